// Fused REGAT / REGATv2 kernels: edge logits + LeakyReLU + per-destination online softmax +
// weighted aggregation in ONE pass over the in-edges of each destination row, and the two-pass
// deterministic backward (destination-major, then source-major over the transposed view).
// Reference call sites: layer/REGATConv.py:71-92, layer/REGATv2Conv.py:133-152 (DGL: gsddmm(add),
// 4-kernel edge_softmax, broadcast gspmm; autograd: gspmm on the reverse graph + gsddmm(dot)).
//
// Mapping: one warp per destination row.  A row of H*D floats is split into 128-bit slices; lane l
// owns slices l, l+32, ... (C per lane), so every gathered source row is read with fully coalesced
// LDG.128.  Slice k of lane l belongs to head (4*(l+32k))/D.  Per-head dot products are reduced
// with xor-shuffles inside the aligned group of D/4 lanes that covers a head.  Gather-bound: one
// H*D*4-byte source row per edge (two in the GATv2 source-major backward pass).
#include "common.cuh"

namespace regnn {

struct AttnArgs {
  const int32_t* indptr;   // CSR (dst-major) or transposed (src-major) row pointers
  const int32_t* indices;
  const int32_t* eid;      // slot -> edge id (dst-major)  |  slot_t (src-major)
  const uint8_t* etype;
  const float* theta;      // [R,H]
  float alpha;
  int R;
  const float* feat;       // GAT: feat [N,H,D]   GATv2: fs
  const float* fd;         // GATv2 only
  const float* el;         // GAT: el [N,H]       GATv2: attn [H,D]
  const float* er;         // GAT: er [N,H]
  float slope;
  const float* keep;       // [E,H] edge-id order or null
  const float* out;
  const float* rowmax;
  const float* rowsum;
  const float* G;
  int H, D;
  int64_t row_begin, row_end;
  float* o0;               // fwd: out      bwd_dst: a_csr     bwd_src: d_feat / d_fs
  float* o1;               // fwd: rowmax   bwd_dst: dpre/dl   bwd_src: d_el
  float* o2;               // fwd: rowsum   bwd_dst: d_er / d_fd
  float* o3;               // fwd: attn_out
  double* partials;
  int partial_stride;
  const float* a_csr;      // bwd_src inputs
  const float* d_csr;
  // long-row fragments (regnn_rowsplit_t): items [0, nfrag) are fragments (padded to nfrag_pad in the
  // grid-mapped kernels); their partial results go to p0/p1/p2 and are merged by the finalize kernels
  const int32_t* frag_row;
  const int32_t* frag_begin;
  int nfrag, nfrag_pad, threshold;
  float* p0;               // [nfrag][H*D]  partial rows (fwd: un-normalised acc; bwd: d_feat / d_fd / d_fs)
  float* p1;               // [nfrag][H]    fwd: fragment max      bwd: d_er / d_el partial
  float* p2;               // [nfrag][H]    fwd: fragment sum
};

struct WorkItem {
  int64_t v, fi;
  int s0, len;
  bool ok, frag, first;
};

// Item wi of a kernel whose first `nfrag_items` items are long-row fragments and the rest ordinary rows.
__device__ __forceinline__ WorkItem decode_item(const AttnArgs& a, int64_t wi, int64_t nfrag_items) {
  WorkItem w{0, wi, 0, 0, false, false, true};
  if (wi < nfrag_items) {
    w.frag = true;
    if (wi < a.nfrag) {
      w.v = a.frag_row[wi];
      if (w.v >= a.row_begin && w.v < a.row_end) {
        w.s0 = a.frag_begin[wi];
        w.len = min(a.threshold, a.indptr[w.v + 1] - w.s0);
        w.first = w.s0 == a.indptr[w.v];
        w.ok = true;
      }
    }
  } else {
    w.v = a.row_begin + (wi - nfrag_items);
    if (w.v < a.row_end) {
      w.s0 = a.indptr[w.v];
      w.len = a.indptr[w.v + 1] - w.s0;
      w.ok = w.len <= a.threshold;  // longer rows are covered by fragments
    }
  }
  return w;
}

template <int C> struct UnrollA { static constexpr int U = C <= 1 ? 4 : (C <= 4 ? 2 : 1); };

__device__ __forceinline__ void load_rel_table(float* w_s, const AttnArgs& a) {
  if (a.etype != nullptr) {
    const int n = a.R * a.H;
    for (int i = threadIdx.x; i < n; i += blockDim.x) w_s[i] = leaky(a.theta[i] * a.alpha, kRelationSlope);
  }
  __syncthreads();
}

__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 leaky4(float4 q, float s) {
  return make_float4(leaky(q.x, s), leaky(q.y, s), leaky(q.z, s), leaky(q.w, s));
}
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// Slice bookkeeping shared by all kernels.
template <int C>
struct Slices {
  int head[C];
  bool ok[C];
  bool leader[C];  // first lane of the head's lane group: writes per-(row|edge, head) scalars
  __device__ __forceinline__ Slices(int lane, int H, int D) {
    const int lph = D >> 2;
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const int c4 = lane + 32 * k;
      ok[k] = c4 * 4 < H * D;
      head[k] = ok[k] ? (c4 * 4) / D : 0;
      leader[k] = ok[k] && (c4 % lph == 0);
    }
  }
};

// =================================================================================================
// REGAT forward.  Logits of a batch of 32 edges are computed one edge per lane (all heads); the
// batch max / sum go through warp shuffles; probabilities are staged in shared memory; then the
// whole warp aggregates the batch edge by edge with C coalesced 128-bit loads per lane per edge.
// Dynamic smem: w_s[R*H] | per warp: p_s[32][HP], sc_s[H], m_s[H]     (HP = H|1: conflict-free)
template <int C>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
gat_fwd_kernel(AttnArgs a) {
  constexpr int U = UnrollA<C>::U;
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, D = a.D, HD = H * D, HP = H | 1;
  float* w_s = smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* p_s = smem + a.R * H + warp * (32 * HP + 2 * H);
  float* sc_s = p_s + 32 * HP;
  float* m_s = sc_s + H;
  load_rel_table(w_s, a);
  const WorkItem it = decode_item(a, (int64_t)blockIdx.x * kWarpsPerBlock + warp, a.nfrag_pad);
  if (!it.ok) return;
  const int64_t v = it.v;
  const int s0 = it.s0, len = it.len;
  const Slices<C> sl(lane, H, D);
  float4 acc[C];
#pragma unroll
  for (int k = 0; k < C; ++k) acc[k] = zero4();
  float m_run = -INFINITY, s_run = 0.f;  // lane h (< H) owns the running max / sum of head h
  const float* fcol = a.feat + (size_t)lane * 4;
  const bool need_eid = a.keep != nullptr || a.o3 != nullptr;

  for (int base = 0; base < len; base += 32) {
    const int cnt = min(32, len - base);
    const bool valid = lane < cnt;
    const int slot = s0 + base + lane;
    int idx = 0, et = 0, e = 0;
    if (valid) {
      idx = a.indices[slot];
      if (a.etype != nullptr) et = a.etype[slot];
      if (need_eid) e = a.eid[slot];
    }
    for (int h = 0; h < H; ++h) {
      float x = -INFINITY;
      if (valid) {
        float pre = __ldg(a.el + (size_t)idx * H + h) + __ldg(a.er + (size_t)v * H + h);
        if (a.etype != nullptr) pre += w_s[et * H + h];
        x = leaky(pre, a.slope);
        if (a.o3 != nullptr) a.o3[(size_t)e * H + h] = x;  // raw logit; normalised after the row
      }
      const float bm = warp_max(x);
      const float m_old = __shfl_sync(0xffffffffu, m_run, h);
      const float m_new = fmaxf(m_old, bm);
      float p = valid ? expf(x - m_new) : 0.f;
      const float bs = group_sum<32>(p);
      const float sc = expf(m_old - m_new);  // first batch: exp(-inf) = 0
      if (lane == h) {
        m_run = m_new;
        s_run = s_run * sc + bs;
        sc_s[h] = sc;
      }
      if (valid && a.keep != nullptr) p *= __ldg(a.keep + (size_t)e * H + h);
      p_s[lane * HP + h] = p;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < C; ++k) scale4(acc[k], sc_s[sl.head[k]]);
    for (int j = 0; j < cnt; j += U) {
      float4 x[U][C];
      float p[U][C];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = j + u < cnt;
        const int jj = ok ? j + u : j;
        const int sidx = __shfl_sync(0xffffffffu, idx, jj);
#pragma unroll
        for (int k = 0; k < C; ++k) {
          const bool ld = ok && sl.ok[k];
          x[u][k] = ld ? ldg4(fcol + (size_t)sidx * HD + (size_t)k * 128) : zero4();
          p[u][k] = ld ? p_s[jj * HP + sl.head[k]] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < C; ++k) fma4(acc[k], p[u][k], x[u][k]);
    }
    __syncwarp();
  }

  if (it.frag) {  // un-normalised partial + fragment statistics; attn_frag_finalize_kernel merges them
    if (lane < H) {
      a.p1[(size_t)it.fi * H + lane] = m_run;
      a.p2[(size_t)it.fi * H + lane] = s_run;
    }
    float* pcol = a.p0 + (size_t)it.fi * HD + (size_t)lane * 4;
#pragma unroll
    for (int k = 0; k < C; ++k)
      if (sl.ok[k]) st4(pcol + (size_t)k * 128, acc[k]);
    return;
  }
  if (lane < H) {
    const float m = len > 0 ? m_run : 0.f;
    a.o1[(size_t)v * H + lane] = m;
    a.o2[(size_t)v * H + lane] = s_run;
    sc_s[lane] = s_run > 0.f ? 1.f / s_run : 0.f;
    m_s[lane] = m;
  }
  __syncwarp();
  float* ocol = a.o0 + (size_t)v * HD + (size_t)lane * 4;
#pragma unroll
  for (int k = 0; k < C; ++k)
    if (sl.ok[k]) {
      scale4(acc[k], sc_s[sl.head[k]]);
      st4(ocol + (size_t)k * 128, acc[k]);
    }
  if (a.o3 != nullptr) {  // get_attention: stored logits -> a*keep, edge-id order
    for (int i = lane; i < len * H; i += 32) {
      const int h = i % H;
      const size_t o = (size_t)a.eid[s0 + i / H] * H + h;
      float av = expf(a.o3[o] - m_s[h]) * sc_s[h];
      if (a.keep != nullptr) av *= a.keep[o];
      a.o3[o] = av;
    }
  }
}

// =================================================================================================
// REGAT backward, destination-major.  Per edge: da = <feat[src,h,:], G[v,h,:]>, recomputed a from
// the saved row max / sum, dl = a*keep*da - a*S with S = <out[v,h,:], G[v,h,:]>.
// Dynamic smem: w_s[R*H] | per warp: binsw[R*H]
template <int C>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
gat_bwd_dst_kernel(AttnArgs a) {
  constexpr int U = UnrollA<C>::U;
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, D = a.D, HD = H * D, RH = a.etype != nullptr ? a.R * H : 0, lph = D >> 2;
  float* w_s = smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* binsw = smem + RH + warp * RH;
  for (int i = lane; i < RH; i += 32) binsw[i] = 0.f;
  load_rel_table(w_s, a);
  const Slices<C> sl(lane, H, D);
  const float* fcol = a.feat + (size_t)lane * 4;
  const int64_t items = a.nfrag + (a.row_end - a.row_begin);

  for (int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp; r < items;
       r += (int64_t)gridDim.x * kWarpsPerBlock) {
    const WorkItem it = decode_item(a, r, a.nfrag);
    if (!it.ok) continue;
    const int64_t v = it.v;
    const int s0 = it.s0, len = it.len;
    float4 g[C];
    float S[C], m[C], inv[C], erv[C], der[C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
      float part = 0.f;
      g[k] = zero4();
      m[k] = inv[k] = erv[k] = 0.f;
      der[k] = 0.f;
      if (sl.ok[k]) {
        const size_t c = (size_t)v * HD + (size_t)(lane + 32 * k) * 4;
        g[k] = ldg4(a.G + c);
        part = dot4(ldg4(a.out + c), g[k]);
        const size_t vh = (size_t)v * H + sl.head[k];
        m[k] = a.rowmax[vh];
        const float s = a.rowsum[vh];
        inv[k] = s > 0.f ? 1.f / s : 0.f;
        erv[k] = a.er[vh];
      }
      S[k] = group_sum_rt(part, lph);
    }
    for (int base = 0; base < len; base += 32) {
      const int cnt = min(32, len - base);
      const int slot = s0 + base + lane;
      int idx = 0, et = 0, e = 0;
      if (lane < cnt) {
        idx = a.indices[slot];
        if (a.etype != nullptr) et = a.etype[slot];
        if (a.keep != nullptr) e = a.eid[slot];
      }
      for (int j = 0; j < cnt; j += U) {
        float4 x[U][C];
        int sidx[U], set[U], se[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jj = min(j + u, cnt - 1);
          sidx[u] = __shfl_sync(0xffffffffu, idx, jj);
          set[u] = __shfl_sync(0xffffffffu, et, jj);
          se[u] = __shfl_sync(0xffffffffu, e, jj);
#pragma unroll
          for (int k = 0; k < C; ++k)
            x[u][k] = (j + u < cnt && sl.ok[k]) ? ldg4(fcol + (size_t)sidx[u] * HD + (size_t)k * 128) : zero4();
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (j + u < cnt) {  // warp-uniform
            const size_t sh = (size_t)(s0 + base + j + u) * H;
#pragma unroll
            for (int k = 0; k < C; ++k) {
              const float da = group_sum_rt(dot4(x[u][k], g[k]), lph);
              if (sl.ok[k]) {
                const int h = sl.head[k];
                float pre = __ldg(a.el + (size_t)sidx[u] * H + h) + erv[k];
                if (a.etype != nullptr) pre += w_s[set[u] * H + h];
                const float aa = expf(leaky(pre, a.slope) - m[k]) * inv[k];
                const float at = a.keep != nullptr ? aa * __ldg(a.keep + (size_t)se[u] * H + h) : aa;
                const float dp = (at * da - aa * S[k]) * leaky_grad(pre, a.slope);
                if (sl.leader[k]) {
                  a.o0[sh + h] = at;
                  a.o1[sh + h] = dp;
                  der[k] += dp;
                  if (a.etype != nullptr) binsw[set[u] * H + h] += dp;
                }
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < C; ++k)
      if (sl.leader[k]) {
        if (it.frag) a.p1[(size_t)it.fi * H + sl.head[k]] = der[k];
        else a.o2[(size_t)v * H + sl.head[k]] = der[k];
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RH; i += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < kWarpsPerBlock; ++w) s += (double)smem[RH + w * RH + i];
    a.partials[(size_t)blockIdx.x * a.partial_stride + i] = s;
  }
}

// =================================================================================================
// Source-major aggregation with precomputed per-slot, per-head weights (REGAT backward w.r.t. feat,
// and the el-gradient reduction):  d_feat[u] = sum_j a_csr[slot_t[j]] * G[indices_t[j]].
// Dynamic smem per warp: p_s[32][HP]
template <int C>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
gat_bwd_src_kernel(AttnArgs a) {
  constexpr int U = UnrollA<C>::U;
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, D = a.D, HD = H * D, HP = H | 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* p_s = smem + warp * (32 * HP);
  const WorkItem it = decode_item(a, (int64_t)blockIdx.x * kWarpsPerBlock + warp, a.nfrag_pad);
  if (!it.ok) return;
  const int64_t u_row = it.v;
  const int t0 = it.s0, len = it.len;
  const Slices<C> sl(lane, H, D);
  float4 acc[C];
#pragma unroll
  for (int k = 0; k < C; ++k) acc[k] = zero4();
  float del_run = 0.f;
  const float* gcol = a.G + (size_t)lane * 4;

  for (int base = 0; base < len; base += 32) {
    const int cnt = min(32, len - base);
    const bool valid = lane < cnt;
    int d = 0, slot = 0;
    if (valid) {
      d = a.indices[t0 + base + lane];
      slot = a.eid[t0 + base + lane];
    }
    for (int h = 0; h < H; ++h) {
      p_s[lane * HP + h] = valid ? __ldg(a.a_csr + (size_t)slot * H + h) : 0.f;
      if (a.d_csr != nullptr) {
        const float tot = group_sum<32>(valid ? __ldg(a.d_csr + (size_t)slot * H + h) : 0.f);
        if (lane == h) del_run += tot;
      }
    }
    __syncwarp();
    for (int j = 0; j < cnt; j += U) {
      float4 x[U][C];
      float p[U][C];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = j + u < cnt;
        const int jj = ok ? j + u : j;
        const int sd = __shfl_sync(0xffffffffu, d, jj);
#pragma unroll
        for (int k = 0; k < C; ++k) {
          const bool ld = ok && sl.ok[k];
          x[u][k] = ld ? ldg4(gcol + (size_t)sd * HD + (size_t)k * 128) : zero4();
          p[u][k] = ld ? p_s[jj * HP + sl.head[k]] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < C; ++k) fma4(acc[k], p[u][k], x[u][k]);
    }
    __syncwarp();
  }
  float* ocol = it.frag ? a.p0 + (size_t)it.fi * HD + (size_t)lane * 4 : a.o0 + (size_t)u_row * HD + (size_t)lane * 4;
#pragma unroll
  for (int k = 0; k < C; ++k)
    if (sl.ok[k]) st4(ocol + (size_t)k * 128, acc[k]);
  if (a.d_csr != nullptr && lane < H) {
    if (it.frag) a.p1[(size_t)it.fi * H + lane] = del_run;
    else a.o1[(size_t)u_row * H + lane] = del_run;
  }
}

// =================================================================================================
// REGATv2 forward: per edge the gathered fs[src] row feeds both the logit
// (sum_d attn*LeakyReLU(fs+fd)) and the aggregation; online softmax per edge.
// Dynamic smem: w_s[R*H] | per warp: m_s[H], inv_s[H]
template <int C>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
gatv2_fwd_kernel(AttnArgs a) {
  constexpr int U = UnrollA<C>::U;
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, D = a.D, HD = H * D, RH = a.etype != nullptr ? a.R * H : 0, lph = D >> 2;
  float* w_s = smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* m_s = smem + RH + warp * 2 * H;
  float* inv_s = m_s + H;
  load_rel_table(w_s, a);
  const WorkItem it = decode_item(a, (int64_t)blockIdx.x * kWarpsPerBlock + warp, a.nfrag_pad);
  if (!it.ok) return;
  const int64_t v = it.v;
  const int s0 = it.s0, len = it.len;
  const Slices<C> sl(lane, H, D);
  float4 acc[C], fdv[C], at[C];
  float m[C], s[C];
#pragma unroll
  for (int k = 0; k < C; ++k) {
    acc[k] = fdv[k] = at[k] = zero4();
    m[k] = -INFINITY;
    s[k] = 0.f;
    if (sl.ok[k]) {
      fdv[k] = ldg4(a.fd + (size_t)v * HD + (size_t)(lane + 32 * k) * 4);
      at[k] = ldg4(a.el + (size_t)(lane + 32 * k) * 4);
    }
  }
  const float* fcol = a.feat + (size_t)lane * 4;
  const bool need_eid = a.keep != nullptr || a.o3 != nullptr;

  for (int base = 0; base < len; base += 32) {
    const int cnt = min(32, len - base);
    const int slot = s0 + base + lane;
    int idx = 0, et = 0, e = 0;
    if (lane < cnt) {
      idx = a.indices[slot];
      if (a.etype != nullptr) et = a.etype[slot];
      if (need_eid) e = a.eid[slot];
    }
    for (int j = 0; j < cnt; j += U) {
      float4 x[U][C];
      int set[U], se[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int jj = min(j + u, cnt - 1);
        const int sidx = __shfl_sync(0xffffffffu, idx, jj);
        set[u] = __shfl_sync(0xffffffffu, et, jj);
        se[u] = __shfl_sync(0xffffffffu, e, jj);
#pragma unroll
        for (int k = 0; k < C; ++k)
          x[u][k] = (j + u < cnt && sl.ok[k]) ? ldg4(fcol + (size_t)sidx * HD + (size_t)k * 128) : zero4();
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j + u < cnt) {  // warp-uniform
#pragma unroll
          for (int k = 0; k < C; ++k) {
            const float part = dot4(at[k], leaky4(add4(x[u][k], fdv[k]), a.slope));
            float l = group_sum_rt(part, lph);
            if (sl.ok[k]) {
              const int h = sl.head[k];
              if (a.etype != nullptr) l += w_s[set[u] * H + h];
              if (a.o3 != nullptr && sl.leader[k]) a.o3[(size_t)se[u] * H + h] = l;
              const float m_new = fmaxf(m[k], l);
              const float sc = expf(m[k] - m_new);
              float p = expf(l - m_new);
              s[k] = s[k] * sc + p;
              m[k] = m_new;
              if (a.keep != nullptr) p *= __ldg(a.keep + (size_t)se[u] * H + h);
              scale4(acc[k], sc);
              fma4(acc[k], p, x[u][k]);
            }
          }
        }
      }
    }
  }
  if (it.frag) {
#pragma unroll
    for (int k = 0; k < C; ++k)
      if (sl.ok[k]) {
        st4(a.p0 + (size_t)it.fi * HD + (size_t)(lane + 32 * k) * 4, acc[k]);
        if (sl.leader[k]) {
          a.p1[(size_t)it.fi * H + sl.head[k]] = m[k];
          a.p2[(size_t)it.fi * H + sl.head[k]] = s[k];
        }
      }
    return;
  }
#pragma unroll
  for (int k = 0; k < C; ++k) {
    if (sl.ok[k]) {
      const float inv = s[k] > 0.f ? 1.f / s[k] : 0.f;
      const float mm = len > 0 ? m[k] : 0.f;
      scale4(acc[k], inv);
      st4(a.o0 + (size_t)v * HD + (size_t)(lane + 32 * k) * 4, acc[k]);
      if (sl.leader[k]) {
        const int h = sl.head[k];
        a.o1[(size_t)v * H + h] = mm;
        a.o2[(size_t)v * H + h] = s[k];
        m_s[h] = mm;
        inv_s[h] = inv;
      }
    }
  }
  if (a.o3 != nullptr) {
    __syncwarp();
    for (int i = lane; i < len * H; i += 32) {
      const int h = i % H;
      const size_t o = (size_t)a.eid[s0 + i / H] * H + h;
      float av = expf(a.o3[o] - m_s[h]) * inv_s[h];
      if (a.keep != nullptr) av *= a.keep[o];
      a.o3[o] = av;
    }
  }
}

// =================================================================================================
// REGATv2 backward, destination-major: a_csr, dl_csr, d_fd rows, per-block partials of d_attn and
// of the relation-gradient table.
// Dynamic smem: w_s[R*H] | per warp: binsw[R*H] | per warp: dat_s[H*D]
template <int C>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
gatv2_bwd_dst_kernel(AttnArgs a) {
  constexpr int U = UnrollA<C>::U;
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, D = a.D, HD = H * D, RH = a.etype != nullptr ? a.R * H : 0, lph = D >> 2;
  float* w_s = smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int RHp = (RH + 3) & ~3;  // keeps dat_all 16-byte aligned
  float* binsw = smem + RHp + warp * RHp;
  float* dat_all = smem + RHp * (1 + kWarpsPerBlock);
  for (int i = lane; i < RH; i += 32) binsw[i] = 0.f;
  load_rel_table(w_s, a);
  const Slices<C> sl(lane, H, D);
  const float* fcol = a.feat + (size_t)lane * 4;
  const int64_t items = a.nfrag + (a.row_end - a.row_begin);
  float4 at[C], dat[C];
#pragma unroll
  for (int k = 0; k < C; ++k) {
    dat[k] = zero4();
    at[k] = sl.ok[k] ? ldg4(a.el + (size_t)(lane + 32 * k) * 4) : zero4();
  }

  for (int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp; r < items;
       r += (int64_t)gridDim.x * kWarpsPerBlock) {
    const WorkItem it = decode_item(a, r, a.nfrag);
    if (!it.ok) continue;
    const int64_t v = it.v;
    const int s0 = it.s0, len = it.len;
    float4 g[C], fdv[C], dfd[C];
    float S[C], m[C], inv[C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
      float part = 0.f;
      g[k] = fdv[k] = dfd[k] = zero4();
      m[k] = inv[k] = 0.f;
      if (sl.ok[k]) {
        const size_t c = (size_t)v * HD + (size_t)(lane + 32 * k) * 4;
        g[k] = ldg4(a.G + c);
        fdv[k] = ldg4(a.fd + c);
        part = dot4(ldg4(a.out + c), g[k]);
        const size_t vh = (size_t)v * H + sl.head[k];
        m[k] = a.rowmax[vh];
        const float s = a.rowsum[vh];
        inv[k] = s > 0.f ? 1.f / s : 0.f;
      }
      S[k] = group_sum_rt(part, lph);
    }
    for (int base = 0; base < len; base += 32) {
      const int cnt = min(32, len - base);
      const int slot = s0 + base + lane;
      int idx = 0, et = 0, e = 0;
      if (lane < cnt) {
        idx = a.indices[slot];
        if (a.etype != nullptr) et = a.etype[slot];
        if (a.keep != nullptr) e = a.eid[slot];
      }
      for (int j = 0; j < cnt; j += U) {
        float4 x[U][C];
        int set[U], se[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jj = min(j + u, cnt - 1);
          const int sidx = __shfl_sync(0xffffffffu, idx, jj);
          set[u] = __shfl_sync(0xffffffffu, et, jj);
          se[u] = __shfl_sync(0xffffffffu, e, jj);
#pragma unroll
          for (int k = 0; k < C; ++k)
            x[u][k] = (j + u < cnt && sl.ok[k]) ? ldg4(fcol + (size_t)sidx * HD + (size_t)k * 128) : zero4();
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (j + u < cnt) {  // warp-uniform
            const size_t sh = (size_t)(s0 + base + j + u) * H;
#pragma unroll
            for (int k = 0; k < C; ++k) {
              const float4 q = add4(x[u][k], fdv[k]);
              const float4 lr = leaky4(q, a.slope);
              float l = group_sum_rt(dot4(at[k], lr), lph);
              const float da = group_sum_rt(dot4(x[u][k], g[k]), lph);
              if (sl.ok[k]) {
                const int h = sl.head[k];
                if (a.etype != nullptr) l += w_s[set[u] * H + h];
                const float aa = expf(l - m[k]) * inv[k];
                const float att = a.keep != nullptr ? aa * __ldg(a.keep + (size_t)se[u] * H + h) : aa;
                const float dl = att * da - aa * S[k];
                if (sl.leader[k]) {
                  a.o0[sh + h] = att;
                  a.o1[sh + h] = dl;
                  if (a.etype != nullptr) binsw[set[u] * H + h] += dl;
                }
                dfd[k].x = fmaf(dl * at[k].x, leaky_grad(q.x, a.slope), dfd[k].x);
                dfd[k].y = fmaf(dl * at[k].y, leaky_grad(q.y, a.slope), dfd[k].y);
                dfd[k].z = fmaf(dl * at[k].z, leaky_grad(q.z, a.slope), dfd[k].z);
                dfd[k].w = fmaf(dl * at[k].w, leaky_grad(q.w, a.slope), dfd[k].w);
                fma4(dat[k], dl, lr);
              }
            }
          }
        }
      }
    }
    float* drow = it.frag ? a.p0 + (size_t)it.fi * HD : a.o2 + (size_t)v * HD;
#pragma unroll
    for (int k = 0; k < C; ++k)
      if (sl.ok[k]) st4(drow + (size_t)(lane + 32 * k) * 4, dfd[k]);
  }
#pragma unroll
  for (int k = 0; k < C; ++k)
    if (sl.ok[k]) st4(dat_all + (size_t)warp * HD + (size_t)(lane + 32 * k) * 4, dat[k]);
  __syncthreads();
  double* outp = a.partials + (size_t)blockIdx.x * a.partial_stride;
  for (int i = threadIdx.x; i < RH; i += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < kWarpsPerBlock; ++w) s += (double)smem[RHp + w * RHp + i];
    outp[i] = s;
  }
  for (int i = threadIdx.x; i < HD; i += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < kWarpsPerBlock; ++w) s += (double)dat_all[w * HD + i];
    outp[RH + i] = s;
  }
}

// =================================================================================================
// REGATv2 backward, source-major: d_fs[u] = sum_j a_csr*G[dst] + dl_csr*attn*LeakyReLU'(fs[u]+fd[dst]).
// Dynamic smem per warp: p_s[32][HP], q_s[32][HP]
template <int C>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
gatv2_bwd_src_kernel(AttnArgs a) {
  constexpr int U = C <= 2 ? 2 : 1;
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, D = a.D, HD = H * D, HP = H | 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* p_s = smem + warp * (64 * HP);
  float* q_s = p_s + 32 * HP;
  const WorkItem it = decode_item(a, (int64_t)blockIdx.x * kWarpsPerBlock + warp, a.nfrag_pad);
  if (!it.ok) return;
  const int64_t u_row = it.v;
  const int t0 = it.s0, len = it.len;
  const Slices<C> sl(lane, H, D);
  float4 acc[C], fsu[C], at[C];
#pragma unroll
  for (int k = 0; k < C; ++k) {
    acc[k] = fsu[k] = at[k] = zero4();
    if (sl.ok[k]) {
      fsu[k] = ldg4(a.feat + (size_t)u_row * HD + (size_t)(lane + 32 * k) * 4);
      at[k] = ldg4(a.el + (size_t)(lane + 32 * k) * 4);
    }
  }
  const float* gcol = a.G + (size_t)lane * 4;
  const float* dcol = a.fd + (size_t)lane * 4;

  for (int base = 0; base < len; base += 32) {
    const int cnt = min(32, len - base);
    const bool valid = lane < cnt;
    int d = 0, slot = 0;
    if (valid) {
      d = a.indices[t0 + base + lane];
      slot = a.eid[t0 + base + lane];
    }
    for (int h = 0; h < H; ++h) {
      p_s[lane * HP + h] = valid ? __ldg(a.a_csr + (size_t)slot * H + h) : 0.f;
      q_s[lane * HP + h] = valid ? __ldg(a.d_csr + (size_t)slot * H + h) : 0.f;
    }
    __syncwarp();
    for (int j = 0; j < cnt; j += U) {
      float4 xg[U][C], xd[U][C];
      float p[U][C], q[U][C];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = j + u < cnt;
        const int jj = ok ? j + u : j;
        const int sd = __shfl_sync(0xffffffffu, d, jj);
#pragma unroll
        for (int k = 0; k < C; ++k) {
          const bool ld = ok && sl.ok[k];
          xg[u][k] = ld ? ldg4(gcol + (size_t)sd * HD + (size_t)k * 128) : zero4();
          xd[u][k] = ld ? ldg4(dcol + (size_t)sd * HD + (size_t)k * 128) : zero4();
          p[u][k] = ld ? p_s[jj * HP + sl.head[k]] : 0.f;
          q[u][k] = ld ? q_s[jj * HP + sl.head[k]] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < C; ++k) {
          fma4(acc[k], p[u][k], xg[u][k]);
          const float4 z = add4(fsu[k], xd[u][k]);
          acc[k].x = fmaf(q[u][k] * at[k].x, leaky_grad(z.x, a.slope), acc[k].x);
          acc[k].y = fmaf(q[u][k] * at[k].y, leaky_grad(z.y, a.slope), acc[k].y);
          acc[k].z = fmaf(q[u][k] * at[k].z, leaky_grad(z.z, a.slope), acc[k].z);
          acc[k].w = fmaf(q[u][k] * at[k].w, leaky_grad(z.w, a.slope), acc[k].w);
        }
    }
    __syncwarp();
  }
  float* orow = it.frag ? a.p0 + (size_t)it.fi * HD : a.o0 + (size_t)u_row * HD;
#pragma unroll
  for (int k = 0; k < C; ++k)
    if (sl.ok[k]) st4(orow + (size_t)(lane + 32 * k) * 4, acc[k]);
}

// =================================================================================================
// Merges the fragments of every long row of a forward pass (online-softmax combine in fragment order):
//   M = max_f m_f;  S = sum_f s_f*exp(m_f-M);  out = sum_f acc_f*exp(m_f-M) / S
// and, for get_attention, normalises the row's stored logits.  One warp per long row.
template <int C>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
attn_frag_finalize_kernel(AttnArgs a, const int32_t* __restrict__ long_rows,
                          const int32_t* __restrict__ frag_ptr, int num_long) {
  const int H = a.H, D = a.D, HD = H * D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l = blockIdx.x * kWarpsPerBlock + warp;
  if (l >= num_long) return;
  const int64_t v = long_rows[l];
  if (v < a.row_begin || v >= a.row_end) return;
  const int f0 = frag_ptr[l], f1 = frag_ptr[l + 1];
  const Slices<C> sl(lane, H, D);
  float M = -INFINITY;  // lane h (< H) owns head h
  if (lane < H)
    for (int f = f0; f < f1; ++f) M = fmaxf(M, a.p1[(size_t)f * H + lane]);
  float S = 0.f;
  float4 acc[C];
#pragma unroll
  for (int k = 0; k < C; ++k) acc[k] = zero4();
  for (int f = f0; f < f1; ++f) {
    float sc = 0.f;
    if (lane < H) {
      sc = expf(a.p1[(size_t)f * H + lane] - M);
      S += a.p2[(size_t)f * H + lane] * sc;
    }
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const float sk = __shfl_sync(0xffffffffu, sc, sl.head[k]);
      if (sl.ok[k]) fma4(acc[k], sk, ldg4(a.p0 + (size_t)f * HD + (size_t)(lane + 32 * k) * 4));
    }
  }
  const float inv = S > 0.f ? 1.f / S : 0.f;
  if (lane < H) {
    a.o1[(size_t)v * H + lane] = M;
    a.o2[(size_t)v * H + lane] = S;
  }
#pragma unroll
  for (int k = 0; k < C; ++k) {
    const float ik = __shfl_sync(0xffffffffu, inv, sl.head[k]);
    if (sl.ok[k]) {
      scale4(acc[k], ik);
      st4(a.o0 + (size_t)v * HD + (size_t)(lane + 32 * k) * 4, acc[k]);
    }
  }
  if (a.o3 != nullptr) {
    __syncwarp();
    const int s0 = a.indptr[v], len = a.indptr[v + 1] - s0;
    for (int i = lane; i < len * H; i += 32) {
      const int h = i % H;
      const size_t o = (size_t)a.eid[s0 + i / H] * H + h;
      const float s = a.o2[(size_t)v * H + h];
      float av = expf(a.o3[o] - a.o1[(size_t)v * H + h]) * (s > 0.f ? 1.f / s : 0.f);
      if (a.keep != nullptr) av *= a.keep[o];
      a.o3[o] = av;
    }
  }
}

// out[v][c] = sum over the fragments of long row v of partial[f][c]  (fragment order; block per row)
__global__ void frag_rowsum_kernel(const int32_t* __restrict__ long_rows, const int32_t* __restrict__ frag_ptr,
                                   int num_long, const float* __restrict__ partial, int W,
                                   float* __restrict__ out, int64_t row_begin, int64_t row_end) {
  const int l = blockIdx.x;
  if (l >= num_long) return;
  const int64_t v = long_rows[l];
  if (v < row_begin || v >= row_end) return;
  const int f0 = frag_ptr[l], f1 = frag_ptr[l + 1];
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    float s = 0.f;
    for (int f = f0; f < f1; ++f) s += partial[(size_t)f * W + c];
    out[(size_t)v * W + c] = s;
  }
}

// ---- dispatch ---------------------------------------------------------------------------------------
static bool aligned16(const void* p) { return p == nullptr || ((uintptr_t)p & 15) == 0; }

static int pick_c(int HD) {
  const int c = (HD / 4 + 31) / 32;
  if (c <= 1) return 1;
  if (c <= 2) return 2;
  if (c <= 4) return 4;
  if (c <= 8) return 8;
  return 0;
}

// Fills the fragment fields of `a` from the caller's row split; returns false if it is incomplete.
static bool apply_split(AttnArgs& a, const regnn_rowsplit_t* split, float* ws) {
  a.nfrag = a.nfrag_pad = 0;
  a.threshold = 0x7fffffff;
  if (split == nullptr || split->num_frags <= 0) return true;
  if (!ws || !split->long_rows || !split->frag_ptr || !split->frag_row || !split->frag_begin || split->threshold <= 0)
    return false;
  a.frag_row = split->frag_row;
  a.frag_begin = split->frag_begin;
  a.nfrag = split->num_frags;
  a.nfrag_pad = (a.nfrag + kWarpsPerBlock - 1) / kWarpsPerBlock * kWarpsPerBlock;
  a.threshold = split->threshold;
  const size_t HD = (size_t)a.H * a.D;
  a.p0 = ws;
  a.p1 = ws + (size_t)a.nfrag * HD;
  a.p2 = a.p1 + (size_t)a.nfrag * a.H;
  return true;
}

static void launch_rowsum(const regnn_rowsplit_t* split, const float* partial, int W, float* out, int64_t rb,
                          int64_t re, cudaStream_t stream) {
  frag_rowsum_kernel<<<split->num_long, 128, 0, stream>>>(split->long_rows, split->frag_ptr, split->num_long, partial,
                                                          W, out, rb, re);
}

static int check_shape(const char* who, int H, int D, int R, bool has_rel, bool needs_dot) {
  REGNN_REQUIRE(H >= 1 && H <= REGNN_MAX_HEADS, REGNN_ERR_UNSUPPORTED_SHAPE, "%s: num_heads=%d outside [1,%d]", who, H, REGNN_MAX_HEADS);
  REGNN_REQUIRE(D >= 4 && D % 4 == 0, REGNN_ERR_UNSUPPORTED_SHAPE, "%s: head_dim=%d must be a positive multiple of 4", who, D);
  REGNN_REQUIRE(pick_c(H * D) != 0, REGNN_ERR_UNSUPPORTED_SHAPE, "%s: H*D=%d exceeds 1024", who, H * D);
  if (needs_dot)
    REGNN_REQUIRE(D <= 128 && (D & (D - 1)) == 0, REGNN_ERR_UNSUPPORTED_SHAPE,
                  "%s: head_dim=%d must be a power of two in [4,128]", who, D);
  if (has_rel)
    REGNN_REQUIRE(R >= 1 && R <= REGNN_MAX_RELATIONS && R * H <= 4096, REGNN_ERR_UNSUPPORTED_SHAPE,
                  "%s: num_relations=%d (x %d heads) unsupported", who, R, H);
  return REGNN_OK;
}

#define REGNN_DISPATCH_C(KERNEL, GRID, SMEM)                                           \
  do {                                                                                 \
    int rc_ = REGNN_OK;                                                                \
    switch (pick_c(a.H * a.D)) {                                                       \
      case 1: rc_ = set_smem(KERNEL<1>, SMEM); if (rc_ == REGNN_OK) KERNEL<1><<<GRID, kWarpsPerBlock * 32, SMEM, stream>>>(a); break; \
      case 2: rc_ = set_smem(KERNEL<2>, SMEM); if (rc_ == REGNN_OK) KERNEL<2><<<GRID, kWarpsPerBlock * 32, SMEM, stream>>>(a); break; \
      case 4: rc_ = set_smem(KERNEL<4>, SMEM); if (rc_ == REGNN_OK) KERNEL<4><<<GRID, kWarpsPerBlock * 32, SMEM, stream>>>(a); break; \
      default: rc_ = set_smem(KERNEL<8>, SMEM); if (rc_ == REGNN_OK) KERNEL<8><<<GRID, kWarpsPerBlock * 32, SMEM, stream>>>(a); break; \
    }                                                                                  \
    if (rc_ != REGNN_OK) return rc_;                                                   \
  } while (0)

#define REGNN_DISPATCH_FINALIZE(GRID)                                                                                  \
  do {                                                                                                                 \
    switch (pick_c(a.H * a.D)) {                                                                                       \
      case 1: attn_frag_finalize_kernel<1><<<GRID, kWarpsPerBlock * 32, 0, stream>>>(a, split->long_rows, split->frag_ptr, split->num_long); break; \
      case 2: attn_frag_finalize_kernel<2><<<GRID, kWarpsPerBlock * 32, 0, stream>>>(a, split->long_rows, split->frag_ptr, split->num_long); break; \
      case 4: attn_frag_finalize_kernel<4><<<GRID, kWarpsPerBlock * 32, 0, stream>>>(a, split->long_rows, split->frag_ptr, split->num_long); break; \
      default: attn_frag_finalize_kernel<8><<<GRID, kWarpsPerBlock * 32, 0, stream>>>(a, split->long_rows, split->frag_ptr, split->num_long); break; \
    }                                                                                                                  \
  } while (0)

}  // namespace regnn

using namespace regnn;

extern "C" int regnn_gat_fwd(const int32_t* indptr, const int32_t* indices, const int32_t* eid,
                             const uint8_t* etype_csr, const float* theta, float alpha,
                             int num_relations, const float* feat, const float* el, const float* er,
                             float negative_slope, const float* keep, int num_heads, int head_dim,
                             int64_t row_begin, int64_t row_end, float* out, float* rowmax,
                             float* rowsum, float* attn_out, const regnn_rowsplit_t* split, float* split_workspace,
    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  REGNN_REQUIRE(indptr && indices && feat && el && er && out && rowmax && rowsum, REGNN_ERR_INVALID_ARG, "gat_fwd: null pointer");
  REGNN_REQUIRE((keep == nullptr && attn_out == nullptr) || eid != nullptr, REGNN_ERR_INVALID_ARG, "gat_fwd: keep/attn_out need eid");
  REGNN_REQUIRE(etype_csr == nullptr || theta != nullptr, REGNN_ERR_INVALID_ARG, "gat_fwd: etype without theta");
  int rc = check_shape("gat_fwd", num_heads, head_dim, num_relations, etype_csr != nullptr, false);
  if (rc != REGNN_OK) return rc;
  REGNN_REQUIRE(aligned16(feat) && aligned16(out), REGNN_ERR_INVALID_ARG, "gat_fwd: feat/out must be 16-byte aligned");
  const int64_t rows = row_end - row_begin;
  REGNN_REQUIRE(rows >= 0, REGNN_ERR_INVALID_ARG, "gat_fwd: negative row range");
  if (rows == 0) return REGNN_OK;
  AttnArgs a{};
  a.indptr = indptr; a.indices = indices; a.eid = eid; a.etype = etype_csr; a.theta = theta; a.alpha = alpha;
  a.R = etype_csr ? num_relations : 0; a.feat = feat; a.el = el; a.er = er; a.slope = negative_slope; a.keep = keep;
  a.H = num_heads; a.D = head_dim; a.row_begin = row_begin; a.row_end = row_end;
  a.o0 = out; a.o1 = rowmax; a.o2 = rowsum; a.o3 = attn_out;
  REGNN_REQUIRE(apply_split(a, split, split_workspace), REGNN_ERR_INVALID_ARG, "gat_fwd: incomplete row split");
  const int H = num_heads, HP = H | 1;
  const size_t smem = sizeof(float) * ((size_t)a.R * H + (size_t)kWarpsPerBlock * (32 * HP + 2 * H));
  const unsigned grid = (unsigned)((rows + a.nfrag_pad + kWarpsPerBlock - 1) / kWarpsPerBlock);
  REGNN_DISPATCH_C(gat_fwd_kernel, grid, smem);
  if (a.nfrag > 0) {
    const unsigned fgrid = (unsigned)((split->num_long + kWarpsPerBlock - 1) / kWarpsPerBlock);
    REGNN_DISPATCH_FINALIZE(fgrid);
  }
  return check_launch("regnn_gat_fwd");
}

extern "C" int regnn_gat_bwd_dst(const int32_t* indptr, const int32_t* indices, const int32_t* eid,
                                 const uint8_t* etype_csr, const float* theta, float alpha,
                                 int num_relations, const float* feat, const float* el,
                                 const float* er, float negative_slope, const float* keep,
                                 const float* out, const float* rowmax, const float* rowsum,
                                 const float* Gd, int num_heads, int head_dim, int64_t row_begin,
                                 int64_t row_end, float* a_csr, float* dpre_csr, float* d_er,
                                 double* partials, float* d_theta, const regnn_rowsplit_t* split, float* split_workspace,
    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  REGNN_REQUIRE(indptr && indices && feat && el && er && out && rowmax && rowsum && Gd && a_csr && dpre_csr && d_er,
                REGNN_ERR_INVALID_ARG, "gat_bwd_dst: null pointer");
  REGNN_REQUIRE(keep == nullptr || eid != nullptr, REGNN_ERR_INVALID_ARG, "gat_bwd_dst: keep needs eid");
  REGNN_REQUIRE(etype_csr == nullptr || (theta && partials && d_theta), REGNN_ERR_INVALID_ARG, "gat_bwd_dst: null relation buffers");
  int rc = check_shape("gat_bwd_dst", num_heads, head_dim, num_relations, etype_csr != nullptr, true);
  if (rc != REGNN_OK) return rc;
  REGNN_REQUIRE(aligned16(feat) && aligned16(out) && aligned16(Gd), REGNN_ERR_INVALID_ARG, "gat_bwd_dst: 16-byte alignment required");
  const int64_t rows = row_end - row_begin;
  REGNN_REQUIRE(rows >= 0, REGNN_ERR_INVALID_ARG, "gat_bwd_dst: negative row range");
  AttnArgs a{};
  a.indptr = indptr; a.indices = indices; a.eid = eid; a.etype = etype_csr; a.theta = theta; a.alpha = alpha;
  a.R = etype_csr ? num_relations : 0; a.feat = feat; a.el = el; a.er = er; a.slope = negative_slope; a.keep = keep;
  a.out = out; a.rowmax = rowmax; a.rowsum = rowsum; a.G = Gd; a.H = num_heads; a.D = head_dim;
  a.row_begin = row_begin; a.row_end = row_end; a.o0 = a_csr; a.o1 = dpre_csr; a.o2 = d_er;
  a.partials = partials; a.partial_stride = a.R * num_heads;
  REGNN_REQUIRE(apply_split(a, split, split_workspace), REGNN_ERR_INVALID_ARG, "gat_bwd_dst: incomplete row split");
  const int RH = a.R * num_heads;
  const size_t smem = sizeof(float) * ((size_t)RH * (1 + kWarpsPerBlock)) + 16;
  const int nb = partial_blocks(rows + a.nfrag);
  REGNN_DISPATCH_C(gat_bwd_dst_kernel, nb, smem);
  if (a.nfrag > 0) launch_rowsum(split, a.p1, num_heads, d_er, row_begin, row_end, stream);
  if (etype_csr != nullptr) launch_relation_grad_finalize(partials, nb, RH, RH, theta, alpha, d_theta, stream);
  return check_launch("regnn_gat_bwd_dst");
}

extern "C" int regnn_gat_bwd_src(const int32_t* indptr_t, const int32_t* indices_t,
                                 const int32_t* slot_t, const float* a_csr, const float* dpre_csr,
                                 const float* Gd, int num_heads, int head_dim, int64_t row_begin,
                                 int64_t row_end, float* d_feat, float* d_el, const regnn_rowsplit_t* split, float* split_workspace,
    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  REGNN_REQUIRE(indptr_t && indices_t && slot_t && a_csr && Gd && d_feat, REGNN_ERR_INVALID_ARG, "gat_bwd_src: null pointer");
  REGNN_REQUIRE(dpre_csr == nullptr || d_el != nullptr, REGNN_ERR_INVALID_ARG, "gat_bwd_src: dpre_csr without d_el");
  int rc = check_shape("gat_bwd_src", num_heads, head_dim, 0, false, false);
  if (rc != REGNN_OK) return rc;
  REGNN_REQUIRE(aligned16(Gd) && aligned16(d_feat), REGNN_ERR_INVALID_ARG, "gat_bwd_src: 16-byte alignment required");
  const int64_t rows = row_end - row_begin;
  REGNN_REQUIRE(rows >= 0, REGNN_ERR_INVALID_ARG, "gat_bwd_src: negative row range");
  if (rows == 0) return REGNN_OK;
  AttnArgs a{};
  a.indptr = indptr_t; a.indices = indices_t; a.eid = slot_t; a.a_csr = a_csr; a.d_csr = dpre_csr; a.G = Gd;
  a.H = num_heads; a.D = head_dim; a.row_begin = row_begin; a.row_end = row_end; a.o0 = d_feat; a.o1 = d_el;
  REGNN_REQUIRE(apply_split(a, split, split_workspace), REGNN_ERR_INVALID_ARG, "gat_bwd_src: incomplete row split");
  const int HP = num_heads | 1;
  const size_t smem = sizeof(float) * (size_t)kWarpsPerBlock * 32 * HP;
  const unsigned grid = (unsigned)((rows + a.nfrag_pad + kWarpsPerBlock - 1) / kWarpsPerBlock);
  REGNN_DISPATCH_C(gat_bwd_src_kernel, grid, smem);
  if (a.nfrag > 0) {
    launch_rowsum(split, a.p0, num_heads * head_dim, d_feat, row_begin, row_end, stream);
    if (dpre_csr != nullptr) launch_rowsum(split, a.p1, num_heads, d_el, row_begin, row_end, stream);
  }
  return check_launch("regnn_gat_bwd_src");
}

extern "C" int regnn_gatv2_fwd(const int32_t* indptr, const int32_t* indices, const int32_t* eid,
                               const uint8_t* etype_csr, const float* theta, float alpha,
                               int num_relations, const float* fs, const float* fd,
                               const float* attn, float negative_slope, const float* keep,
                               int num_heads, int head_dim, int64_t row_begin, int64_t row_end,
                               float* out, float* rowmax, float* rowsum, float* attn_out,
                               const regnn_rowsplit_t* split, float* split_workspace,
    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  REGNN_REQUIRE(indptr && indices && fs && fd && attn && out && rowmax && rowsum, REGNN_ERR_INVALID_ARG, "gatv2_fwd: null pointer");
  REGNN_REQUIRE((keep == nullptr && attn_out == nullptr) || eid != nullptr, REGNN_ERR_INVALID_ARG, "gatv2_fwd: keep/attn_out need eid");
  REGNN_REQUIRE(etype_csr == nullptr || theta != nullptr, REGNN_ERR_INVALID_ARG, "gatv2_fwd: etype without theta");
  int rc = check_shape("gatv2_fwd", num_heads, head_dim, num_relations, etype_csr != nullptr, true);
  if (rc != REGNN_OK) return rc;
  REGNN_REQUIRE(aligned16(fs) && aligned16(fd) && aligned16(attn) && aligned16(out), REGNN_ERR_INVALID_ARG, "gatv2_fwd: 16-byte alignment required");
  const int64_t rows = row_end - row_begin;
  REGNN_REQUIRE(rows >= 0, REGNN_ERR_INVALID_ARG, "gatv2_fwd: negative row range");
  if (rows == 0) return REGNN_OK;
  AttnArgs a{};
  a.indptr = indptr; a.indices = indices; a.eid = eid; a.etype = etype_csr; a.theta = theta; a.alpha = alpha;
  a.R = etype_csr ? num_relations : 0; a.feat = fs; a.fd = fd; a.el = attn; a.slope = negative_slope; a.keep = keep;
  a.H = num_heads; a.D = head_dim; a.row_begin = row_begin; a.row_end = row_end;
  a.o0 = out; a.o1 = rowmax; a.o2 = rowsum; a.o3 = attn_out;
  REGNN_REQUIRE(apply_split(a, split, split_workspace), REGNN_ERR_INVALID_ARG, "gatv2_fwd: incomplete row split");
  const size_t smem = sizeof(float) * ((size_t)a.R * num_heads + (size_t)kWarpsPerBlock * 2 * num_heads) + 16;
  const unsigned grid = (unsigned)((rows + a.nfrag_pad + kWarpsPerBlock - 1) / kWarpsPerBlock);
  REGNN_DISPATCH_C(gatv2_fwd_kernel, grid, smem);
  if (a.nfrag > 0) {
    const unsigned fgrid = (unsigned)((split->num_long + kWarpsPerBlock - 1) / kWarpsPerBlock);
    REGNN_DISPATCH_FINALIZE(fgrid);
  }
  return check_launch("regnn_gatv2_fwd");
}

extern "C" int regnn_gatv2_bwd_dst(const int32_t* indptr, const int32_t* indices, const int32_t* eid,
                                   const uint8_t* etype_csr, const float* theta, float alpha,
                                   int num_relations, const float* fs, const float* fd,
                                   const float* attn, float negative_slope, const float* keep,
                                   const float* out, const float* rowmax, const float* rowsum,
                                   const float* Gd, int num_heads, int head_dim, int64_t row_begin,
                                   int64_t row_end, float* a_csr, float* dl_csr, float* d_fd,
                                   float* d_attn, double* partials, float* d_theta, const regnn_rowsplit_t* split, float* split_workspace,
    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  REGNN_REQUIRE(indptr && indices && fs && fd && attn && out && rowmax && rowsum && Gd && a_csr && dl_csr && d_fd && d_attn && partials,
                REGNN_ERR_INVALID_ARG, "gatv2_bwd_dst: null pointer");
  REGNN_REQUIRE(keep == nullptr || eid != nullptr, REGNN_ERR_INVALID_ARG, "gatv2_bwd_dst: keep needs eid");
  REGNN_REQUIRE(etype_csr == nullptr || (theta && d_theta), REGNN_ERR_INVALID_ARG, "gatv2_bwd_dst: null relation buffers");
  int rc = check_shape("gatv2_bwd_dst", num_heads, head_dim, num_relations, etype_csr != nullptr, true);
  if (rc != REGNN_OK) return rc;
  REGNN_REQUIRE(aligned16(fs) && aligned16(fd) && aligned16(attn) && aligned16(out) && aligned16(Gd) && aligned16(d_fd),
                REGNN_ERR_INVALID_ARG, "gatv2_bwd_dst: 16-byte alignment required");
  const int64_t rows = row_end - row_begin;
  REGNN_REQUIRE(rows >= 0, REGNN_ERR_INVALID_ARG, "gatv2_bwd_dst: negative row range");
  AttnArgs a{};
  a.indptr = indptr; a.indices = indices; a.eid = eid; a.etype = etype_csr; a.theta = theta; a.alpha = alpha;
  a.R = etype_csr ? num_relations : 0; a.feat = fs; a.fd = fd; a.el = attn; a.slope = negative_slope; a.keep = keep;
  a.out = out; a.rowmax = rowmax; a.rowsum = rowsum; a.G = Gd; a.H = num_heads; a.D = head_dim;
  a.row_begin = row_begin; a.row_end = row_end; a.o0 = a_csr; a.o1 = dl_csr; a.o2 = d_fd;
  const int RH = a.R * num_heads, HD = num_heads * head_dim;
  a.partials = partials; a.partial_stride = RH + HD;
  const size_t smem = sizeof(float) * ((size_t)((RH + 3) & ~3) * (1 + kWarpsPerBlock) + (size_t)kWarpsPerBlock * HD) + 16;
  REGNN_REQUIRE(apply_split(a, split, split_workspace), REGNN_ERR_INVALID_ARG, "gatv2_bwd_dst: incomplete row split");
  const int nb = partial_blocks(rows + a.nfrag);
  REGNN_DISPATCH_C(gatv2_bwd_dst_kernel, nb, smem);
  if (a.nfrag > 0) launch_rowsum(split, a.p0, HD, d_fd, row_begin, row_end, stream);
  if (etype_csr != nullptr) launch_relation_grad_finalize(partials, nb, RH + HD, RH, theta, alpha, d_theta, stream);
  launch_colsum_finalize(partials, nb, RH + HD, RH, HD, d_attn, stream);
  return check_launch("regnn_gatv2_bwd_dst");
}

extern "C" int regnn_gatv2_bwd_src(const int32_t* indptr_t, const int32_t* indices_t,
                                   const int32_t* slot_t, const float* a_csr, const float* dl_csr,
                                   const float* fs, const float* fd, const float* attn,
                                   float negative_slope, const float* Gd, int num_heads,
                                   int head_dim, int64_t row_begin, int64_t row_end, float* d_fs,
                                   const regnn_rowsplit_t* split, float* split_workspace,
    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  REGNN_REQUIRE(indptr_t && indices_t && slot_t && a_csr && dl_csr && fs && fd && attn && Gd && d_fs,
                REGNN_ERR_INVALID_ARG, "gatv2_bwd_src: null pointer");
  int rc = check_shape("gatv2_bwd_src", num_heads, head_dim, 0, false, false);
  if (rc != REGNN_OK) return rc;
  REGNN_REQUIRE(aligned16(fs) && aligned16(fd) && aligned16(attn) && aligned16(Gd) && aligned16(d_fs),
                REGNN_ERR_INVALID_ARG, "gatv2_bwd_src: 16-byte alignment required");
  const int64_t rows = row_end - row_begin;
  REGNN_REQUIRE(rows >= 0, REGNN_ERR_INVALID_ARG, "gatv2_bwd_src: negative row range");
  if (rows == 0) return REGNN_OK;
  AttnArgs a{};
  a.indptr = indptr_t; a.indices = indices_t; a.eid = slot_t; a.a_csr = a_csr; a.d_csr = dl_csr;
  a.feat = fs; a.fd = fd; a.el = attn; a.slope = negative_slope; a.G = Gd;
  a.H = num_heads; a.D = head_dim; a.row_begin = row_begin; a.row_end = row_end; a.o0 = d_fs;
  const int HP = num_heads | 1;
  REGNN_REQUIRE(apply_split(a, split, split_workspace), REGNN_ERR_INVALID_ARG, "gatv2_bwd_src: incomplete row split");
  const size_t smem = sizeof(float) * (size_t)kWarpsPerBlock * 64 * HP;
  const unsigned grid = (unsigned)((rows + a.nfrag_pad + kWarpsPerBlock - 1) / kWarpsPerBlock);
  REGNN_DISPATCH_C(gatv2_bwd_src_kernel, grid, smem);
  if (a.nfrag > 0) launch_rowsum(split, a.p0, num_heads * head_dim, d_fs, row_begin, row_end, stream);
  return check_launch("regnn_gatv2_bwd_src");
}
