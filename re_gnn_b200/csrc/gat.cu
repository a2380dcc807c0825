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
#include <type_traits>

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
  int d_shift;             // log2(D) when D is a power of two (the common case), else -1: avoids integer divisions
  int hg_count;            // ceil(H*D / 128)
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

constexpr int kUA = 4;  // source rows gathered per lane before any math; kept small so that 5 blocks (40 warps) fit
                         // per SM: with short rows it is resident warps, not loads per warp, that hide the latency chain

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

// Head-group mapping: a row of H*D floats is cut into 128-float slices ("head groups"); ONE WARP owns
// one (row | fragment, head group) item, lane l holding the 128-bit slice [128*hg + 4*l, +4).  Heads are
// independent in both layers, so the HG = ceil(H*D/128) warps of a row never have to talk to each other,
// every gathered source slice is one fully coalesced 512-byte request, and each lane needs a single
// float4 per edge -- which leaves registers for kUA rows in flight per lane.
struct Group {
  int hg, col, hl, h_lo, nh, lph;
  bool ok, leader;
};
__device__ __forceinline__ int div_d(const AttnArgs& a, int x) { return a.d_shift >= 0 ? (x >> a.d_shift) : x / a.D; }
__device__ __forceinline__ Group make_group(const AttnArgs& a, int hg, int lane) {
  Group g;
  const int HD = a.H * a.D;
  g.hg = hg;
  g.col = hg * 128 + lane * 4;
  g.ok = g.col < HD;
  g.h_lo = div_d(a, hg * 128);
  const int h_hi = min(a.H - 1, div_d(a, hg * 128 + 127));
  g.nh = h_hi - g.h_lo + 1;
  g.hl = g.ok ? div_d(a, g.col) : g.h_lo;
  const int lanes_per_head = a.D >> 2;
  g.lph = min(32, lanes_per_head);
  g.leader = g.ok && (g.col - g.hl * a.D == 0);   // first lane of its head
  return g;
}
__device__ __forceinline__ int num_groups(const AttnArgs& a) { return a.hg_count; }
// (item index, head group) of warp item wi; 32-bit arithmetic whenever the item count allows it
__device__ __forceinline__ void split_item(int64_t wi, int HG, int64_t* ri, int* hg) {
  if (wi < 0x7fffffffLL) {
    const unsigned w = (unsigned)wi;
    *ri = w / (unsigned)HG;
    *hg = (int)(w % (unsigned)HG);
  } else {
    *ri = wi / HG;
    *hg = (int)(wi % HG);
  }
}

// =================================================================================================
// REGAT forward.  Logits of a batch of 32 edges are computed one edge per lane (heads of this group);
// batch max / sum through warp shuffles; probabilities staged in shared memory; then the warp
// aggregates the batch with kUA coalesced 128-bit loads in flight per lane.
// Dynamic smem: w_s[R*H] | per warp: p_s[32][HP], sc_s[H], m_s[H]     (HP = H|1: conflict-free)
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 5)
gat_fwd_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, HD = H * a.D, HP = H | 1;
  float* w_s = smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* p_s = smem + a.R * H + warp * (32 * HP + 2 * H);
  float* sc_s = p_s + 32 * HP;
  float* m_s = sc_s + H;
  load_rel_table(w_s, a);
  const int HG = num_groups(a);
  const int64_t wi = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  int64_t ri;
  int hg;
  split_item(wi, HG, &ri, &hg);
  const WorkItem it = decode_item(a, ri, a.nfrag);
  if (!it.ok) return;
  const Group g = make_group(a, hg, lane);
  const int64_t v = it.v;
  const int s0 = it.s0, len = it.len;
  float4 acc = zero4();
  float m_run = -INFINITY, s_run = 0.f;  // lane t (< nh) owns the running max / sum of head h_lo + t
  const float* fcol = a.feat + g.col;
  const bool need_eid = a.keep != nullptr || a.o3 != nullptr;
  const int tl = g.hl - g.h_lo;

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
    for (int t = 0; t < g.nh; ++t) {
      const int h = g.h_lo + t;
      float x = -INFINITY;
      if (valid) {
        float pre = __ldg(a.el + (size_t)idx * H + h) + __ldg(a.er + (size_t)v * H + h);
        if (a.etype != nullptr) pre += w_s[et * H + h];
        x = leaky(pre, a.slope);
        if (a.o3 != nullptr) a.o3[(size_t)e * H + h] = x;  // raw logit; normalised after the row
      }
      const float bm = warp_max(x);
      const float m_old = __shfl_sync(0xffffffffu, m_run, t);
      const float m_new = fmaxf(m_old, bm);
      float p = valid ? expf(x - m_new) : 0.f;
      const float bs = group_sum<32>(p);
      const float sc = expf(m_old - m_new);  // first batch: exp(-inf) = 0
      if (lane == t) {
        m_run = m_new;
        s_run = s_run * sc + bs;
        sc_s[t] = sc;
      }
      if (valid && a.keep != nullptr) p *= __ldg(a.keep + (size_t)e * H + h);
      p_s[lane * HP + t] = p;
    }
    __syncwarp();
    scale4(acc, sc_s[tl]);
    // full groups of kUA edges run without per-edge predicates; the tail group is predicated
    auto body = [&](int j, auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
      float4 x[kUA];
      float p[kUA];
#pragma unroll
      for (int u = 0; u < kUA; ++u) {
        const bool ok = FULL || j + u < cnt;
        const int jj = ok ? j + u : j;
        const int sidx = __shfl_sync(0xffffffffu, idx, jj);
        const bool ld = ok && g.ok;
        x[u] = ld ? ldg4(fcol + (size_t)sidx * HD) : zero4();
        p[u] = ld ? p_s[jj * HP + tl] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kUA; ++u) fma4(acc, p[u], x[u]);
    };
    int j = 0;
    for (; j + kUA <= cnt; j += kUA) body(j, std::true_type{});
    if (j < cnt) body(j, std::false_type{});
    __syncwarp();
  }

  if (it.frag) {  // un-normalised partial + fragment statistics; attn_frag_finalize_kernel merges them
    if (lane < g.nh) {
      a.p1[(size_t)it.fi * H + g.h_lo + lane] = m_run;
      a.p2[(size_t)it.fi * H + g.h_lo + lane] = s_run;
    }
    if (g.ok) st4(a.p0 + (size_t)it.fi * HD + g.col, acc);
    return;
  }
  if (lane < g.nh) {
    const float m = len > 0 ? m_run : 0.f;
    a.o1[(size_t)v * H + g.h_lo + lane] = m;
    a.o2[(size_t)v * H + g.h_lo + lane] = s_run;
    sc_s[lane] = s_run > 0.f ? 1.f / s_run : 0.f;
    m_s[lane] = m;
  }
  __syncwarp();
  if (g.ok) {
    scale4(acc, sc_s[tl]);
    st4(a.o0 + (size_t)v * HD + g.col, acc);
  }
  if (a.o3 != nullptr) {  // get_attention: stored logits -> a*keep, edge-id order (heads of this group)
    for (int i = lane; i < len * g.nh; i += 32) {
      const int t = i % g.nh;
      const size_t o = (size_t)a.eid[s0 + i / g.nh] * H + g.h_lo + t;
      float av = expf(a.o3[o] - m_s[t]) * sc_s[t];
      if (a.keep != nullptr) av *= a.keep[o];
      a.o3[o] = av;
    }
  }
}

// =================================================================================================
// REGAT backward, destination-major.  Per edge: da = <feat[src,h,:], G[v,h,:]>, a recomputed from the
// saved row max / sum, dl = a*keep*da - a*S with S = <out[v,h,:], G[v,h,:]>.
// Two phases per batch of 32 slots, like the forward: (1) one edge per lane recomputes a, a*keep and the
// LeakyReLU slope for the heads of this group (the el[src] gathers and exps are spread over 32 lanes instead
// of serialising on the head-leader lanes); (2) the warp gathers the feat[src] slices, reduces the dots in the
// D/4-lane head groups and the leader lanes finish dpre; (3) each lane writes the values of its edge.
// Dynamic smem: w_s[R*H] | per warp: binsw[R*H] | per warp: pa_s, pt_s, pg_s, dp_s [32][HP]
template <int LPH>   // lanes per head = min(32, D/4), a power of two
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
gat_bwd_dst_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, HD = H * a.D, HP = H | 1, RH = a.etype != nullptr ? a.R * H : 0;
  float* w_s = smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* binsw = smem + RH + warp * RH;
  float* pa_s = smem + RH * (1 + kWarpsPerBlock) + warp * (4 * 32 * HP);
  float* pt_s = pa_s + 32 * HP;
  float* pg_s = pt_s + 32 * HP;
  float* dp_s = pg_s + 32 * HP;
  for (int i = lane; i < RH; i += 32) binsw[i] = 0.f;
  load_rel_table(w_s, a);
  const int HG = num_groups(a);
  const int64_t items = (a.nfrag + (a.row_end - a.row_begin)) * HG;

  for (int64_t wi = (int64_t)blockIdx.x * kWarpsPerBlock + warp; wi < items;
       wi += (int64_t)gridDim.x * kWarpsPerBlock) {
    int64_t ri;
    int hg;
    split_item(wi, HG, &ri, &hg);
    const WorkItem it = decode_item(a, ri, a.nfrag);
    if (!it.ok) continue;
    const Group g = make_group(a, hg, lane);
    const int64_t v = it.v;
    const int s0 = it.s0, len = it.len;
    const int tl = g.hl - g.h_lo;
    const float* fcol = a.feat + g.col;
    float4 gv = zero4();
    float part = 0.f, der = 0.f;
    if (g.ok) {
      gv = ldg4(a.G + (size_t)v * HD + g.col);
      part = dot4(ldg4(a.out + (size_t)v * HD + g.col), gv);
    }
    const float S = group_sum<LPH>(part);
    for (int base = 0; base < len; base += 32) {
      const int cnt = min(32, len - base);
      const bool valid = lane < cnt;
      const int slot = s0 + base + lane;
      int idx = 0, et = 0, e = 0;
      if (valid) {
        idx = a.indices[slot];
        if (a.etype != nullptr) et = a.etype[slot];
        if (a.keep != nullptr) e = a.eid[slot];
      }
      for (int t = 0; t < g.nh; ++t) {  // phase 1: one edge per lane
        const int h = g.h_lo + t;
        float aa = 0.f, at = 0.f, gr = 0.f;
        if (valid) {
          const size_t vh = (size_t)v * H + h;
          float pre = __ldg(a.el + (size_t)idx * H + h) + __ldg(a.er + vh);
          if (a.etype != nullptr) pre += w_s[et * H + h];
          const float sm = __ldg(a.rowsum + vh);
          aa = expf(leaky(pre, a.slope) - __ldg(a.rowmax + vh)) * (sm > 0.f ? 1.f / sm : 0.f);
          at = a.keep != nullptr ? aa * __ldg(a.keep + (size_t)e * H + h) : aa;
          gr = leaky_grad(pre, a.slope);
        }
        pa_s[lane * HP + t] = aa;
        pt_s[lane * HP + t] = at;
        pg_s[lane * HP + t] = gr;
      }
      __syncwarp();
      auto body = [&](int j, auto full_tag) {  // phase 2: gather + per-head dots
        constexpr bool FULL = decltype(full_tag)::value;
        float4 x[kUA];
        float da[kUA];
#pragma unroll
        for (int u = 0; u < kUA; ++u) {
          const bool ok = FULL || j + u < cnt;
          const int sidx = __shfl_sync(0xffffffffu, idx, ok ? j + u : j);
          x[u] = (ok && g.ok) ? ldg4(fcol + (size_t)sidx * HD) : zero4();
        }
#pragma unroll
        for (int u = 0; u < kUA; ++u) da[u] = group_sum<LPH>(dot4(x[u], gv));  // independent reductions: ILP
#pragma unroll
        for (int u = 0; u < kUA; ++u) {
          if (FULL || j + u < cnt) {  // warp-uniform
            const int set = __shfl_sync(0xffffffffu, et, j + u);
            if (g.leader) {
              const int o = (j + u) * HP + tl;
              const float dp = (pt_s[o] * da[u] - pa_s[o] * S) * pg_s[o];
              dp_s[o] = dp;
              der += dp;
              if (a.etype != nullptr) binsw[set * H + g.hl] += dp;
            }
          }
        }
      };
      {
        int j = 0;
        for (; j + kUA <= cnt; j += kUA) body(j, std::true_type{});
        if (j < cnt) body(j, std::false_type{});
      }
      __syncwarp();
      if (valid) {  // phase 3: each lane writes the heads of its own edge
        float* ao = a.o0 + (size_t)slot * H + g.h_lo;
        float* po = a.o1 + (size_t)slot * H + g.h_lo;
        for (int t = 0; t < g.nh; ++t) {
          ao[t] = pt_s[lane * HP + t];
          po[t] = dp_s[lane * HP + t];
        }
      }
      __syncwarp();
    }
    if (g.leader) {
      if (it.frag) a.p1[(size_t)it.fi * H + g.hl] = der;
      else a.o2[(size_t)v * H + g.hl] = der;
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
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 5)
gat_bwd_src_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, HD = H * a.D, HP = H | 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* p_s = smem + warp * (32 * HP);
  const int HG = num_groups(a);
  const int64_t wi = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  int64_t ri;
  int hg;
  split_item(wi, HG, &ri, &hg);
  const WorkItem it = decode_item(a, ri, a.nfrag);
  if (!it.ok) return;
  const Group g = make_group(a, hg, lane);
  const int64_t u_row = it.v;
  const int t0 = it.s0, len = it.len;
  const int tl = g.hl - g.h_lo;
  float4 acc = zero4();
  float del_run = 0.f;
  const float* gcol = a.G + g.col;

  for (int base = 0; base < len; base += 32) {
    const int cnt = min(32, len - base);
    const bool valid = lane < cnt;
    int d = 0, slot = 0;
    if (valid) {
      d = a.indices[t0 + base + lane];
      slot = a.eid[t0 + base + lane];
    }
    for (int t = 0; t < g.nh; ++t) {
      const int h = g.h_lo + t;
      p_s[lane * HP + t] = valid ? __ldg(a.a_csr + (size_t)slot * H + h) : 0.f;
      if (a.d_csr != nullptr) {
        const float tot = group_sum<32>(valid ? __ldg(a.d_csr + (size_t)slot * H + h) : 0.f);
        if (lane == t) del_run += tot;
      }
    }
    __syncwarp();
    auto body = [&](int j, auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
      float4 x[kUA];
      float p[kUA];
#pragma unroll
      for (int u = 0; u < kUA; ++u) {
        const bool ok = FULL || j + u < cnt;
        const int jj = ok ? j + u : j;
        const int sd = __shfl_sync(0xffffffffu, d, jj);
        const bool ld = ok && g.ok;
        x[u] = ld ? ldg4(gcol + (size_t)sd * HD) : zero4();
        p[u] = ld ? p_s[jj * HP + tl] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kUA; ++u) fma4(acc, p[u], x[u]);
    };
    int j = 0;
    for (; j + kUA <= cnt; j += kUA) body(j, std::true_type{});
    if (j < cnt) body(j, std::false_type{});
    __syncwarp();
  }
  if (g.ok) st4((it.frag ? a.p0 + (size_t)it.fi * HD : a.o0 + (size_t)u_row * HD) + g.col, acc);
  if (a.d_csr != nullptr && lane < g.nh) {
    if (it.frag) a.p1[(size_t)it.fi * H + g.h_lo + lane] = del_run;
    else a.o1[(size_t)u_row * H + g.h_lo + lane] = del_run;
  }
}

// =================================================================================================
// REGATv2 forward: the gathered fs[src] slice feeds both the logit (sum_d attn*LeakyReLU(fs+fd),
// reduced by xor-shuffles over the D/4 lanes of a head) and the aggregation; online softmax per
// group of kUA edges.  Nothing [E,H,D]-sized is ever written.
// Dynamic smem: w_s[R*H] | per warp: m_s[H], inv_s[H]
template <int LPH>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 5)
gatv2_fwd_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, HD = H * a.D, RH = a.etype != nullptr ? a.R * H : 0;
  float* w_s = smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* m_s = smem + RH + warp * 2 * H;
  float* inv_s = m_s + H;
  load_rel_table(w_s, a);
  const int HG = num_groups(a);
  const int64_t wi = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  int64_t ri;
  int hg;
  split_item(wi, HG, &ri, &hg);
  const WorkItem it = decode_item(a, ri, a.nfrag);
  if (!it.ok) return;
  const Group g = make_group(a, hg, lane);
  const int64_t v = it.v;
  const int s0 = it.s0, len = it.len;
  float4 acc = zero4(), fdv = zero4(), at = zero4();
  float m = -INFINITY, s = 0.f;
  if (g.ok) {
    fdv = ldg4(a.fd + (size_t)v * HD + g.col);
    at = ldg4(a.el + g.col);
  }
  const float* fcol = a.feat + g.col;
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
    auto body = [&](int j, auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
      float4 x[kUA];
      float l[kUA];
#pragma unroll
      for (int u = 0; u < kUA; ++u) {
        const bool ok = FULL || j + u < cnt;
        const int sidx = __shfl_sync(0xffffffffu, idx, ok ? j + u : j);
        x[u] = (ok && g.ok) ? ldg4(fcol + (size_t)sidx * HD) : zero4();
      }
      float bm = -INFINITY;
#pragma unroll
      for (int u = 0; u < kUA; ++u) {  // kUA independent head reductions
        const bool ok = FULL || j + u < cnt;
        l[u] = group_sum<LPH>(dot4(at, leaky4(add4(x[u], fdv), a.slope)));
        if (a.etype != nullptr) l[u] += w_s[__shfl_sync(0xffffffffu, et, ok ? j + u : j) * H + g.hl];
        if (!ok) l[u] = -INFINITY;
        bm = fmaxf(bm, l[u]);
      }
      const float m_new = fmaxf(m, bm);
      const float sc = expf(m - m_new);
      m = m_new;
      s *= sc;
      scale4(acc, sc);
#pragma unroll
      for (int u = 0; u < kUA; ++u) {
        if (FULL || j + u < cnt) {  // warp-uniform
          int se = 0;
          if (need_eid) se = __shfl_sync(0xffffffffu, e, j + u);
          if (g.ok) {
            if (a.o3 != nullptr && g.leader) a.o3[(size_t)se * H + g.hl] = l[u];
            float p = expf(l[u] - m_new);
            s += p;
            if (a.keep != nullptr) p *= __ldg(a.keep + (size_t)se * H + g.hl);
            fma4(acc, p, x[u]);
          }
        }
      }
    };
    int j = 0;
    for (; j + kUA <= cnt; j += kUA) body(j, std::true_type{});
    if (j < cnt) body(j, std::false_type{});
  }
  if (it.frag) {
    if (g.ok) {
      st4(a.p0 + (size_t)it.fi * HD + g.col, acc);
      if (g.leader) {
        a.p1[(size_t)it.fi * H + g.hl] = m;
        a.p2[(size_t)it.fi * H + g.hl] = s;
      }
    }
    return;
  }
  if (g.ok) {
    const float inv = s > 0.f ? 1.f / s : 0.f;
    const float mm = len > 0 ? m : 0.f;
    scale4(acc, inv);
    st4(a.o0 + (size_t)v * HD + g.col, acc);
    if (g.leader) {
      a.o1[(size_t)v * H + g.hl] = mm;
      a.o2[(size_t)v * H + g.hl] = s;
      m_s[g.hl] = mm;
      inv_s[g.hl] = inv;
    }
  }
  if (a.o3 != nullptr) {
    __syncwarp();
    for (int i = lane; i < len * g.nh; i += 32) {
      const int h = g.h_lo + i % g.nh;
      const size_t o = (size_t)a.eid[s0 + i / g.nh] * H + h;
      float av = expf(a.o3[o] - m_s[h]) * inv_s[h];
      if (a.keep != nullptr) av *= a.keep[o];
      a.o3[o] = av;
    }
  }
}

// =================================================================================================
// REGATv2 backward, destination-major: a_csr, dl_csr, d_fd rows, per-block partials of d_attn and
// of the relation-gradient table.
// Dynamic smem: w_s[R*H] | per warp: binsw[R*H] | per warp: dat_s[128]
template <int LPH>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
gatv2_bwd_dst_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, HD = H * a.D, RH = a.etype != nullptr ? a.R * H : 0;
  const int RHp = (RH + 3) & ~3;  // keeps dat_all 16-byte aligned
  float* w_s = smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* binsw = smem + RHp + warp * RHp;
  float* dat_all = smem + RHp * (1 + kWarpsPerBlock);
  for (int i = lane; i < RH; i += 32) binsw[i] = 0.f;
  load_rel_table(w_s, a);
  const int HG = num_groups(a);
  // every warp keeps ONE head group for all its items (its d_attn slice accumulates in registers): warp wg owns
  // group wg % HG and walks the row items wg / HG, + usable / HG, ...; the last (nW % HG) warps stay idle
  const int64_t nW = (int64_t)gridDim.x * kWarpsPerBlock, usable = nW - nW % HG;
  const int64_t wg = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  const Group g = make_group(a, (int)(wg % HG), lane);
  const int64_t row_items = a.nfrag + (a.row_end - a.row_begin);
  const float* fcol = a.feat + g.col;
  const float4 at = g.ok ? ldg4(a.el + g.col) : zero4();
  float4 dat = zero4();

  for (int64_t ri = wg / HG; wg < usable && ri < row_items; ri += usable / HG) {
    const WorkItem it = decode_item(a, ri, a.nfrag);
    if (!it.ok) continue;
    const int64_t v = it.v;
    const int s0 = it.s0, len = it.len;
    float4 gv = zero4(), fdv = zero4(), dfd = zero4();
    float part = 0.f, m = 0.f, inv = 0.f;
    if (g.ok) {
      gv = ldg4(a.G + (size_t)v * HD + g.col);
      fdv = ldg4(a.fd + (size_t)v * HD + g.col);
      part = dot4(ldg4(a.out + (size_t)v * HD + g.col), gv);
      const size_t vh = (size_t)v * H + g.hl;
      m = a.rowmax[vh];
      const float sm = a.rowsum[vh];
      inv = sm > 0.f ? 1.f / sm : 0.f;
    }
    const float S = group_sum<LPH>(part);
    for (int base = 0; base < len; base += 32) {
      const int cnt = min(32, len - base);
      const int slot = s0 + base + lane;
      int idx = 0, et = 0, e = 0;
      if (lane < cnt) {
        idx = a.indices[slot];
        if (a.etype != nullptr) et = a.etype[slot];
        if (a.keep != nullptr) e = a.eid[slot];
      }
      for (int j = 0; j < cnt; j += kUA) {
        float4 x[kUA];
        float l[kUA], da[kUA];
#pragma unroll
        for (int u = 0; u < kUA; ++u) {
          const int sidx = __shfl_sync(0xffffffffu, idx, min(j + u, cnt - 1));
          x[u] = (j + u < cnt && g.ok) ? ldg4(fcol + (size_t)sidx * HD) : zero4();
        }
#pragma unroll
        for (int u = 0; u < kUA; ++u) {
          l[u] = group_sum<LPH>(dot4(at, leaky4(add4(x[u], fdv), a.slope)));
          da[u] = group_sum<LPH>(dot4(x[u], gv));
        }
#pragma unroll
        for (int u = 0; u < kUA; ++u) {
          if (j + u < cnt) {  // warp-uniform
            const int set = __shfl_sync(0xffffffffu, et, j + u);
            const int se = __shfl_sync(0xffffffffu, e, j + u);
            if (g.ok) {
              const int h = g.hl;
              float lu = l[u];
              if (a.etype != nullptr) lu += w_s[set * H + h];
              const float aa = expf(lu - m) * inv;
              const float att = a.keep != nullptr ? aa * __ldg(a.keep + (size_t)se * H + h) : aa;
              const float dl = att * da[u] - aa * S;
              if (g.leader) {
                const size_t sh = (size_t)(s0 + base + j + u) * H + h;
                a.o0[sh] = att;
                a.o1[sh] = dl;
                if (a.etype != nullptr) binsw[set * H + h] += dl;
              }
              const float4 q = add4(x[u], fdv);
              dfd.x = fmaf(dl * at.x, leaky_grad(q.x, a.slope), dfd.x);
              dfd.y = fmaf(dl * at.y, leaky_grad(q.y, a.slope), dfd.y);
              dfd.z = fmaf(dl * at.z, leaky_grad(q.z, a.slope), dfd.z);
              dfd.w = fmaf(dl * at.w, leaky_grad(q.w, a.slope), dfd.w);
              fma4(dat, dl, leaky4(q, a.slope));
            }
          }
        }
      }
    }
    if (g.ok) st4((it.frag ? a.p0 + (size_t)it.fi * HD : a.o2 + (size_t)v * HD) + g.col, dfd);
  }
  st4(dat_all + (size_t)warp * 128 + lane * 4, dat);
  __syncthreads();
  double* outp = a.partials + (size_t)blockIdx.x * a.partial_stride;
  for (int i = threadIdx.x; i < RH; i += blockDim.x) {
    double sum = 0.0;
    for (int w = 0; w < kWarpsPerBlock; ++w) sum += (double)smem[RHp + w * RHp + i];
    outp[i] = sum;
  }
  for (int i = threadIdx.x; i < HD; i += blockDim.x) {
    double sum = 0.0;
    for (int w = 0; w < kWarpsPerBlock; ++w)  // warps of this block that own head group i/128, in warp order
      if (((int64_t)blockIdx.x * kWarpsPerBlock + w) % HG == i / 128) sum += (double)dat_all[w * 128 + (i & 127)];
    outp[RH + i] = sum;
  }
}

// =================================================================================================
// REGATv2 backward, source-major: d_fs[u] = sum_j a_csr*G[dst] + dl_csr*attn*LeakyReLU'(fs[u]+fd[dst]).
// Dynamic smem per warp: p_s[32][HP], q_s[32][HP]
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 5)
gatv2_bwd_src_kernel(AttnArgs a) {
  constexpr int U = kUA / 2;  // two gathered rows per edge
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, HD = H * a.D, HP = H | 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* p_s = smem + warp * (64 * HP);
  float* q_s = p_s + 32 * HP;
  const int HG = num_groups(a);
  const int64_t wi = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  int64_t ri;
  int hg;
  split_item(wi, HG, &ri, &hg);
  const WorkItem it = decode_item(a, ri, a.nfrag);
  if (!it.ok) return;
  const Group g = make_group(a, hg, lane);
  const int64_t u_row = it.v;
  const int t0 = it.s0, len = it.len;
  const int tl = g.hl - g.h_lo;
  float4 acc = zero4(), fsu = zero4(), at = zero4();
  if (g.ok) {
    fsu = ldg4(a.feat + (size_t)u_row * HD + g.col);
    at = ldg4(a.el + g.col);
  }
  const float* gcol = a.G + g.col;
  const float* dcol = a.fd + g.col;

  for (int base = 0; base < len; base += 32) {
    const int cnt = min(32, len - base);
    const bool valid = lane < cnt;
    int d = 0, slot = 0;
    if (valid) {
      d = a.indices[t0 + base + lane];
      slot = a.eid[t0 + base + lane];
    }
    for (int t = 0; t < g.nh; ++t) {
      const int h = g.h_lo + t;
      p_s[lane * HP + t] = valid ? __ldg(a.a_csr + (size_t)slot * H + h) : 0.f;
      q_s[lane * HP + t] = valid ? __ldg(a.d_csr + (size_t)slot * H + h) : 0.f;
    }
    __syncwarp();
    for (int j = 0; j < cnt; j += U) {
      float4 xg[U], xd[U];
      float p[U], q[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = j + u < cnt;
        const int jj = ok ? j + u : j;
        const int sd = __shfl_sync(0xffffffffu, d, jj);
        const bool ld = ok && g.ok;
        xg[u] = ld ? ldg4(gcol + (size_t)sd * HD) : zero4();
        xd[u] = ld ? ldg4(dcol + (size_t)sd * HD) : zero4();
        p[u] = ld ? p_s[jj * HP + tl] : 0.f;
        q[u] = ld ? q_s[jj * HP + tl] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        fma4(acc, p[u], xg[u]);
        const float4 z = add4(fsu, xd[u]);
        acc.x = fmaf(q[u] * at.x, leaky_grad(z.x, a.slope), acc.x);
        acc.y = fmaf(q[u] * at.y, leaky_grad(z.y, a.slope), acc.y);
        acc.z = fmaf(q[u] * at.z, leaky_grad(z.z, a.slope), acc.z);
        acc.w = fmaf(q[u] * at.w, leaky_grad(z.w, a.slope), acc.w);
      }
    }
    __syncwarp();
  }
  if (g.ok) st4((it.frag ? a.p0 + (size_t)it.fi * HD : a.o0 + (size_t)u_row * HD) + g.col, acc);
}

// =================================================================================================
// Merges the fragments of every long row of a forward pass (online-softmax combine in fragment order):
//   M = max_f m_f;  S = sum_f s_f*exp(m_f-M);  out = sum_f acc_f*exp(m_f-M) / S
// and, for get_attention, normalises the row's stored logits.  One warp per (long row, head group).
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
attn_frag_finalize_kernel(AttnArgs a, const int32_t* __restrict__ long_rows,
                          const int32_t* __restrict__ frag_ptr, int num_long) {
  const int H = a.H, HD = H * a.D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int HG = num_groups(a);
  const int64_t wi = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  const int l = (int)(wi / HG);
  if (l >= num_long) return;
  const int64_t v = long_rows[l];
  if (v < a.row_begin || v >= a.row_end) return;
  const Group g = make_group(a, (int)(wi % HG), lane);
  const int f0 = frag_ptr[l], f1 = frag_ptr[l + 1];
  const size_t hh = g.hl;  // every lane tracks the statistics of its own head (redundant within a head)
  float M = -INFINITY;
  for (int f = f0; f < f1; ++f) M = fmaxf(M, a.p1[(size_t)f * H + hh]);
  float S = 0.f;
  float4 acc = zero4();
  for (int f = f0; f < f1; ++f) {
    const float sc = expf(a.p1[(size_t)f * H + hh] - M);
    S += a.p2[(size_t)f * H + hh] * sc;
    if (g.ok) fma4(acc, sc, ldg4(a.p0 + (size_t)f * HD + g.col));
  }
  const float inv = S > 0.f ? 1.f / S : 0.f;
  if (g.ok) {
    scale4(acc, inv);
    st4(a.o0 + (size_t)v * HD + g.col, acc);
    if (g.leader) {
      a.o1[(size_t)v * H + hh] = M;
      a.o2[(size_t)v * H + hh] = S;
    }
  }
  if (a.o3 != nullptr) {
    __syncwarp();
    const int s0 = a.indptr[v], len = a.indptr[v + 1] - s0;
    for (int i = lane; i < len * g.nh; i += 32) {
      const int h = g.h_lo + i % g.nh;
      const size_t o = (size_t)a.eid[s0 + i / g.nh] * H + h;
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
  a.hg_count = (a.H * a.D + 127) / 128;
  a.d_shift = -1;
  for (int b = 2; b <= 10; ++b)
    if ((1 << b) == a.D) a.d_shift = b;
  a.nfrag = a.nfrag_pad = 0;
  a.threshold = 0x7fffffff;
  if (split == nullptr || split->num_frags <= 0) return true;
  if (!ws || !split->long_rows || !split->frag_ptr || !split->frag_row || !split->frag_begin || split->threshold <= 0)
    return false;
  a.frag_row = split->frag_row;
  a.frag_begin = split->frag_begin;
  a.nfrag = split->num_frags;
  a.nfrag_pad = a.nfrag;
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

#define REGNN_DISPATCH_C(KERNEL, GRID, SMEM)                                  \
  do {                                                                        \
    int rc_ = set_smem(KERNEL, SMEM);                                         \
    if (rc_ != REGNN_OK) return rc_;                                          \
    KERNEL<<<GRID, kWarpsPerBlock * 32, SMEM, stream>>>(a);                   \
  } while (0)

#define REGNN_DISPATCH_FINALIZE(GRID) \
  attn_frag_finalize_kernel<<<GRID, kWarpsPerBlock * 32, 0, stream>>>(a, split->long_rows, split->frag_ptr, split->num_long)

static int head_groups(int H, int D) { return (H * D + 127) / 128; }

// kernels that reduce over the D/4 lanes of a head are compiled per lane count (fully unrolled butterflies)
#define REGNN_DISPATCH_LPH(KERNEL, GRID, SMEM)                                          \
  do {                                                                                  \
    const int lph_ = a.D / 4 >= 32 ? 32 : a.D / 4;                                      \
    switch (lph_) {                                                                     \
      case 1: REGNN_DISPATCH_C(KERNEL<1>, GRID, SMEM); break;                           \
      case 2: REGNN_DISPATCH_C(KERNEL<2>, GRID, SMEM); break;                           \
      case 4: REGNN_DISPATCH_C(KERNEL<4>, GRID, SMEM); break;                           \
      case 8: REGNN_DISPATCH_C(KERNEL<8>, GRID, SMEM); break;                           \
      case 16: REGNN_DISPATCH_C(KERNEL<16>, GRID, SMEM); break;                         \
      default: REGNN_DISPATCH_C(KERNEL<32>, GRID, SMEM); break;                         \
    }                                                                                   \
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
  REGNN_REQUIRE(indptr && feat && el && er && out && rowmax && rowsum, REGNN_ERR_INVALID_ARG, "gat_fwd: null pointer");
  REGNN_REQUIRE((keep == nullptr && attn_out == nullptr) || eid != nullptr, REGNN_ERR_INVALID_ARG, "gat_fwd: keep/attn_out need eid");
  REGNN_REQUIRE(etype_csr == nullptr || theta != nullptr, REGNN_ERR_INVALID_ARG, "gat_fwd: etype without theta");
  int rc = check_shape("gat_fwd", num_heads, head_dim, num_relations, etype_csr != nullptr, true);  // as the backward: fail early
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
  const unsigned grid = (unsigned)(((rows + a.nfrag) * head_groups(a.H, a.D) + kWarpsPerBlock - 1) / kWarpsPerBlock);
  REGNN_DISPATCH_C(gat_fwd_kernel, grid, smem);
  if (a.nfrag > 0) {
    const unsigned fgrid = (unsigned)(((int64_t)split->num_long * head_groups(a.H, a.D) + kWarpsPerBlock - 1) / kWarpsPerBlock);
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
  REGNN_REQUIRE(indptr && feat && el && er && out && rowmax && rowsum && Gd && d_er,
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
  const size_t smem = sizeof(float) * ((size_t)RH * (1 + kWarpsPerBlock) + (size_t)kWarpsPerBlock * 4 * 32 * (num_heads | 1)) + 16;
  const int nb = partial_blocks((rows + a.nfrag) * head_groups(a.H, a.D));
  REGNN_DISPATCH_LPH(gat_bwd_dst_kernel, nb, smem);
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
  REGNN_REQUIRE(indptr_t && Gd && d_feat, REGNN_ERR_INVALID_ARG, "gat_bwd_src: null pointer");
  REGNN_REQUIRE(dpre_csr == nullptr || d_el != nullptr, REGNN_ERR_INVALID_ARG, "gat_bwd_src: dpre_csr without d_el");
  int rc = check_shape("gat_bwd_src", num_heads, head_dim, 0, false, true);
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
  const unsigned grid = (unsigned)(((rows + a.nfrag) * head_groups(a.H, a.D) + kWarpsPerBlock - 1) / kWarpsPerBlock);
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
  REGNN_REQUIRE(indptr && fs && fd && attn && out && rowmax && rowsum, REGNN_ERR_INVALID_ARG, "gatv2_fwd: null pointer");
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
  const unsigned grid = (unsigned)(((rows + a.nfrag) * head_groups(a.H, a.D) + kWarpsPerBlock - 1) / kWarpsPerBlock);
  REGNN_DISPATCH_LPH(gatv2_fwd_kernel, grid, smem);
  if (a.nfrag > 0) {
    const unsigned fgrid = (unsigned)(((int64_t)split->num_long * head_groups(a.H, a.D) + kWarpsPerBlock - 1) / kWarpsPerBlock);
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
  REGNN_REQUIRE(indptr && fs && fd && attn && out && rowmax && rowsum && Gd && d_fd && d_attn && partials,
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
  const size_t smem = sizeof(float) * ((size_t)((RH + 3) & ~3) * (1 + kWarpsPerBlock) + (size_t)kWarpsPerBlock * 128) + 16;
  REGNN_REQUIRE(apply_split(a, split, split_workspace), REGNN_ERR_INVALID_ARG, "gatv2_bwd_dst: incomplete row split");
  const int nb = partial_blocks((rows + a.nfrag) * head_groups(a.H, a.D));
  REGNN_DISPATCH_LPH(gatv2_bwd_dst_kernel, nb, smem);
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
  REGNN_REQUIRE(indptr_t && fs && fd && attn && Gd && d_fs,
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
  const unsigned grid = (unsigned)(((rows + a.nfrag) * head_groups(a.H, a.D) + kWarpsPerBlock - 1) / kWarpsPerBlock);
  REGNN_DISPATCH_C(gatv2_bwd_src_kernel, grid, smem);
  if (a.nfrag > 0) launch_rowsum(split, a.p0, num_heads * head_dim, d_fs, row_begin, row_end, stream);
  return check_launch("regnn_gatv2_bwd_src");
}
