// Grouped per-node-type input projection: out[i,:] = W[type(i)] * x_{type(i)}[row(i),:] + b[type(i)] for all node types in
// ONE launch, and its weight / bias gradients.
// Reference call sites: model/REGCN.py:36-39, model/REGAT.py:56-58, model/REMixHop.py:89-92 (`fc_list[t](features_list[t])`
// per type, then torch.cat) and mag/regnn_ns.py:300-326 (`group_input`: per type a boolean mask, an index_select of the
// type's feature table, a Linear and a masked scatter -- T kernels chains and T host synchronisations per batch).
//
// Rows are processed in type-sorted order (`perm`: sorted position -> output row; NULL when the rows already are
// type-contiguous, as in the HGB models), so every 64-row tile has ONE weight matrix: a plain tiled GEMM with the row
// gather folded into the A-tile load and the scatter into the epilogue.  The dense projections are the one place of this
// library where the tensor pipe applies: fp32 operands are split into two TF32 terms and multiplied with
// three mma.sync.m16n8k8 instructions per product (hi*hi + hi*lo + lo*hi, fp32 accumulate), which reproduces fp32 GEMM
// accuracy (~1e-6 relative) -- the parity contract (1e-5) forbids plain TF32.  (tcgen05 + TMEM is the sm_100a-native
// path for large dense GEMMs; at K = 20 ... 4231, 64-row type segments and gathered A rows the legacy warp-level MMA is
// the practical choice, and the projection is not the hot path.)
#include "common.cuh"

namespace regnn {

struct GroupedArgs {
  const float* x[REGNN_MAX_NODE_TYPES];   // feature table of type t: [n_t, k[t]], leading dimension ldx[t]
  const float* w[REGNN_MAX_NODE_TYPES];   // nn.Linear weight of type t: [n_out, k[t]] row-major
  const float* b[REGNN_MAX_NODE_TYPES];   // bias [n_out] or null
  float* dw[REGNN_MAX_NODE_TYPES];        // backward: per-split partial weight gradients [splits, n_out, k[t]]
  float* db[REGNN_MAX_NODE_TYPES];        // backward: per-split partial bias gradients [splits, n_out] (or null)
  int k[REGNN_MAX_NODE_TYPES];
  int64_t ldx[REGNN_MAX_NODE_TYPES];
  int num_types, n_out;
  const int32_t* seg_ptr;     // [T+1]: rows of type t occupy sorted positions [seg_ptr[t], seg_ptr[t+1])
  const int64_t* perm;        // sorted position -> row of out / dout (null: identity)
  const int64_t* local_idx;   // row of out -> row of its type's table (null: sorted position - seg_ptr[t])
  float* out;                 // forward result [M, n_out] (leading dimension ldo); backward: dL/d out
  int64_t ldo;
  int splits;
};

constexpr int BM = 64, BN = 64, BK = 32, PAD = 4;   // block tile; 4 warps as 2 x 2, each 32 x 32 (2 x 4 mma tiles)
constexpr int LDT = BK + PAD;                          // shared-memory row pitch (words): fragment loads conflict-free

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// D(16x8) += A(16x8, row) * B(8x8, col), TF32 operands, fp32 accumulate
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Operand tiles live in shared memory already split into their two TF32 terms (hi = tf32(x), lo = tf32(x - hi)): the
// split costs 3 instructions per element and is done ONCE by the thread that stages the element, not by every warp
// that multiplies with it.
struct Tiles {
  uint32_t ah[BM][LDT], al[BM][LDT], bh[BN][LDT], bl[BN][LDT];
};
__device__ __forceinline__ void put(uint32_t (*hi)[LDT], uint32_t (*lo)[LDT], int r, int c, float x) {
  const uint32_t h = to_tf32(x);
  hi[r][c] = h;
  lo[r][c] = to_tf32(x - __uint_as_float(h));
}

// One BK-deep step of the 64 x 64 block tile (A[m][k], B[n][k], both k-contiguous).
__device__ __forceinline__ void tile_step(const Tiles& s, int wm, int wn, int lane, float (&acc)[2][4][4]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kk = 0; kk < BK; kk += 8) {
    uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = wm * 32 + i * 16 + g;
      ah[i][0] = s.ah[r][kk + t];         al[i][0] = s.al[r][kk + t];
      ah[i][1] = s.ah[r + 8][kk + t];     al[i][1] = s.al[r + 8][kk + t];
      ah[i][2] = s.ah[r][kk + t + 4];     al[i][2] = s.al[r][kk + t + 4];
      ah[i][3] = s.ah[r + 8][kk + t + 4]; al[i][3] = s.al[r + 8][kk + t + 4];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = wn * 32 + j * 8 + g;
      bh[j][0] = s.bh[c][kk + t];     bl[j][0] = s.bl[c][kk + t];
      bh[j][1] = s.bh[c][kk + t + 4]; bl[j][1] = s.bl[c][kk + t + 4];
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {   // small terms first
        mma_tf32(acc[i][j], al[i], bh[j]);
        mma_tf32(acc[i][j], ah[i], bl[j]);
        mma_tf32(acc[i][j], ah[i], bh[j]);
      }
  }
}

constexpr int kFlush = 4;   // k-tiles (of BK) collected in the MMA accumulator before a round-to-nearest flush
__device__ __forceinline__ void flush_acc(float (&acc)[2][4][4], float (&tot)[2][4][4]) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        tot[i][j][q] += acc[i][j][q];
        acc[i][j][q] = 0.f;
      }
}

// Type and row range of forward tile `tile` (tiles are numbered type by type); false when past the last tile.
__device__ __forceinline__ bool locate_tile(const GroupedArgs& a, int tile, int* type, int* row0, int* row1) {
  int first = 0;
  for (int t = 0; t < a.num_types; ++t) {
    const int s0 = a.seg_ptr[t], s1 = a.seg_ptr[t + 1];
    const int nt = (s1 - s0 + BM - 1) / BM;
    if (tile < first + nt) {
      *type = t;
      *row0 = s0 + (tile - first) * BM;
      *row1 = min(s1, *row0 + BM);
      return true;
    }
    first += nt;
  }
  return false;
}

// 4 consecutive floats of a row (k .. k+3) with the tail / alignment handled: a 128-bit load when the row allows it
__device__ __forceinline__ float4 load4(const float* row, int k, int K, bool vec) {
  if (vec && k + 3 < K) return ldg4(row + k);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k < K) v.x = __ldg(row + k);
  if (k + 1 < K) v.y = __ldg(row + k + 1);
  if (k + 2 < K) v.z = __ldg(row + k + 2);
  if (k + 3 < K) v.w = __ldg(row + k + 3);
  return v;
}

constexpr int kLd = (BM * BK / 4) / 128;   // float4 loads per thread and operand tile (= 4)

__global__ void __launch_bounds__(128)
grouped_linear_fwd_kernel(GroupedArgs a) {
  __shared__ __align__(16) Tiles sm;
  __shared__ int64_t src_row[BM], dst_row[BM];
  int type, row0, row1;
  if (!locate_tile(a, blockIdx.x, &type, &row0, &row1)) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int n0 = blockIdx.y * BN;
  const int K = a.k[type];
  const float* X = a.x[type];
  const float* W = a.w[type];
  const int64_t ldx = a.ldx[type];
  const bool vec_x = (ldx & 3) == 0 && ((uintptr_t)X & 15) == 0;
  const bool vec_w = (K & 3) == 0 && ((uintptr_t)W & 15) == 0;
  if (tid < BM) {
    const int p = row0 + tid;
    int64_t o = -1, s = 0;
    if (p < row1) {
      o = a.perm != nullptr ? a.perm[p] : p;
      s = a.local_idx != nullptr ? a.local_idx[o] : p - a.seg_ptr[type];
    }
    dst_row[tid] = o;
    src_row[tid] = s;
  }
  __syncthreads();
  // The tensor pipe adds into its fp32 accumulator with truncation: over thousands of k steps that bias shows (measured:
  // 2.5e-5 of max|out| at K = 4231).  The MMA accumulator therefore only collects kFlush k-tiles (128 values of k) and is
  // then added, round-to-nearest, into a second set of registers.
  float acc[2][4][4] = {}, tot[2][4][4] = {};
  float4 pa[kLd], pb[kLd];
  auto fetch = [&](int k0) {   // global -> registers: 8 threads read the 128 contiguous bytes of one row
#pragma unroll
    for (int i = 0; i < kLd; ++i) {
      const int e = tid + i * 128, r = e >> 3, k = k0 + (e & 7) * 4;
      pa[i] = dst_row[r] >= 0 ? load4(X + src_row[r] * ldx, k, K, vec_x) : make_float4(0.f, 0.f, 0.f, 0.f);
      pb[i] = n0 + r < a.n_out ? load4(W + (size_t)(n0 + r) * K, k, K, vec_w) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stage = [&]() {         // registers -> shared memory, split into the two TF32 terms
#pragma unroll
    for (int i = 0; i < kLd; ++i) {
      const int e = tid + i * 128, r = e >> 3, c = (e & 7) * 4;
      put(sm.ah, sm.al, r, c, pa[i].x); put(sm.ah, sm.al, r, c + 1, pa[i].y);
      put(sm.ah, sm.al, r, c + 2, pa[i].z); put(sm.ah, sm.al, r, c + 3, pa[i].w);
      put(sm.bh, sm.bl, r, c, pb[i].x); put(sm.bh, sm.bl, r, c + 1, pb[i].y);
      put(sm.bh, sm.bl, r, c + 2, pb[i].z); put(sm.bh, sm.bl, r, c + 3, pb[i].w);
    }
  };
  fetch(0);
  int since_flush = 0;
  for (int k0 = 0; k0 < K; k0 += BK) {
    stage();
    __syncthreads();
    if (k0 + BK < K) fetch(k0 + BK);   // the next tile's loads fly behind this tile's MMAs
    tile_step(sm, wm, wn, lane, acc);
    __syncthreads();
    if (++since_flush == kFlush) {
      flush_acc(acc, tot);
      since_flush = 0;
    }
  }
  flush_acc(acc, tot);
  const int g = lane >> 2, t = lane & 3;
  const float* bias = a.b[type];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = wm * 32 + i * 16 + g + half * 8;
      const int64_t o = dst_row[r];
      if (o < 0) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + wn * 32 + j * 8 + 2 * t;
        if (n < a.n_out) a.out[o * a.ldo + n] = tot[i][j][half * 2] + (bias != nullptr ? bias[n] : 0.f);
        if (n + 1 < a.n_out) a.out[o * a.ldo + n + 1] = tot[i][j][half * 2 + 1] + (bias != nullptr ? bias[n + 1] : 0.f);
      }
    }
}

// d W_t[n,k] = sum_{rows of type t} dout[row,n] * x_t[src(row),k]   (and d b_t[n] = sum dout[row,n]):
// block (x: k tile, y: n tile, z: type * splits + split) reduces the rows of its split of the type's segment in order
// (deterministic), BK rows per step, and writes ONE partial tile; the caller adds the `splits` partials in split order.
// The operand tiles are the TRANSPOSES of row chunks: A[n][row] = dout[row][n], B[k][row] = x[row][k].
__global__ void __launch_bounds__(128)
grouped_linear_bwd_w_kernel(GroupedArgs a) {
  __shared__ __align__(16) Tiles sm;
  const int type = blockIdx.z / a.splits, split = blockIdx.z % a.splits;
  const int K = a.k[type];
  const int k0 = blockIdx.x * BN, n0 = blockIdx.y * BM;
  if (k0 >= K) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int s0 = a.seg_ptr[type], s1 = a.seg_ptr[type + 1];
  const int per = (((s1 - s0) + a.splits - 1) / a.splits + BK - 1) / BK * BK;   // rows per split, a multiple of BK
  const int r_begin = s0 + split * per, r_end = min(s1, r_begin + per);
  const float* X = a.x[type];
  const int64_t ldx = a.ldx[type];
  const bool vec_x = (ldx & 3) == 0 && ((uintptr_t)X & 15) == 0;
  const bool vec_d = (a.ldo & 3) == 0 && ((uintptr_t)a.out & 15) == 0;
  float acc[2][4][4] = {}, tot[2][4][4] = {};
  float bsum = 0.f;   // thread n < 64 of the k-tile-0 blocks: column sum of dout
  float4 pa[kLd], pb[kLd];
  auto fetch = [&](int r0) {   // 16 threads read the 64 contiguous n (resp. k) values of one row of the chunk
#pragma unroll
    for (int i = 0; i < kLd; ++i) {
      const int e = tid + i * 128, c = e >> 4, m = (e & 15) * 4;
      const int p = r0 + c;
      pa[i] = pb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < r_end) {
        const int64_t o = a.perm != nullptr ? a.perm[p] : p;
        const int64_t s = a.local_idx != nullptr ? a.local_idx[o] : p - s0;
        pa[i] = load4(a.out + o * a.ldo, n0 + m, a.n_out, vec_d);
        pb[i] = load4(X + s * ldx, k0 + m, K, vec_x);
      }
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int i = 0; i < kLd; ++i) {
      const int e = tid + i * 128, c = e >> 4, m = (e & 15) * 4;
      put(sm.ah, sm.al, m, c, pa[i].x); put(sm.ah, sm.al, m + 1, c, pa[i].y);
      put(sm.ah, sm.al, m + 2, c, pa[i].z); put(sm.ah, sm.al, m + 3, c, pa[i].w);
      put(sm.bh, sm.bl, m, c, pb[i].x); put(sm.bh, sm.bl, m + 1, c, pb[i].y);
      put(sm.bh, sm.bl, m + 2, c, pb[i].z); put(sm.bh, sm.bl, m + 3, c, pb[i].w);
    }
  };
  fetch(r_begin);
  int since_flush = 0;
  for (int r0 = r_begin; r0 < r_end; r0 += BK) {
    stage();
    __syncthreads();
    if (r0 + BK < r_end) fetch(r0 + BK);
    if (blockIdx.x == 0 && tid < BM) {   // d b: column sums of dout, rows in order (hi + lo restores the fp32 value to ~2^-22)
#pragma unroll 8
      for (int c = 0; c < BK; ++c) bsum += __uint_as_float(sm.ah[tid][c]) + __uint_as_float(sm.al[tid][c]);
    }
    tile_step(sm, wm, wn, lane, acc);
    __syncthreads();
    if (++since_flush == kFlush) {
      flush_acc(acc, tot);
      since_flush = 0;
    }
  }
  flush_acc(acc, tot);
  const int g = lane >> 2, t = lane & 3;
  float* dw = a.dw[type] + (size_t)split * a.n_out * K;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int n = n0 + wm * 32 + i * 16 + g + half * 8;
      if (n >= a.n_out) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + wn * 32 + j * 8 + 2 * t;
        if (k < K) dw[(size_t)n * K + k] = tot[i][j][half * 2];
        if (k + 1 < K) dw[(size_t)n * K + k + 1] = tot[i][j][half * 2 + 1];
      }
    }
  if (blockIdx.x == 0 && tid < BM && a.db[type] != nullptr && n0 + tid < a.n_out)
    a.db[type][(size_t)split * a.n_out + n0 + tid] = bsum;
}

static int fill_args(GroupedArgs* a, int num_types, const float* const* x, const int64_t* ldx, const int* k,
                     const float* const* w, const float* const* b, int n_out, const int32_t* seg_ptr, const int64_t* perm,
                     const int64_t* local_idx, const char* who) {
  REGNN_REQUIRE(num_types >= 1 && num_types <= REGNN_MAX_NODE_TYPES, REGNN_ERR_UNSUPPORTED_SHAPE,
                "%s: num_types=%d outside [1,%d]", who, num_types, REGNN_MAX_NODE_TYPES);
  REGNN_REQUIRE(x && ldx && k && w && seg_ptr && n_out >= 1, REGNN_ERR_INVALID_ARG, "%s: null pointer", who);
  *a = GroupedArgs{};
  for (int t = 0; t < num_types; ++t) {
    REGNN_REQUIRE(x[t] && w[t] && k[t] >= 1 && ldx[t] >= k[t], REGNN_ERR_INVALID_ARG, "%s: bad table / weight of type %d", who, t);
    a->x[t] = x[t];
    a->w[t] = w[t];
    a->b[t] = b != nullptr ? b[t] : nullptr;
    a->k[t] = k[t];
    a->ldx[t] = ldx[t];
  }
  a->num_types = num_types;
  a->n_out = n_out;
  a->seg_ptr = seg_ptr;
  a->perm = perm;
  a->local_idx = local_idx;
  return REGNN_OK;
}

}  // namespace regnn

using namespace regnn;

extern "C" int regnn_grouped_linear_fwd(int num_types, const float* const* x, const int64_t* ldx, const int* k,
                                        const float* const* w, const float* const* b, int n_out, int64_t num_rows,
                                        const int32_t* seg_ptr, const int64_t* perm, const int64_t* local_idx, float* out,
                                        int64_t ldo, void* stream) {
  GroupedArgs a;
  int rc = fill_args(&a, num_types, x, ldx, k, w, b, n_out, seg_ptr, perm, local_idx, "grouped_linear_fwd");
  if (rc != REGNN_OK) return rc;
  REGNN_REQUIRE(out && ldo >= n_out && num_rows >= 0, REGNN_ERR_INVALID_ARG, "grouped_linear_fwd: bad output");
  if (num_rows == 0) return REGNN_OK;
  a.out = out;
  a.ldo = ldo;
  const unsigned tiles = (unsigned)((num_rows + BM - 1) / BM + num_types);   // upper bound: blocks past the last tile exit
  grouped_linear_fwd_kernel<<<dim3(tiles, (unsigned)((n_out + BN - 1) / BN)), 128, 0, (cudaStream_t)stream>>>(a);
  return check_launch("regnn_grouped_linear_fwd");
}

extern "C" int regnn_grouped_linear_bwd(int num_types, const float* const* x, const int64_t* ldx, const int* k, int n_out,
                                        int64_t num_rows, const int32_t* seg_ptr, const int64_t* perm,
                                        const int64_t* local_idx, const float* dout, int64_t ldo, int splits,
                                        float* const* dw_partials, float* const* db_partials, void* stream) {
  REGNN_REQUIRE(dw_partials && dout && splits >= 1 && splits <= 64, REGNN_ERR_INVALID_ARG, "grouped_linear_bwd: bad argument");
  GroupedArgs a;
  // the weight pointers are not read by the backward kernel: the partial buffers stand in for the null check
  int rc = fill_args(&a, num_types, x, ldx, k, reinterpret_cast<const float* const*>(dw_partials), nullptr, n_out, seg_ptr,
                     perm, local_idx, "grouped_linear_bwd");
  if (rc != REGNN_OK) return rc;
  if (num_rows == 0) return REGNN_OK;
  int kmax = 0;
  for (int t = 0; t < num_types; ++t) {
    a.dw[t] = dw_partials[t];
    a.db[t] = db_partials != nullptr ? db_partials[t] : nullptr;
    kmax = k[t] > kmax ? k[t] : kmax;
  }
  a.out = const_cast<float*>(dout);
  a.ldo = ldo;
  a.splits = splits;
  grouped_linear_bwd_w_kernel<<<dim3((unsigned)((kmax + BN - 1) / BN), (unsigned)((n_out + BM - 1) / BM),
                                     (unsigned)(num_types * splits)),
                                128, 0, (cudaStream_t)stream>>>(a);
  return check_launch("regnn_grouped_linear_bwd");
}
