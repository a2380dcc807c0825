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

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;   // block tile; 4 warps as 2 x 2, each 32 x 32 (2 x 4 mma tiles)

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
}
// D(16x8) += A(16x8, row) * B(8x8, col), TF32 operands, fp32 accumulate
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// One BK-deep step of the 64 x 64 block tile: As[m][k], Bs[n][k] (both k-contiguous, row pitch BK + PAD).
__device__ __forceinline__ void tile_step(const float (*As)[BK + PAD], const float (*Bs)[BK + PAD], int wm, int wn, int lane,
                                          float (&acc)[2][4][4]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int kk = 0; kk < BK; kk += 8) {
    uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = wm * 32 + i * 16 + g;
      split_tf32(As[r][kk + t], ah[i][0], al[i][0]);
      split_tf32(As[r + 8][kk + t], ah[i][1], al[i][1]);
      split_tf32(As[r][kk + t + 4], ah[i][2], al[i][2]);
      split_tf32(As[r + 8][kk + t + 4], ah[i][3], al[i][3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = wn * 32 + j * 8 + g;
      split_tf32(Bs[c][kk + t], bh[j][0], bl[j][0]);
      split_tf32(Bs[c][kk + t + 4], bh[j][1], bl[j][1]);
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

__global__ void __launch_bounds__(128)
grouped_linear_fwd_kernel(GroupedArgs a) {
  __shared__ __align__(16) float As[BM][BK + PAD];
  __shared__ __align__(16) float Bs[BN][BK + PAD];
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
  float acc[2][4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    // A tile: 64 gathered rows x BK; B tile: 64 weight rows x BK -- 8 scalars per thread each, coalesced along k
#pragma unroll
    for (int i = 0; i < (BM * BK) / 128; ++i) {
      const int e = tid + i * 128;
      const int r = e / BK, c = e % BK;
      const int k = k0 + c;
      As[r][c] = (dst_row[r] >= 0 && k < K) ? __ldg(X + src_row[r] * ldx + k) : 0.f;
      const int n = n0 + r;
      Bs[r][c] = (n < a.n_out && k < K) ? __ldg(W + (size_t)n * K + k) : 0.f;
    }
    __syncthreads();
    tile_step(As, Bs, wm, wn, lane, acc);
    __syncthreads();
  }
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
        if (n < a.n_out) a.out[o * a.ldo + n] = acc[i][j][half * 2] + (bias != nullptr ? bias[n] : 0.f);
        if (n + 1 < a.n_out) a.out[o * a.ldo + n + 1] = acc[i][j][half * 2 + 1] + (bias != nullptr ? bias[n + 1] : 0.f);
      }
    }
}

// d W_t[n,k] = sum_{rows of type t} dout[row,n] * x_t[src(row),k]   (and d b_t[n] = sum dout[row,n]):
// block (x: k tile, y: n tile, z: type * splits + split) reduces the rows of its split of the type's segment in order
// (deterministic), 16 rows per step, and writes ONE partial tile; the caller adds the `splits` partials in split order.
__global__ void __launch_bounds__(128)
grouped_linear_bwd_w_kernel(GroupedArgs a) {
  __shared__ __align__(16) float As[BM][BK + PAD];   // dout^T: [n][row]
  __shared__ __align__(16) float Bs[BN][BK + PAD];   // x^T:    [k][row]
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
  float acc[2][4][4] = {};
  float bsum = 0.f;   // thread n < 64 of the k-tile-0 blocks: column sum of dout
  for (int r0 = r_begin; r0 < r_end; r0 += BK) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / 128; ++i) {
      const int e = tid + i * 128;
      const int c = e / BM, m = e % BM;          // row r0 + c of the segment, element m of the 64-wide n / k tile
      const int p = r0 + c;
      float dv = 0.f, xv = 0.f;
      if (p < r_end) {
        const int64_t o = a.perm != nullptr ? a.perm[p] : p;
        const int64_t s = a.local_idx != nullptr ? a.local_idx[o] : p - s0;
        if (n0 + m < a.n_out) dv = __ldg(a.out + o * a.ldo + n0 + m);
        if (k0 + m < K) xv = __ldg(X + s * ldx + k0 + m);
      }
      As[m][c] = dv;
      Bs[m][c] = xv;
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid < BM) {
#pragma unroll
      for (int c = 0; c < BK; ++c) bsum += As[tid][c];
    }
    tile_step(As, Bs, wm, wn, lane, acc);
    __syncthreads();
  }
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
        if (k < K) dw[(size_t)n * K + k] = acc[i][j][half * 2];
        if (k + 1 < K) dw[(size_t)n * K + k + 1] = acc[i][j][half * 2 + 1];
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
