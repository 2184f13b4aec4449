// UVd preconditioner  Q = (I + U V^T) diag(d)  -- streaming sm_100a kernels.
//
// Replaces the TensorFlow op sequences of update_precond_UVd_math_ (psgd.py:554-617),
// precond_grad_UVd_math (psgd.py:619-627) and IpUVtmatvec (psgd.py:540-544).
//
// Design: every big operand is indexed by parameter row, so each call is a short chain of
// bandwidth-bound sweeps separated by tiny r x r kernels.  What sets the byte count is how many global reductions gate
// the element-wise maps.  All reductions over a = Qh = dh + U p and b = Q^-T v = w - V s1 are expanded through ONE Gram
// table, so only max|nablaD| is left as a second dependency:
//
//   update (default, "uvd_fused" = 1):
//            sweep 1  (reduce)     G = [U V dh w]^T [U V dh w]   (dh = d*h, w = v/d)
//            small 1               p, t, s1, s2 (two r x r LU solves) AND the rank-2 step: a.a, b.b, a.b, a^T X, b^T X
//                                  expanded through G, Frobenius normaliser, coefficient vectors
//            sweep 2  (map)        a, b, nablaD per row; max|nablaD|; U (or V) -= rank-2 update, written back
//            pass 3                d -= mu_d d nablaD
//   update ("uvd_fused" = 0, cross-check): sweep 2 stores a, b, nablaD and reduces a.a, ... directly; small 2; sweep 3
//            writes d and U (or V).
//   apply:   sweep A1 (reduce)  U^T U, U^T(d g), V^T(d g);  small A;  sweep A2 (map)  out = d (y + V t),  y = d g + U p
//   update+apply (psgd_uvd_update_apply): sweep 1 as above; sweep 2 also accumulates [U' V']^T [d g | d nablaD g] over the
//            UPDATED rows (U'^T U' is expanded through G once more); small UA; sweep 3 = d update fused with A2.
//
// Sweeps stage row tiles in shared memory with TMA bulk copies (cp.async.bulk, mbarrier pipeline,
// one producer lane per CTA) so every global access is a full-line coalesced transaction although
// rows are only 4r bytes; consumer warps read rows conflict-free from shared memory.  Cross-CTA
// reductions are two-stage and fixed-order (per-CTA float partials -> one CTA summing in float64),
// so results are deterministic and identical for any grid size.
#include <math.h>

#include "comm.cuh"

namespace psgd {
namespace uvd {

// Pipeline depth knobs (tools/uvd_variants.py builds A/B libraries with other values)
#ifndef PSGD_UVD_SMEM_KB
#define PSGD_UVD_SMEM_KB 160
#endif
#ifndef PSGD_UVD_MAX_STAGES
#define PSGD_UVD_MAX_STAGES 8
#endif
#ifndef PSGD_GRAM_RPL
#define PSGD_GRAM_RPL 0            // rows per lane and pipeline stage in the Gram sweep; 0 = by warps per role
#endif
#ifndef PSGD_MAP_RPL
#define PSGD_MAP_RPL 0             // rows per consumer lane and pipeline stage in the map sweeps; 0 = by stage size
#endif
constexpr int kMapLaneRows = 256;    // consumer lanes of a map sweep = rows it handles per step
constexpr int kMaxStages = PSGD_UVD_MAX_STAGES;
constexpr int kMaxRank = 16;
constexpr size_t kSmemBudget = PSGD_UVD_SMEM_KB * 1024;

enum Mode { kUpdate = 0, kApply = 1, kMatvec = 2 };

// ---------------------------------------------------------------------------------------------
// compile-time plan for the Gram sweep: which entries of  Z^T [Z | X]  (Z = [U V]) are needed and
// how rows of that (upper-trapezoidal) table are split across warp roles.
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr bool gp_needed(int R, int MODE, int i, int j) {
  const int W = 2 * R;
  if (MODE == kUpdate) return j >= i;                                   // full upper triangle of [Z X]^T [Z X]
  if (MODE == kApply) return (j >= W) || (i < R && j < R && j >= i);    // U^T U, U^T x, V^T x
  return (i >= R) && (j >= W);                                          // kMatvec: V^T x only
}
__host__ __device__ constexpr int gp_nx(int MODE) { return MODE == kUpdate ? 2 : 1; }
// table rows: the 2R columns of Z; the update also keeps the X rows (dh.dh, dh.w, w.w) so that every reduction over
// a = dh + U p and b = w - V s1 can be expanded through the table (no second reduction sweep)
__host__ __device__ constexpr int gp_nrows(int R, int MODE) { return MODE == kUpdate ? 2 * R + 2 : 2 * R; }
// Accumulators are kept as float2 pairs (columns 2k, 2k+1) so that one packed FFMA2 (fma.rn.f32x2, sm_100) updates
// two table entries; a pair is live when either of its columns is needed.
__host__ __device__ constexpr bool gp_pair_needed(int R, int MODE, int i, int k) {
  const int E = 2 * R + gp_nx(MODE);
  return gp_needed(R, MODE, i, 2 * k) || (2 * k + 1 < E && gp_needed(R, MODE, i, 2 * k + 1));
}
// registers (floats) held for table row i
__host__ __device__ constexpr int gp_row_count(int R, int MODE, int i) {
  int c = 0;
  const int E2 = (2 * R + gp_nx(MODE) + 1) / 2;
  for (int k = 0; k < E2; ++k) c += gp_pair_needed(R, MODE, i, k) ? 2 : 0;
  return c;
}
__host__ __device__ constexpr int gp_total(int R, int MODE) {
  int c = 0;
  for (int i = 0; i < gp_nrows(R, MODE); ++i) c += gp_row_count(R, MODE, i);
  return c;
}
// Tuning knobs of the Gram sweep (tools/uvd_variants.py builds A/B libraries with other values):
//   PSGD_GRAM_ACC   accumulator floats per warp role (fewer -> more roles, more warps, each re-reading the row)
//   PSGD_GRAM_WPR4  warps per role when the plan has 4 roles
//   PSGD_GRAM_RCP   1: w = v * rcp(d) (2 roundings) instead of the IEEE division in the Gram sweep's x columns
#ifndef PSGD_GRAM_ACC
#define PSGD_GRAM_ACC 104
#endif
#ifndef PSGD_GRAM_WPR4
#define PSGD_GRAM_WPR4 2
#endif
#ifndef PSGD_GRAM_RCP
#define PSGD_GRAM_RCP 0
#endif
constexpr int kAccPerRole = PSGD_GRAM_ACC;
__host__ __device__ constexpr int gp_nroles(int R, int MODE) { return (gp_total(R, MODE) + kAccPerRole - 1) / kAccPerRole; }
// first table row of role `role` (role >= nroles gives the row count)
__host__ __device__ constexpr int gp_begin(int R, int MODE, int role) {
  const int nroles = gp_nroles(R, MODE);
  const int nrows = gp_nrows(R, MODE);
  if (role >= nroles) return nrows;
  const int target = (gp_total(R, MODE) * role) / nroles;
  int c = 0;
  for (int i = 0; i < nrows; ++i) {
    if (c >= target) return i;
    c += gp_row_count(R, MODE, i);
  }
  return nrows;
}

template <int R, int MODE>
struct GramPlan {
  static constexpr int W = 2 * R;
  static constexpr int NX = gp_nx(MODE);
  static constexpr int E = W + NX;
  static constexpr int E2 = (E + 1) / 2;          // float2 column pairs
  static constexpr int NR = gp_nrows(R, MODE);    // table rows
  static constexpr int TABLE = NR * E;            // floats per table ([NR][E], only the needed entries are meaningful)
  static constexpr int NROLES = gp_nroles(R, MODE);
  // warps per role: keep the CTA at <= 12 warps so ptxas may use > 128 registers per thread
  static constexpr int WPR =
      NROLES == 1 ? 8 : (NROLES == 2 ? 4 : (NROLES == 3 ? 3 : (NROLES == 4 ? PSGD_GRAM_WPR4 : (NROLES <= 5 ? 2 : 1))));
  // Rows per lane and pipeline stage.  Bigger tiles mean fewer barrier round trips and TMA operations per byte (measured
  // at r = 10: 96 / 192 / 384-row tiles -> 2.15 / 1.63 / 1.46 ms per sweep); take the largest of 4, 2, 1 that still
  // leaves at least four stages in the shared-memory budget.
  static constexpr int ROW_BYTES = 4 * (W + (MODE == kUpdate ? 3 : (MODE == kApply ? 2 : 1)));
  static constexpr int rpl_auto() {
    for (int rpl = 4; rpl > 1; rpl /= 2)
      if (kSmemBudget / ((size_t)WPR * 32 * rpl * ROW_BYTES) >= 4) return rpl;
    return 1;
  }
  static constexpr int ROWS_PER_LANE = PSGD_GRAM_RPL ? PSGD_GRAM_RPL : rpl_auto();
  static constexpr int TILE = WPR * 32 * ROWS_PER_LANE;          // rows per pipeline stage
  static constexpr int CONSUMER_WARPS = NROLES * WPR;
  static constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
  __host__ __device__ static constexpr bool needed(int i, int j) { return gp_needed(R, MODE, i, j); }
  __host__ __device__ static constexpr bool pair_needed(int i, int k) { return gp_pair_needed(R, MODE, i, k); }
  __host__ __device__ static constexpr int begin(int role) { return gp_begin(R, MODE, role); }
};

// ---------------------------------------------------------------------------------------------
// row loaders (rows are 4R bytes; with a 16-byte aligned base the row start is aligned to
// gcd(4R,16) bytes, which is the vector width used)
// ---------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&out)[R]) {
  if constexpr (R % 4 == 0) {
#pragma unroll
    for (int k = 0; k < R / 4; ++k) {
      float4 t = reinterpret_cast<const float4*>(p)[k];
      out[4 * k] = t.x; out[4 * k + 1] = t.y; out[4 * k + 2] = t.z; out[4 * k + 3] = t.w;
    }
  } else if constexpr (R % 2 == 0) {
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
      float2 t = reinterpret_cast<const float2*>(p)[k];
      out[2 * k] = t.x; out[2 * k + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < R; ++k) out[k] = p[k];
  }
}
template <int R>
__device__ __forceinline__ void store_row(float* __restrict__ p, const float (&in)[R]) {
  if constexpr (R % 4 == 0) {
#pragma unroll
    for (int k = 0; k < R / 4; ++k)
      reinterpret_cast<float4*>(p)[k] = make_float4(in[4 * k], in[4 * k + 1], in[4 * k + 2], in[4 * k + 3]);
  } else if constexpr (R % 2 == 0) {
#pragma unroll
    for (int k = 0; k < R / 2; ++k) reinterpret_cast<float2*>(p)[k] = make_float2(in[2 * k], in[2 * k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < R; ++k) p[k] = in[k];
  }
}
template <int R>
__device__ __forceinline__ float dot_row(const float (&a)[R], const float* __restrict__ b) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < R; ++k) s = fmaf(a[k], b[k], s);
  return s;
}

// ---------------------------------------------------------------------------------------------
// shared-memory tile pipeline: NM matrices [kTile, R] + NV vectors [kTile] per stage
// ---------------------------------------------------------------------------------------------
template <int R, int NM, int NV, int TILE>
struct TileLayout {
  static constexpr int kTile = TILE;
  static constexpr int kMatFloats = kTile * R;
  static constexpr int kStageFloats = NM * kMatFloats + NV * kTile;
  static constexpr uint32_t kStageBytes = kStageFloats * 4;
  static constexpr int kStages =
      (kSmemBudget / kStageBytes) < (size_t)kMaxStages ? (int)(kSmemBudget / kStageBytes) : kMaxStages;
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 2 * kMaxStages * sizeof(uint64_t) + 16;
  __device__ static float* mat(float* stage_base, int m) { return stage_base + m * kMatFloats; }
  __device__ static float* vec(float* stage_base, int v) { return stage_base + NM * kMatFloats + v * kTile; }
};

struct SweepArgs {
  const float* mat[2];
  const float* vec[4];
  int64_t n;          // rows
  int direct;         // 1: bypass the TMA pipeline (debug / cross-check)
  unsigned int* zero_word;   // set to 0 by CTA 0 (the ticket of the mid kernel that consumes this sweep's partials)
};

// Producer lane: stream this CTA's full tiles into the ring.
template <int R, int NM, int NV, int TILE>
__device__ __forceinline__ void producer_loop(const SweepArgs& a, float* smem, uint64_t* full, uint64_t* empty,
                                              int64_t n_full_tiles) {
  using L = TileLayout<R, NM, NV, TILE>;
  constexpr int kTile = TILE;
  int it = 0;
  for (int64_t tile = blockIdx.x; tile < n_full_tiles; tile += gridDim.x, ++it) {
    const int stage = it % L::kStages;
    const uint32_t phase = (it / L::kStages) & 1;
    mbar_wait(&empty[stage], phase ^ 1);
    mbar_arrive_expect_tx(&full[stage], L::kStageBytes);
    float* sb = smem + (size_t)stage * L::kStageFloats;
#pragma unroll
    for (int m = 0; m < NM; ++m)
      bulk_g2s(L::mat(sb, m), a.mat[m] + tile * (int64_t)L::kMatFloats, L::kMatFloats * 4, &full[stage]);
#pragma unroll
    for (int v = 0; v < NV; ++v) bulk_g2s(L::vec(sb, v), a.vec[v] + tile * (int64_t)kTile, kTile * 4, &full[stage]);
  }
}

template <int R, int NM, int NV, int TILE>
__device__ __forceinline__ void pipeline_init(float*& smem, uint64_t*& full, uint64_t*& empty, unsigned char* raw,
                                              int consumer_arrivals) {
  using L = TileLayout<R, NM, NV, TILE>;
  smem = reinterpret_cast<float*>(raw);
  full = reinterpret_cast<uint64_t*>(raw + (size_t)L::kStages * L::kStageBytes);
  empty = full + kMaxStages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < L::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], consumer_arrivals);
    }
    mbar_fence_init();
  }
  __syncthreads();
}

// =============================================================================================
// Sweep 1 / A1 / matvec-reduce:  G += Z^T [Z | X]  restricted to GramPlan::needed
// =============================================================================================
template <int R, int MODE>
struct RowX {  // the X columns of one row
  float x[GramPlan<R, MODE>::NX];
};

// X columns from the per-row vectors: update: (d*h, v/d) ; apply: d*g ; matvec: x
template <int R, int MODE>
__device__ __forceinline__ RowX<R, MODE> make_x(float v0, float v1, float v2) {
  RowX<R, MODE> o;
  if constexpr (MODE == kUpdate) {
    o.x[0] = v0 * v1;   // d*h           psgd.py:569
#if PSGD_GRAM_RCP
    o.x[1] = v2 * __frcp_rn(v0);
#else
    o.x[1] = v2 / v0;   // v/d           psgd.py:576
#endif
  } else if constexpr (MODE == kApply) {
    o.x[0] = v0 * v1;   // d*g           psgd.py:625
  } else {
    o.x[0] = v0;
  }
  return o;
}

template <int R, int MODE, int ROLE>
struct GramRole {
  using P = GramPlan<R, MODE>;
  static constexpr int I0 = P::begin(ROLE);
  static constexpr int I1 = P::begin(ROLE + 1);
  static constexpr int NI = (I1 - I0) > 0 ? (I1 - I0) : 1;
  float2 acc[NI][P::E2];

  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int k = 0; k < P::E2; ++k) acc[i][k] = make_float2(0.f, 0.f);
  }
  __device__ __forceinline__ void row(const float* __restrict__ urow, const float* __restrict__ vrow,
                                      const RowX<R, MODE>& rx) {
    float z[2 * P::E2];
    {
      float u[R], v[R];
      load_row<R>(urow, u);
      load_row<R>(vrow, v);
#pragma unroll
      for (int k = 0; k < R; ++k) { z[k] = u[k]; z[R + k] = v[k]; }
#pragma unroll
      for (int k = 0; k < P::NX; ++k) z[P::W + k] = rx.x[k];
      if constexpr (P::E & 1) z[P::E] = 0.f;
    }
    // acc[i][2k..2k+1] += z[i] * (z[2k], z[2k+1]): one FFMA2 per pair (SASS: FFMA2 with a scalar-broadcast operand)
#pragma unroll
    for (int i = I0; i < I1; ++i)
#pragma unroll
      for (int k = 0; k < P::E2; ++k)
        if (P::pair_needed(i, k))
          acc[i - I0][k] = ffma2(make_float2(z[i], z[i]), make_float2(z[2 * k], z[2 * k + 1]), acc[i - I0][k]);
  }
  // warp butterfly, then lane 0 writes this warp's table rows into `dst` ([NR][E] floats)
  __device__ __forceinline__ void flush(float* dst, int lane) {
#pragma unroll
    for (int i = I0; i < I1; ++i)
#pragma unroll
      for (int j = 0; j < P::E; ++j)
        if (P::needed(i, j)) {
          float s = warp_sum((j & 1) ? acc[i - I0][j / 2].y : acc[i - I0][j / 2].x);
          if (lane == 0) dst[i * P::E + j] = s;
        }
  }
};

template <int R, int MODE, int ROLE>
__device__ __forceinline__ void gram_consumer(const SweepArgs& a, float* smem, uint64_t* full, uint64_t* empty,
                                              int64_t n_full_tiles, int warp_in_role, int lane, float* warp_out) {
  using P = GramPlan<R, MODE>;
  constexpr int NV = (MODE == kUpdate) ? 3 : (MODE == kApply ? 2 : 1);
  using L = TileLayout<R, 2, NV, P::TILE>;
  GramRole<R, MODE, ROLE> role;
  role.zero();
  if (!a.direct) {
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_full_tiles; tile += gridDim.x, ++it) {
      const int stage = it % L::kStages;
      const uint32_t phase = (it / L::kStages) & 1;
      mbar_wait(&full[stage], phase);
      float* sb = smem + (size_t)stage * L::kStageFloats;
      const float* su = L::mat(sb, 0);
      const float* sv = L::mat(sb, 1);
#pragma unroll
      for (int r0 = warp_in_role * 32 + lane; r0 < P::TILE; r0 += P::WPR * 32) {
        const float v0 = L::vec(sb, 0)[r0];
        const float v1 = NV > 1 ? L::vec(sb, NV > 1 ? 1 : 0)[r0] : 0.f;
        const float v2 = NV > 2 ? L::vec(sb, NV > 2 ? 2 : 0)[r0] : 0.f;
        role.row(su + r0 * R, sv + r0 * R, make_x<R, MODE>(v0, v1, v2));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }
  }
  // rows not covered by full tiles (one CTA owns that tail), or every row in direct mode:
  // straight from global memory
  {
    int64_t first, step;
    if (a.direct) {
      first = ((int64_t)blockIdx.x * P::WPR + warp_in_role) * 32 + lane;
      step = (int64_t)gridDim.x * P::WPR * 32;
    } else {
      const bool owner = blockIdx.x == (unsigned)(n_full_tiles % gridDim.x);
      first = owner ? n_full_tiles * P::TILE + warp_in_role * 32 + lane : a.n;
      step = P::WPR * 32;
    }
    for (int64_t r0 = first; r0 < a.n; r0 += step) {
      const float v0 = a.vec[0][r0];
      const float v1 = NV > 1 ? a.vec[NV > 1 ? 1 : 0][r0] : 0.f;
      const float v2 = NV > 2 ? a.vec[NV > 2 ? 2 : 0][r0] : 0.f;
      role.row(a.mat[0] + r0 * R, a.mat[1] + r0 * R, make_x<R, MODE>(v0, v1, v2));
    }
  }
  role.flush(warp_out, lane);
}

template <int R, int MODE, int ROLE>
__device__ __forceinline__ void gram_dispatch(int role, const SweepArgs& a, float* smem, uint64_t* full,
                                              uint64_t* empty, int64_t n_full_tiles, int warp_in_role, int lane,
                                              float* warp_out) {
  using P = GramPlan<R, MODE>;
  if constexpr (ROLE < P::NROLES) {
    if (role == ROLE)
      gram_consumer<R, MODE, ROLE>(a, smem, full, empty, n_full_tiles, warp_in_role, lane, warp_out);
    else
      gram_dispatch<R, MODE, ROLE + 1>(role, a, smem, full, empty, n_full_tiles, warp_in_role, lane, warp_out);
  }
}

// partial: [gridDim.x][NR*E] floats
template <int R, int MODE>
__global__ void __launch_bounds__(GramPlan<R, MODE>::THREADS, 1)
    gram_sweep_kernel(SweepArgs a, float* __restrict__ partial) {
  using P = GramPlan<R, MODE>;
  constexpr int NV = (MODE == kUpdate) ? 3 : (MODE == kApply ? 2 : 1);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ float warp_tab[P::WPR][P::TABLE];      // per (warp-in-role) table, rows owned by its role

  float* smem; uint64_t* full; uint64_t* empty;
  pipeline_init<R, 2, NV, P::TILE>(smem, full, empty, smem_raw, P::CONSUMER_WARPS);
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.zero_word) *a.zero_word = 0u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_full_tiles = a.n / P::TILE;
  if (warp == P::CONSUMER_WARPS) {
    if (lane == 0 && !a.direct) producer_loop<R, 2, NV, P::TILE>(a, smem, full, empty, n_full_tiles);
    __syncwarp();
  } else {
    const int role = warp / P::WPR, wir = warp % P::WPR;
    gram_dispatch<R, MODE, 0>(role, a, smem, full, empty, n_full_tiles, wir, lane, warp_tab[wir]);
  }
  __syncthreads();
  // fixed-order sum over the WPR warps of each role -> CTA partial
  for (int k = threadIdx.x; k < P::TABLE; k += blockDim.x) {
    const int i = k / P::E, j = k % P::E;
    float s = 0.f;
    if (P::needed(i, j)) {
#pragma unroll
      for (int w = 0; w < P::WPR; ++w) s += warp_tab[w][k];
    }
    partial[(size_t)blockIdx.x * P::TABLE + k] = s;
  }
}

// =============================================================================================
// small kernels (one CTA): fixed-order float64 reduction of CTA partials, r x r algebra
// =============================================================================================
// out[k] = sum over CTAs of partial[b][k], in float64 and in an order that depends only on (nblocks, count): each of
// `parts` threads per entry sums a strided subset of the CTAs (independent loads, pipelined), the subsets are then
// combined in index order.  One CTA of 1024 threads; latency bound (a few microseconds).
constexpr int kReduceThreads = 1024;
constexpr int kReduceMaxParts = 32;
__global__ void __launch_bounds__(kReduceThreads) reduce_partials_kernel(const float* __restrict__ partial, int nblocks,
                                                                         int count, double* __restrict__ out) {
  __shared__ double part_sum[kReduceThreads];
  int parts = kReduceThreads / count;
  if (parts > kReduceMaxParts) parts = kReduceMaxParts;
  if (parts < 1) parts = 1;
  for (int k0 = 0; k0 < count; k0 += kReduceThreads / parts) {       // count > 1024 never happens for r <= 16, but stay general
    const int slots = min(count - k0, kReduceThreads / parts);
    const int k = k0 + (int)threadIdx.x % slots;
    const int part = (int)threadIdx.x / slots;
    if (part < parts) {
      double s = 0.0;
#pragma unroll 8
      for (int b = part; b < nblocks; b += parts) s += (double)partial[(size_t)b * count + k];
      part_sum[part * slots + (k - k0)] = s;
    }
    __syncthreads();
    if ((int)threadIdx.x < slots) {
      double s = 0.0;
      for (int q = 0; q < parts; ++q) s += part_sum[q * slots + threadIdx.x];
      out[k0 + threadIdx.x] = s;
    }
    __syncthreads();
  }
}

// LU with partial pivoting (the algorithm class of tf.linalg.solve / LAPACK gesv), n <= kMaxRank, executed by ONE WARP on
// an augmented system [A | b] in shared memory (row stride kLuLd); the solution replaces the last column.
constexpr int kLuLd = kMaxRank + 1;
__device__ __forceinline__ void warp_lu_solve(int n, double (*A)[kLuLd], int lane) {
  for (int k = 0; k < n; ++k) {
    // pivot: first row of maximal |A[i][k]|, i >= k
    double v = (lane >= k && lane < n) ? fabs(A[lane][k]) : -1.0;
    int idx = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (idx != k && lane <= n) { const double t = A[k][lane]; A[k][lane] = A[idx][lane]; A[idx][lane] = t; }
    __syncwarp();
    const double inv = 1.0 / A[k][k];
    const int rows = n - k - 1, cols = n - k;          // columns k+1 .. n (the last one is the right-hand side)
    for (int e = lane; e < rows * cols; e += 32) {
      const int i = k + 1 + e / cols, j = k + 1 + e % cols;
      A[i][j] -= (A[i][k] * inv) * A[k][j];
    }
    __syncwarp();
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = (lane > i && lane < n) ? A[i][lane] * A[lane][n] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) A[i][n] = (A[i][n] - s) / A[i][i];
    __syncwarp();
  }
}

// device-resident r-sized state shared by the sweeps of one call
struct SmallState {
  // after small 1
  float p[kMaxRank];     // V^T (d h)
  float t[kMaxRank];     // U^T Qh
  float s1[kMaxRank];    // solve(IpVtU^T, U^T (v/d))
  float s2[kMaxRank];    // solve(IpVtU,  V^T invQtv)
  double IpVtU[kMaxRank * kMaxRank];
  double UtU[kMaxRank * kMaxRank];
  double VtV[kMaxRank * kMaxRank];
  // after small 2
  float mu_d;
  float mu;              // V branch
  float c1[kMaxRank];    // U branch: mu * atV IpVtU    | V branch: atU
  float c2[kMaxRank];    // U branch: mu * btV IpVtU    | V branch: btU
  float max_nabla;       // atomic max target of sweep 2
  float maxU, maxV;      // balance
  float rho;
  // fused forms: reductions over a = Qh, b = invQtv expanded through the Gram table (uvd_small1_kernel)
  double aa, bb, ab;     // a.a, b.b, a.b
  double Uta[kMaxRank];  // U^T a
  double Utb[kMaxRank];  // U^T b
  unsigned int ticket;   // mid kernel: CTAs done with their slice (zeroed by the sweep that precedes it)
};

// table accessors for G = Z^T [Z | X] stored [W][E] with only j >= i filled
__device__ __forceinline__ double Gsym(const double* G, int E, int i, int j) { return i <= j ? G[i * E + j] : G[j * E + i]; }

// Rank-2 step of U (or V) from the r-sized reductions over a = Qh, b = invQtv (one warp; psgd.py:589-615):
//   normaliser ||(a atX - b btX) X^T||_F evaluated through the r x r Gram X^T X (X = V on the U branch, U on the V
//   branch), mu = step / (norm + tiny), coefficient vectors of the row update.
__device__ __forceinline__ void rank2_coeffs(double aa, double bb, double ab, const double* atX, const double* btX, int r,
                                             int update_U, float step, float tiny, SmallState* st, int lane) {
  const double* XtX = update_U ? st->VtV : st->UtU;
  // ||X atX^T||^2 = atX (X^T X) atX^T  etc.                                  psgd.py:594-596 / :608-610
  double qaa = 0, qbb = 0, qab = 0;
  if (lane < r) {
    double ra = 0, rb = 0;
    for (int j = 0; j < r; ++j) { ra += XtX[lane * r + j] * atX[j]; rb += XtX[lane * r + j] * btX[j]; }
    qaa = atX[lane] * ra; qbb = btX[lane] * rb; qab = atX[lane] * rb;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    qaa += __shfl_xor_sync(0xffffffffu, qaa, o);
    qbb += __shfl_xor_sync(0xffffffffu, qbb, o);
    qab += __shfl_xor_sync(0xffffffffu, qab, o);
  }
  const float norm = sqrtf(fabsf((float)(aa * qaa + bb * qbb - 2.0 * ab * qab)));
  const float mu = step / (norm + tiny);                                      // psgd.py:597 / :611
  if (lane == 0) st->mu = mu;
  if (lane < r) {
    if (update_U) {                                                           // psgd.py:600-601
      double s1 = 0, s2 = 0;
      for (int i = 0; i < r; ++i) { s1 += atX[i] * st->IpVtU[i * r + lane]; s2 += btX[i] * st->IpVtU[i * r + lane]; }
      st->c1[lane] = mu * (float)s1;
      st->c2[lane] = mu * (float)s2;
    } else {
      st->c1[lane] = (float)atX[lane];
      st->c2[lane] = (float)btX[lane];
    }
  }
}

// One warp.  G: reduced table of sweep 1, [2r+2][2r+2] with the upper triangle filled.
// fused != 0: the reductions over a and b that the rank-2 step needs (a.a, b.b, a.b, a^T X, b^T X) are expanded
// through the table (a = dh + U p, b = w - V s1), so the step's coefficients are ready before sweep 2 and the row
// update rides in that sweep (tests/uvd_pipeline_model.py restates this algebra on the host).
__device__ __forceinline__ void small1_device(const double* __restrict__ G, int r, int fused, int update_U, float step,
                                              float tiny, SmallState* __restrict__ st, int lane) {
  __shared__ double A[kMaxRank][kLuLd];
  __shared__ double sp[kMaxRank], ss1[kMaxRank], sat[kMaxRank], sbt[kMaxRank];
  const int W = 2 * r, E = W + 2;
  for (int e = lane; e < r * r; e += 32) {
    const int i = e / r, j = e % r;
    st->UtU[e] = Gsym(G, E, i, j);
    st->VtV[e] = Gsym(G, E, r + i, r + j);
    // VtU[i][j] = sum_k V[k,i] U[k,j] = G[U_j][V_i]                         psgd.py:574-575
    st->IpVtU[e] = G[j * E + (r + i)] + (i == j ? 1.0 : 0.0);
  }
  // p = V^T(dh);  t = U^T Qh = U^T dh + (U^T U) p                            psgd.py:569-570
  if (lane < r) sp[lane] = G[(r + lane) * E + W];
  __syncwarp();
  double tl = 0.0;
  if (lane < r) {
    double s = G[lane * E + W];
    for (int j = 0; j < r; ++j) s += Gsym(G, E, lane, j) * sp[j];
    st->p[lane] = (float)sp[lane];
    st->t[lane] = (float)s;
    tl = s;
  }
  // s1 = solve(IpVtU^T, U^T w)                                               psgd.py:577
  for (int e = lane; e < r * r; e += 32) {
    const int i = e / r, j = e % r;
    A[i][j] = G[i * E + (r + j)] + (i == j ? 1.0 : 0.0);      // IpVtU[j][i]
  }
  if (lane < r) A[lane][r] = G[lane * E + W + 1];
  __syncwarp();
  warp_lu_solve(r, A, lane);
  if (lane < r) { ss1[lane] = A[lane][r]; st->s1[lane] = (float)A[lane][r]; }
  __syncwarp();
  // s2 = solve(IpVtU, V^T invQtv),  V^T invQtv = V^T w - (V^T V) s1          psgd.py:578
  for (int e = lane; e < r * r; e += 32) {
    const int i = e / r, j = e % r;
    A[i][j] = G[j * E + (r + i)] + (i == j ? 1.0 : 0.0);      // IpVtU[i][j]
  }
  if (lane < r) {
    double s = G[(r + lane) * E + W + 1];
    for (int j = 0; j < r; ++j) s -= Gsym(G, E, r + lane, r + j) * ss1[j];
    A[lane][r] = s;
  }
  __syncwarp();
  warp_lu_solve(r, A, lane);
  if (lane < r) st->s2[lane] = (float)A[lane][r];
  if (lane == 0) st->max_nabla = 0.f;
  if (!fused) return;

  // ---- rank-2 step from the table ---------------------------------------------------------------
  double aa = 0, bb = 0, ab = 0;
  if (lane < r) {
    const int i = lane;
    double uup = 0, vvs = 0, uvs = 0;          // (U^T U p)_i, (V^T V s1)_i, (U^T V s1)_i
    for (int j = 0; j < r; ++j) {
      uup += Gsym(G, E, i, j) * sp[j];
      vvs += Gsym(G, E, r + i, r + j) * ss1[j];
      uvs += G[i * E + (r + j)] * ss1[j];
    }
    aa = sp[i] * (2.0 * G[i * E + W] + uup);                                   // 2 p.U^T dh + p^T U^T U p
    bb = ss1[i] * (vvs - 2.0 * G[(r + i) * E + W + 1]);                        // s1^T V^T V s1 - 2 s1.V^T w
    ab = sp[i] * (G[i * E + W + 1] - uvs) - ss1[i] * sp[i];                    // p.U^T w - p^T U^T V s1 - s1.V^T dh
    st->Uta[i] = tl;                                                           // U^T a = U^T Qh
    st->Utb[i] = G[i * E + W + 1] - uvs;                                       // U^T b = U^T w - U^T V s1
    if (update_U) {
      double pv = 0;                           // (p^T U^T V)_i
      for (int j = 0; j < r; ++j) pv += sp[j] * G[j * E + (r + i)];
      sat[i] = sp[i] + pv;                                                     // a^T V               psgd.py:589
      sbt[i] = G[(r + i) * E + W + 1] - vvs;                                   // b^T V               psgd.py:591
    } else {
      sat[i] = tl;                                                             // a^T U = U^T Qh      psgd.py:603
      sbt[i] = G[i * E + W + 1] - uvs;                                         // b^T U               psgd.py:604
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    aa += __shfl_xor_sync(0xffffffffu, aa, o);
    bb += __shfl_xor_sync(0xffffffffu, bb, o);
    ab += __shfl_xor_sync(0xffffffffu, ab, o);
  }
  aa += G[W * E + W];                                                          // dh.dh
  bb += G[(W + 1) * E + W + 1];                                                // w.w
  ab += G[W * E + W + 1];                                                      // dh.w
  if (lane == 0) { st->aa = aa; st->bb = bb; st->ab = ab; }
  __syncwarp();                                // sat/sbt and st->UtU/VtV/IpVtU written above are read across lanes
  rank2_coeffs(aa, bb, ab, sat, sbt, r, update_U, step, tiny, st, lane);
}
__global__ void __launch_bounds__(32) uvd_small1_kernel(const double* __restrict__ G, int r, int fused, int update_U,
                                                        float step, float tiny, SmallState* __restrict__ st) {
  small1_device(G, r, fused, update_U, step, tiny, st, threadIdx.x);
}

// G2 layout: [0]=a.a [1]=b.b [2]=a.b [3..3+r)=a^T X  [3+r..3+2r)=b^T X   (X = V on the U branch, U on the V branch)
__device__ __forceinline__ void small2_device(const double* __restrict__ G2, int r, int update_U, float step, float tiny,
                                              SmallState* __restrict__ st, int lane) {
  if (lane == 0) st->mu_d = step / (st->max_nabla + tiny);                    // psgd.py:582
  rank2_coeffs(G2[0], G2[1], G2[2], G2 + 3, G2 + 3 + r, r, update_U, step, tiny, st, lane);
}
__global__ void __launch_bounds__(32) uvd_small2_kernel(const double* __restrict__ G2, int r, int update_U, float step,
                                                        float tiny, SmallState* __restrict__ st) {
  small2_device(G2, r, update_U, step, tiny, st, threadIdx.x);
}

// After the map sweep of the fused update+apply call.  G3: reduced sums over the UPDATED factors
//   U'^T x0 | U'^T x1 | V'^T x0 | V'^T x1          with x0 = d g and x1 = d nablaD g,
// so that  Z^T (d' g) = Z^T x0 - mu_d Z^T x1  for d' = d - mu_d d nablaD (mu_d is only known once max|nablaD| is, i.e.
// after that sweep).  U'^T U' needs no sums at all: U' = U - a c1^T + b c2^T (U branch; U' = U on the V branch) expands
// through quantities uvd_small1_kernel already has.  Leaves p, t of the apply in st.                 psgd.py:625-626
__device__ __forceinline__ void small_ua_device(const double* __restrict__ G3, int r, int update_U, float step, float tiny,
                                                SmallState* __restrict__ st, int lane) {
  __shared__ double sp[kMaxRank];
  const float mu_d = step / (st->max_nabla + tiny);                           // psgd.py:582
  if (lane == 0) st->mu_d = mu_d;
  const double* Ux0 = G3;
  const double* Ux1 = Ux0 + r;
  const double* Vx0 = Ux1 + r;
  const double* Vx1 = Vx0 + r;
  if (lane < r) sp[lane] = Vx0[lane] - (double)mu_d * Vx1[lane];
  __syncwarp();
  if (lane < r) {
    const int i = lane;
    double s = Ux0[i] - (double)mu_d * Ux1[i];
    const double aa = st->aa, bb = st->bb, ab = st->ab;
    const double c1i = st->c1[i], c2i = st->c2[i], uai = st->Uta[i], ubi = st->Utb[i];
    for (int j = 0; j < r; ++j) {
      double uu = st->UtU[i * r + j];
      if (update_U) {
        const double c1j = st->c1[j], c2j = st->c2[j];
        uu += -uai * c1j - c1i * st->Uta[j] + ubi * c2j + c2i * st->Utb[j] + aa * c1i * c1j - ab * (c1i * c2j + c2i * c1j) +
              bb * c2i * c2j;
      }
      s += uu * sp[j];
    }
    st->p[i] = (float)sp[i];
    st->t[i] = (float)s;
  }
}
__global__ void __launch_bounds__(32) uvd_small_ua_kernel(const double* __restrict__ G3, int r, int update_U, float step,
                                                          float tiny, SmallState* __restrict__ st) {
  small_ua_device(G3, r, update_U, step, tiny, st, threadIdx.x);
}

// apply / matvec: p = V^T x ; t = U^T x + (U^T U) p
__device__ __forceinline__ void small_apply_device(const double* __restrict__ G, int r, SmallState* __restrict__ st,
                                                   int lane) {
  __shared__ double sp[kMaxRank];
  const int W = 2 * r, E = W + 1;
  if (lane < r) sp[lane] = G[(r + lane) * E + W];
  __syncwarp();
  if (lane < r) {
    double s = G[lane * E + W];
    for (int j = 0; j < r; ++j) s += Gsym(G, E, lane, j) * sp[j];
    st->p[lane] = (float)sp[lane];
    st->t[lane] = (float)s;
  }
}
__global__ void __launch_bounds__(32) uvd_small_apply_kernel(const double* __restrict__ G, int r,
                                                             SmallState* __restrict__ st) {
  small_apply_device(G, r, st, threadIdx.x);
}

// =============================================================================================
// "mid" kernel: everything between two sweeps in ONE launch.
//   A. fixed-order float64 reduction of the per-CTA partial records, spread over kMidCtas CTAs (each owns a slice of the
//      entries; within a slice `parts` threads per entry sum strided subsets of the CTAs, combined in index order) --
//      the single 1024-thread CTA of reduce_partials_kernel took 15 us for the 148 x 484 update table;
//   B. the LAST CTA to finish (atomic ticket) runs the cross-rank exchange of comm.cuh in place on the reduced record
//      (sharded runs): push to every peer's slab over NVLink, flag, wait, fixed-rank-order sum;
//   C. the same CTA stages the record in shared memory and one of its warps does the r x r algebra.
// At 8 GPUs a sweep is ~0.2 ms, and the serial chain reduce -> exchange -> algebra (three launches, ~40 us, twice per
// step) was what kept the chunk-sharded step at 90 % scaling efficiency.  Results are bit-identical to the separate
// launches' only in the algebra; the reduction order differs (still fixed: it depends on (nblocks, count) alone).
// =============================================================================================
constexpr int kMidCtas = 32;
constexpr int kMidThreads = 256;
constexpr int kMidMaxCount = (2 * kMaxRank + 2) * (2 * kMaxRank + 2);
enum MidKind { kMidNone = 0, kMidSmall1 = 1, kMidSmallUA = 2, kMidSmallApply = 3, kMidSmall2 = 4 };

struct MidArgs {
  const float* partial;     // [nblocks][count]
  int nblocks, count;
  double* G;                // [count] reduced (and all-reduced) record
  unsigned int* ticket;     // zeroed by the sweep that produced `partial`; reset by the last CTA
  SmallState* st;
  int kind, r, fused, update_U;
  float step, tiny;
  // cross-rank exchange (world == 0: single GPU)
  comm::Peers peers;
  int rank, world;
  float* max_buf;           // n_max non-negative maxima all-reduced together with the sums (e.g. &st->max_nabla)
  int n_max;
  unsigned int* host_err;
  unsigned long long timeout_ns;
};

__global__ void __launch_bounds__(kMidThreads) uvd_mid_kernel(const __grid_constant__ MidArgs a) {
  __shared__ double part_sum[kMidThreads];
  __shared__ double sG[kMidMaxCount];
  __shared__ unsigned int s_last;
  const int tid = threadIdx.x;
  // ---- A: this CTA's slice of the entries ----------------------------------------------------------------------
  const int per = (a.count + gridDim.x - 1) / gridDim.x;
  const int k0 = blockIdx.x * per;
  const int slots = max(0, min(per, a.count - k0));
  if (slots > 0) {
    int parts = kMidThreads / slots;
    if (parts > 32) parts = 32;
    const int part = tid / slots, k = k0 + tid % slots;
    if (part < parts) {
      double s = 0.0;
#pragma unroll 4
      for (int b = part; b < a.nblocks; b += parts) s += (double)__ldcg(&a.partial[(size_t)b * a.count + k]);
      part_sum[part * slots + (k - k0)] = s;
    }
    __syncthreads();
    if (tid < slots) {
      double s = 0.0;
      for (int q = 0; q < parts; ++q) s += part_sum[q * slots + tid];
      a.G[k0 + tid] = s;
    }
  }
  // ---- ticket: the last CTA to get here goes on ----------------------------------------------------------------
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid == 0) *a.ticket = 0u;            // ready for the next mid kernel that shares this ticket
  // ---- B: cross-rank exchange, in place on G --------------------------------------------------------------------
  if (a.world > 0) {
    comm::exchange_device(a.peers, a.rank, a.world, a.G, a.count, a.max_buf, a.n_max, a.host_err, a.timeout_ns);
    __syncthreads();
  }
  // ---- C: r x r algebra from shared memory ------------------------------------------------------------------------
  if (a.kind == kMidNone) return;
  for (int k = tid; k < a.count; k += kMidThreads) sG[k] = __ldcg(&a.G[k]);
  __syncthreads();
  if (tid < 32) {
    if (a.kind == kMidSmall1) small1_device(sG, a.r, a.fused, a.update_U, a.step, a.tiny, a.st, tid);
    else if (a.kind == kMidSmallUA) small_ua_device(sG, a.r, a.update_U, a.step, a.tiny, a.st, tid);
    else if (a.kind == kMidSmallApply) small_apply_device(sG, a.r, a.st, tid);
    else if (a.kind == kMidSmall2) small2_device(sG, a.r, a.update_U, a.step, a.tiny, a.st, tid);
  }
}

// =============================================================================================
// map sweeps: one lane per row, all consumer warps equal
// =============================================================================================
constexpr int kMapConsumerWarps = 8;
constexpr int kMapThreads = (kMapConsumerWarps + 1) * 32;

enum MapKind { kMapUpd2 = 0, kMapUpd3U = 1, kMapUpd3V = 2, kMapApply2 = 3, kMapMatvec2 = 4,
               kMapApplyNorm = 5,    // apply + sum of squares of the preconditioned gradient (UVd.step clip, psgd.py:752)
               kMapApplyParam = 6,   // apply fused with the parameter update (UVd.step without clipping, psgd.py:757-762)
               kMapUpdFU = 7,        // fused update sweep: a, b, nablaD per row AND the rank-2 update of U (psgd.py:569-601)
               kMapUpdFV = 8,        //   ... of V (psgd.py:603-615)
               kMapUpdAppU = 9,      // kMapUpdFU + Gram sums of the updated factors for the apply that follows
               kMapUpdAppV = 10,     // kMapUpdFV + the same
               kMapApplyD = 11 };    // d update (psgd.py:584) fused with the apply's map sweep (psgd.py:625-626)

// NM matrices + NV vectors staged per tile; kStoreMat / kStoreVec: which staged matrix / vector is updated in the
// shared tile and written back with a TMA bulk store (-1: none)
template <int KIND> struct MapTraits;
template <> struct MapTraits<kMapUpd2>   { static constexpr int NM = 2, NV = 3, kStoreMat = -1, kStoreVec = -1; };
template <> struct MapTraits<kMapUpd3U>  { static constexpr int NM = 1, NV = 4, kStoreMat = 0, kStoreVec = 3; };
template <> struct MapTraits<kMapUpd3V>  { static constexpr int NM = 1, NV = 4, kStoreMat = 0, kStoreVec = 3; };
template <> struct MapTraits<kMapApply2> { static constexpr int NM = 2, NV = 2, kStoreMat = -1, kStoreVec = -1; };
template <> struct MapTraits<kMapMatvec2>{ static constexpr int NM = 1, NV = 1, kStoreMat = -1, kStoreVec = -1; };
template <> struct MapTraits<kMapApplyNorm>  { static constexpr int NM = 2, NV = 2, kStoreMat = -1, kStoreVec = -1; };
template <> struct MapTraits<kMapApplyParam> { static constexpr int NM = 2, NV = 4, kStoreMat = -1, kStoreVec = -1; };
template <> struct MapTraits<kMapUpdFU>   { static constexpr int NM = 2, NV = 3, kStoreMat = 0, kStoreVec = -1; };
template <> struct MapTraits<kMapUpdFV>   { static constexpr int NM = 2, NV = 3, kStoreMat = 1, kStoreVec = -1; };
template <> struct MapTraits<kMapUpdAppU> { static constexpr int NM = 2, NV = 4, kStoreMat = 0, kStoreVec = -1; };
template <> struct MapTraits<kMapUpdAppV> { static constexpr int NM = 2, NV = 4, kStoreMat = 1, kStoreVec = -1; };
template <> struct MapTraits<kMapApplyD>  { static constexpr int NM = 2, NV = 3, kStoreMat = -1, kStoreVec = -1; };

// Rows per consumer lane and pipeline stage of a map sweep: two when at least three stages of 512 rows fit the
// shared-memory budget, else one (measured at r = 10: 512-row tiles run the fused sweeps 2-4 % faster than 256-row ones).
template <int R, int KIND>
struct MapPlan {
  using T = MapTraits<KIND>;
  static constexpr int ROW_BYTES = 4 * (T::NM * R + T::NV);
  static constexpr int RPL = PSGD_MAP_RPL ? PSGD_MAP_RPL : ((kSmemBudget / ((size_t)2 * kMapLaneRows * ROW_BYTES) >= 3) ? 2 : 1);
  static constexpr int TILE = kMapLaneRows * RPL;
};

constexpr bool map_is_fused_update(int KIND) {
  return KIND == kMapUpdFU || KIND == kMapUpdFV || KIND == kMapUpdAppU || KIND == kMapUpdAppV;
}
constexpr bool map_is_updapp(int KIND) { return KIND == kMapUpdAppU || KIND == kMapUpdAppV; }
// floats per CTA partial record of the update+apply sweep: U'^T x0 | U'^T x1 | V'^T x0 | V'^T x1
constexpr int updapp_count(int R) { return 4 * R; }

struct MapOut {
  float* o0; float* o1; float* o2;   // per-row outputs (coalesced direct stores)
  float* mat_out;                    // store kinds: updated matrix
  float* vec_out;                    // store kinds: updated d
  float* partial;                    // Upd2: [grid][3+2r] sums; ApplyNorm: [grid] sums of squares; UpdApp: [grid][updapp_count]
  SmallState* st;
  float lr;                          // ApplyParam: learning rate
  int has_v;                         // ApplyParam: vec[3] holds the finite-difference perturbation to remove
};

// per-lane accumulators of sweep 2
template <int R>
struct Upd2Acc {
  float aa = 0.f, bb = 0.f, ab = 0.f, mx = 0.f;
  float atX[R], btX[R];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int k = 0; k < R; ++k) { atX[k] = 0.f; btX[k] = 0.f; }
  }
};

// per-lane accumulators of the update+apply sweep: [U' V']^T [x0 x1] over the rows this lane owns
template <int R, bool ON>
struct UpdAppAcc {
  float ux0[ON ? R : 1], ux1[ON ? R : 1], vx0[ON ? R : 1], vx1[ON ? R : 1];
  __device__ __forceinline__ void zero() {
    if constexpr (ON) {
#pragma unroll
      for (int k = 0; k < R; ++k) { ux0[k] = 0.f; ux1[k] = 0.f; vx0[k] = 0.f; vx1[k] = 0.f; }
    }
  }
  __device__ __forceinline__ void add(const float (&u)[R], const float (&v)[R], float x0, float x1) {
    if constexpr (ON) {
#pragma unroll
      for (int i = 0; i < R; ++i) {
        ux0[i] = fmaf(u[i], x0, ux0[i]); ux1[i] = fmaf(u[i], x1, ux1[i]);
        vx0[i] = fmaf(v[i], x0, vx0[i]); vx1[i] = fmaf(v[i], x1, vx1[i]);
      }
    }
  }
};

template <int R, int KIND>
struct MapBody {
  // r-sized constants in registers
  float k0[R], k1[R], k2[R], k3[R], k4[R], k5[R];
  float mu_d, mu;
  Upd2Acc<R> acc;
  UpdAppAcc<R, map_is_updapp(KIND)> gacc;
  int update_U;

  __device__ __forceinline__ void init(const SmallState* st, int upd_u) {
    update_U = upd_u;
    if constexpr (KIND == kMapUpd2) {
#pragma unroll
      for (int k = 0; k < R; ++k) { k0[k] = st->p[k]; k1[k] = st->t[k]; k2[k] = st->s1[k]; k3[k] = st->s2[k]; }
      acc.zero();
    } else if constexpr (map_is_fused_update(KIND)) {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        k0[k] = st->p[k]; k1[k] = st->t[k]; k2[k] = st->s1[k]; k3[k] = st->s2[k]; k4[k] = st->c1[k]; k5[k] = st->c2[k];
      }
      mu = st->mu;
      acc.mx = 0.f;
      gacc.zero();
    } else if constexpr (KIND == kMapUpd3U || KIND == kMapUpd3V) {
#pragma unroll
      for (int k = 0; k < R; ++k) { k0[k] = st->c1[k]; k1[k] = st->c2[k]; }
      mu_d = st->mu_d; mu = st->mu;
    } else if constexpr (KIND == kMapApplyNorm) {
#pragma unroll
      for (int k = 0; k < R; ++k) { k0[k] = st->p[k]; k1[k] = st->t[k]; }
      acc.aa = 0.f;
    } else {
#pragma unroll
      for (int k = 0; k < R; ++k) { k0[k] = st->p[k]; k1[k] = st->t[k]; }
      if constexpr (KIND == kMapApplyD) mu_d = st->mu_d;
    }
  }

  // m0/m1: row pointers (shared or global); vec values v0..v3; `row` global row index;
  // mw: where to write the updated matrix row (shared tile or global); vecw: likewise for the updated d
  __device__ __forceinline__ void row(const float* m0, const float* m1, float v0, float v1, float v2, float v3,
                                      int64_t row, const MapOut& o, float* mw, float* vecw) {
    if constexpr (KIND == kMapUpd2) {
      // v0=d v1=h v2=v                                                       psgd.py:569-581
      float u[R], vv[R];
      load_row<R>(m0, u);
      load_row<R>(m1, vv);
      const float dh = v0 * v1;
      const float Qh = dh + dot_row<R>(u, k0);
      const float Ph = v0 * (Qh + dot_row<R>(vv, k1));
      const float w = v2 / v0;
      const float b = w - dot_row<R>(vv, k2);
      const float invPv = (b - dot_row<R>(u, k3)) / v0;
      const float nd = Ph * v1 - v2 * invPv;
      o.o0[row] = Qh; o.o1[row] = b; o.o2[row] = nd;
      acc.mx = fmaxf(acc.mx, fabsf(nd));
      acc.aa = fmaf(Qh, Qh, acc.aa); acc.bb = fmaf(b, b, acc.bb); acc.ab = fmaf(Qh, b, acc.ab);
      if (update_U) {
#pragma unroll
        for (int k = 0; k < R; ++k) { acc.atX[k] = fmaf(Qh, vv[k], acc.atX[k]); acc.btX[k] = fmaf(b, vv[k], acc.btX[k]); }
      } else {
#pragma unroll
        for (int k = 0; k < R; ++k) { acc.atX[k] = fmaf(Qh, u[k], acc.atX[k]); acc.btX[k] = fmaf(b, u[k], acc.btX[k]); }
      }
    } else if constexpr (map_is_fused_update(KIND)) {
      // v0=d v1=h v2=v (v3=g): the arithmetic of kMapUpd2 and kMapUpd3U/V on one pass     psgd.py:569-581, :600-601, :614-615
      float u[R], vv[R];
      load_row<R>(m0, u);
      load_row<R>(m1, vv);
      const float dh = v0 * v1;
      const float Qh = dh + dot_row<R>(u, k0);
      const float Ph = v0 * (Qh + dot_row<R>(vv, k1));
      const float w = v2 / v0;
      const float b = w - dot_row<R>(vv, k2);
      const float invPv = (b - dot_row<R>(u, k3)) / v0;
      const float nd = Ph * v1 - v2 * invPv;
      o.o2[row] = nd;
      acc.mx = fmaxf(acc.mx, fabsf(nd));
      if constexpr (KIND == kMapUpdFU || KIND == kMapUpdAppU) {
#pragma unroll
        for (int k = 0; k < R; ++k) u[k] = u[k] - (Qh * k4[k] - b * k5[k]);
        store_row<R>(mw, u);
      } else {
        const float sa = Qh + dot_row<R>(vv, k4);
        const float sb = b + dot_row<R>(vv, k5);
#pragma unroll
        for (int k = 0; k < R; ++k) vv[k] = vv[k] - mu * (sa * k4[k] - sb * k5[k]);
        store_row<R>(mw, vv);
      }
      if constexpr (map_is_updapp(KIND)) {
        const float x0 = v0 * v3;          // d g
        gacc.add(u, vv, x0, x0 * nd);      // sums over the UPDATED rows; d' g = x0 - mu_d x1
      }
    } else if constexpr (KIND == kMapUpd3U) {
      // v0=a v1=b v2=nablaD v3=d ; U -= mu (a c1 - b c2) (c pre-scaled)      psgd.py:584, :600-601
      float u[R];
      load_row<R>(m0, u);
#pragma unroll
      for (int k = 0; k < R; ++k) u[k] = u[k] - (v0 * k0[k] - v1 * k1[k]);
      store_row<R>(mw, u);
      *vecw = v3 - (mu_d * v3) * v2;
    } else if constexpr (KIND == kMapUpd3V) {
      // V -= mu ((a + V atU^T) atU - (b + V btU^T) btU)                       psgd.py:584, :614-615
      float vv[R];
      load_row<R>(m0, vv);
      const float sa = v0 + dot_row<R>(vv, k0);
      const float sb = v1 + dot_row<R>(vv, k1);
#pragma unroll
      for (int k = 0; k < R; ++k) vv[k] = vv[k] - mu * (sa * k0[k] - sb * k1[k]);
      store_row<R>(mw, vv);
      *vecw = v3 - (mu_d * v3) * v2;
    } else if constexpr (KIND == kMapApplyD) {
      // v0=d v1=nablaD v2=g: d' = d - mu_d d nablaD, then the apply on d'    psgd.py:584, :625-626
      float u[R], vv[R];
      load_row<R>(m0, u);
      load_row<R>(m1, vv);
      const float dn = v0 - (mu_d * v0) * v1;
      const float dg = dn * v2;
      const float y = dg + dot_row<R>(u, k0);
      o.o0[row] = dn * (y + dot_row<R>(vv, k1));
      o.o1[row] = dn;
    } else if constexpr (KIND == kMapApply2 || KIND == kMapApplyNorm || KIND == kMapApplyParam) {
      // v0=d v1=g                                                            psgd.py:625-626
      float u[R], vv[R];
      load_row<R>(m0, u);
      load_row<R>(m1, vv);
      const float dg = v0 * v1;
      const float y = dg + dot_row<R>(u, k0);
      const float pre = v0 * (y + dot_row<R>(vv, k1));
      if constexpr (KIND == kMapApplyParam) {
        // v2=param v3=perturbation: param -= lr*pre (+ v)                     psgd.py:757-762
        float upd = o.lr * pre;
        if (o.has_v) upd = upd + v3;
        o.o0[row] = v2 - upd;
        if (o.o1) o.o1[row] = pre;
      } else {
        o.o0[row] = pre;
        if constexpr (KIND == kMapApplyNorm) acc.aa = fmaf(pre, pre, acc.aa);     // psgd.py:752
      }
    } else {
      // matvec: out = x + U_row . p   (m0 = U)                               psgd.py:544
      float u[R];
      load_row<R>(m0, u);
      o.o0[row] = v0 + dot_row<R>(u, k0);
    }
  }
};

template <int R, int KIND>
__global__ void __launch_bounds__(kMapThreads, 1) map_sweep_kernel(SweepArgs a, MapOut o, int update_U) {
  using T = MapTraits<KIND>;
  constexpr int kMapTile = MapPlan<R, KIND>::TILE;
  constexpr int kMapRowsPerLane = MapPlan<R, KIND>::RPL;
  using L = TileLayout<R, T::NM, T::NV, kMapTile>;
  static_assert(kMapLaneRows == kMapConsumerWarps * 32, "one consumer lane per row of a step");
  constexpr bool kStore = T::kStoreMat >= 0;
  constexpr int SM = kStore ? T::kStoreMat : 0;                 // staged matrix that is updated in place
  constexpr int SV = T::kStoreVec >= 0 ? T::kStoreVec : 0;
  constexpr int kRedCols = 4 * kMaxRank;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ float red[kMapConsumerWarps][kRedCols];
  __shared__ float red_mx[kMapConsumerWarps];

  float* smem; uint64_t* full; uint64_t* empty;
  // store kinds: one elected lane releases the stage after its bulk store has drained the tile
  pipeline_init<R, T::NM, T::NV, kMapTile>(smem, full, empty, smem_raw, kStore ? 1 : kMapConsumerWarps);
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.zero_word) *a.zero_word = 0u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_full_tiles = a.n / kMapTile;

  if (warp == kMapConsumerWarps) {
    if (lane == 0 && !a.direct) producer_loop<R, T::NM, T::NV, kMapTile>(a, smem, full, empty, n_full_tiles);
    __syncwarp();
  } else {
    MapBody<R, KIND> body;
    body.init(o.st, update_U);
    const int ct = warp * 32 + lane;                   // consumer thread id == row within tile
    if (!a.direct) {
      int it = 0;
      for (int64_t tile = blockIdx.x; tile < n_full_tiles; tile += gridDim.x, ++it) {
        const int stage = it % L::kStages;
        const uint32_t phase = (it / L::kStages) & 1;
        mbar_wait(&full[stage], phase);
        float* sb = smem + (size_t)stage * L::kStageFloats;
#pragma unroll
        for (int rr = 0; rr < kMapRowsPerLane; ++rr) {
          const int r0 = ct + rr * (kMapConsumerWarps * 32);
          const float* m0 = L::mat(sb, 0);
          const float* m1 = L::mat(sb, T::NM > 1 ? 1 : 0);
          const float v0 = L::vec(sb, 0)[r0];
          const float v1 = T::NV > 1 ? L::vec(sb, T::NV > 1 ? 1 : 0)[r0] : 0.f;
          const float v2 = T::NV > 2 ? L::vec(sb, T::NV > 2 ? 2 : 0)[r0] : 0.f;
          const float v3 = T::NV > 3 ? L::vec(sb, T::NV > 3 ? 3 : 0)[r0] : 0.f;
          body.row(m0 + r0 * R, m1 + r0 * R, v0, v1, v2, v3, tile * kMapTile + r0, o, L::mat(sb, SM) + r0 * R,
                   &L::vec(sb, SV)[r0]);
        }
        if constexpr (kStore) {
          // rows were updated in place in the shared tile: publish to the async proxy, then one lane
          // stores the tile with TMA and frees the stage once the engine has read it
          fence_proxy_async_smem();
          named_bar_sync(1, kMapConsumerWarps * 32);
          if (ct == 0) {
            bulk_s2g(o.mat_out + tile * (int64_t)L::kMatFloats, L::mat(sb, SM), L::kMatFloats * 4);
            if constexpr (T::kStoreVec >= 0) bulk_s2g(o.vec_out + tile * (int64_t)kMapTile, L::vec(sb, SV), kMapTile * 4);
            bulk_commit();
            bulk_wait_read0();
            mbar_arrive(&empty[stage]);
          }
        } else {
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);
        }
      }
      if constexpr (kStore) {
        if (ct == 0) bulk_wait0();
      }
    }
    // tail rows (one CTA) / every row in direct mode: straight from and to global memory
    {
      int64_t first, step;
      if (a.direct) {
        first = (int64_t)blockIdx.x * (kMapConsumerWarps * 32) + ct;
        step = (int64_t)gridDim.x * (kMapConsumerWarps * 32);
      } else {
        first = (blockIdx.x == (unsigned)(n_full_tiles % gridDim.x)) ? n_full_tiles * kMapTile + ct : a.n;
        step = kMapConsumerWarps * 32;
      }
      for (int64_t r0 = first; r0 < a.n; r0 += step) {
        const float* m0 = a.mat[0] + r0 * R;
        const float* m1 = a.mat[T::NM > 1 ? 1 : 0] + r0 * R;
        const float v0 = a.vec[0][r0];
        const float v1 = T::NV > 1 ? a.vec[T::NV > 1 ? 1 : 0][r0] : 0.f;
        const float v2 = T::NV > 2 ? a.vec[T::NV > 2 ? 2 : 0][r0] : 0.f;
        const float v3 = T::NV > 3 ? a.vec[T::NV > 3 ? 3 : 0][r0] : 0.f;
        body.row(m0, m1, v0, v1, v2, v3, r0, o, kStore ? o.mat_out + r0 * R : nullptr,
                 T::kStoreVec >= 0 ? o.vec_out + r0 : nullptr);
      }
    }
    if constexpr (KIND == kMapApplyNorm) {
      const float s0 = warp_sum(body.acc.aa);
      if (lane == 0) red[warp][0] = s0;
    }
    if constexpr (KIND == kMapUpd2) {
      // warp butterflies -> per-warp slots -> fixed-order CTA partial
      float mx = warp_max(body.acc.mx);
      float s0 = warp_sum(body.acc.aa), s1 = warp_sum(body.acc.bb), s2 = warp_sum(body.acc.ab);
      if (lane == 0) { red[warp][0] = s0; red[warp][1] = s1; red[warp][2] = s2; red[warp][3] = mx; }
#pragma unroll
      for (int k = 0; k < R; ++k) {
        float ta = warp_sum(body.acc.atX[k]), tb = warp_sum(body.acc.btX[k]);
        if (lane == 0) { red[warp][4 + k] = ta; red[warp][4 + R + k] = tb; }
      }
    }
    if constexpr (map_is_fused_update(KIND)) {
      const float mx = warp_max(body.acc.mx);
      if (lane == 0) red_mx[warp] = mx;
    }
    if constexpr (map_is_updapp(KIND)) {
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const float t0 = warp_sum(body.gacc.ux0[k]), t1 = warp_sum(body.gacc.ux1[k]);
        const float t2 = warp_sum(body.gacc.vx0[k]), t3 = warp_sum(body.gacc.vx1[k]);
        if (lane == 0) { red[warp][k] = t0; red[warp][R + k] = t1; red[warp][2 * R + k] = t2; red[warp][3 * R + k] = t3; }
      }
    }
  }
  if constexpr (KIND == kMapApplyNorm) {
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kMapConsumerWarps; ++w) s += red[w][0];
      o.partial[blockIdx.x] = s;
    }
  }
  if constexpr (KIND == kMapUpd2) {
    __syncthreads();
    const int cnt = 3 + 2 * R;
    if (threadIdx.x < cnt) {
      const int src = threadIdx.x < 3 ? threadIdx.x : threadIdx.x + 1;
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kMapConsumerWarps; ++w) s += red[w][src];
      o.partial[(size_t)blockIdx.x * cnt + threadIdx.x] = s;
    }
    if (threadIdx.x == 0) {
      float mx = 0.f;
#pragma unroll
      for (int w = 0; w < kMapConsumerWarps; ++w) mx = fmaxf(mx, red[w][3]);
      atomic_max_nonneg(&o.st->max_nabla, mx);
    }
  }
  if constexpr (map_is_fused_update(KIND)) {
    __syncthreads();
    if constexpr (map_is_updapp(KIND)) {
      constexpr int cnt = updapp_count(R);
      for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kMapConsumerWarps; ++w) s += red[w][k];
        o.partial[(size_t)blockIdx.x * cnt + k] = s;
      }
    }
    if (threadIdx.x == 0) {
      float mx = 0.f;
#pragma unroll
      for (int w = 0; w < kMapConsumerWarps; ++w) mx = fmaxf(mx, red_mx[w]);
      atomic_max_nonneg(&o.st->max_nabla, mx);
    }
  }
}

// d -= (mu_d d) nablaD with mu_d = step / (max|nablaD| + tiny)                 psgd.py:582-584
// (last pass of the fused update: max|nablaD| gates it, everything else already rode in the map sweep)
__global__ void __launch_bounds__(256) d_update_kernel(float* __restrict__ d, const float* __restrict__ nd, int64_t n,
                                                       float step, float tiny, const SmallState* __restrict__ st) {
  const float mu_d = step / (st->max_nabla + tiny);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n4 = n / 4;
  float4* d4 = reinterpret_cast<float4*>(d);
  const float4* g4 = reinterpret_cast<const float4*>(nd);
  int64_t i = tid;
  for (; i + stride < n4; i += 2 * stride) {          // two independent 128-bit loads per array in flight
    float4 x0 = d4[i], x1 = d4[i + stride];
    const float4 y0 = g4[i], y1 = g4[i + stride];
    x0.x -= (mu_d * x0.x) * y0.x; x0.y -= (mu_d * x0.y) * y0.y; x0.z -= (mu_d * x0.z) * y0.z; x0.w -= (mu_d * x0.w) * y0.w;
    x1.x -= (mu_d * x1.x) * y1.x; x1.y -= (mu_d * x1.y) * y1.y; x1.z -= (mu_d * x1.z) * y1.z; x1.w -= (mu_d * x1.w) * y1.w;
    d4[i] = x0; d4[i + stride] = x1;
  }
  for (; i < n4; i += stride) {
    float4 x0 = d4[i];
    const float4 y0 = g4[i];
    x0.x -= (mu_d * x0.x) * y0.x; x0.y -= (mu_d * x0.y) * y0.y; x0.z -= (mu_d * x0.z) * y0.z; x0.w -= (mu_d * x0.w) * y0.w;
    d4[i] = x0;
  }
  for (int64_t k = n4 * 4 + tid; k < n; k += stride) d[k] -= (mu_d * d[k]) * nd[k];
}

// =============================================================================================
// balance (psgd.py:562-567): rho = sqrt(max|U| / max|V|); U /= rho; V *= rho
// =============================================================================================
__global__ void maxabs2_kernel(const float* __restrict__ A, const float* __restrict__ B, int64_t count,
                               SmallState* st) {
  float ma = 0.f, mb = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = count / 4;
  const float4* A4 = reinterpret_cast<const float4*>(A);
  const float4* B4 = reinterpret_cast<const float4*>(B);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 x = A4[i], y = B4[i];
    ma = fmaxf(ma, fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))));
    mb = fmaxf(mb, fmaxf(fmaxf(fabsf(y.x), fabsf(y.y)), fmaxf(fabsf(y.z), fabsf(y.w))));
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    ma = fmaxf(ma, fabsf(A[i])); mb = fmaxf(mb, fabsf(B[i]));
  }
  ma = warp_max(ma); mb = warp_max(mb);
  if ((threadIdx.x & 31) == 0) { atomic_max_nonneg(&st->maxU, ma); atomic_max_nonneg(&st->maxV, mb); }
}
__global__ void balance_rho_kernel(SmallState* st) { st->rho = sqrtf(st->maxU / st->maxV); }
__global__ void balance_scale_kernel(float* __restrict__ U, float* __restrict__ V, int64_t count,
                                     const SmallState* __restrict__ st) {
  const float rho = st->rho;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = count / 4;
  float4* U4 = reinterpret_cast<float4*>(U);
  float4* V4 = reinterpret_cast<float4*>(V);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 x = U4[i], y = V4[i];
    x.x /= rho; x.y /= rho; x.z /= rho; x.w /= rho;
    y.x *= rho; y.y *= rho; y.z *= rho; y.w *= rho;
    U4[i] = x; V4[i] = y;
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    U[i] = U[i] / rho; V[i] = rho * V[i];
  }
}
__global__ void zero_small_kernel(SmallState* st) { st->maxU = 0.f; st->maxV = 0.f; st->max_nabla = 0.f; }

// =============================================================================================
// host orchestration
// =============================================================================================
static int grid_for(const psgd_ctx* ctx, int64_t n) {
  int64_t tiles = (n + kMapLaneRows - 1) / kMapLaneRows;
  int64_t g = ctx->num_sms;
  if (tiles < g) g = tiles > 0 ? tiles : 1;
  return (int)g;
}

template <int R, int MODE>
static int launch_gram(psgd_ctx* ctx, const SweepArgs& a, float* partial, int grid) {
  ProfScope prof(ctx, MODE == kUpdate ? PSGD_K_UVD_GRAM_UPDATE : PSGD_K_UVD_GRAM_APPLY);
  using P = GramPlan<R, MODE>;
  constexpr int NV = (MODE == kUpdate) ? 3 : (MODE == kApply ? 2 : 1);
  using L = TileLayout<R, 2, NV, P::TILE>;
  auto kern = gram_sweep_kernel<R, MODE>;
  kern<<<grid, P::THREADS, L::kSmemBytes, ctx->stream>>>(a, partial);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

template <int R, int KIND>
static int launch_map(psgd_ctx* ctx, const SweepArgs& a, const MapOut& o, int update_U, int grid) {
  ProfScope prof(ctx, KIND == kMapUpd2 ? PSGD_K_UVD_MAP_UPDATE2
                      : (KIND == kMapUpd3U || KIND == kMapUpd3V) ? PSGD_K_UVD_MAP_UPDATE3
                      : (KIND == kMapUpdFU || KIND == kMapUpdFV) ? PSGD_K_UVD_MAP_FUSED
                      : map_is_updapp(KIND) ? PSGD_K_UVD_MAP_UPDAPP
                      : KIND == kMapApplyD ? PSGD_K_UVD_MAP_APPLY_D : PSGD_K_UVD_MAP_APPLY);
  using T = MapTraits<KIND>;
  using L = TileLayout<R, T::NM, T::NV, MapPlan<R, KIND>::TILE>;
  auto kern = map_sweep_kernel<R, KIND>;
  kern<<<grid, kMapThreads, L::kSmemBytes, ctx->stream>>>(a, o, update_U);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

// Opt every sweep kernel of rank R into its dynamic shared-memory size once, up front, so that no attribute call
// happens later inside a CUDA-graph stream capture.
template <int R, int KIND>
static cudaError_t map_attr() {
  using T = MapTraits<KIND>;
  using L = TileLayout<R, T::NM, T::NV, MapPlan<R, KIND>::TILE>;
  return cudaFuncSetAttribute(map_sweep_kernel<R, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kSmemBytes);
}
template <int R, int MODE>
static cudaError_t gram_attr() {
  using P = GramPlan<R, MODE>;
  constexpr int NV = (MODE == kUpdate) ? 3 : (MODE == kApply ? 2 : 1);
  using L = TileLayout<R, 2, NV, P::TILE>;
  return cudaFuncSetAttribute(gram_sweep_kernel<R, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kSmemBytes);
}
template <int R>
static int ensure_attrs(const psgd_ctx* ctx) {
  static DeviceOnce done;                  // per rank instantiation, per device
  if (done.done(ctx->device)) return PSGD_OK;
  PSGD_CUDA_CHECK((gram_attr<R, kUpdate>()));
  PSGD_CUDA_CHECK((gram_attr<R, kApply>()));
  PSGD_CUDA_CHECK((gram_attr<R, kMatvec>()));
  PSGD_CUDA_CHECK((map_attr<R, kMapUpd2>()));
  PSGD_CUDA_CHECK((map_attr<R, kMapUpd3U>()));
  PSGD_CUDA_CHECK((map_attr<R, kMapUpd3V>()));
  PSGD_CUDA_CHECK((map_attr<R, kMapApply2>()));
  PSGD_CUDA_CHECK((map_attr<R, kMapMatvec2>()));
  PSGD_CUDA_CHECK((map_attr<R, kMapApplyNorm>()));
  PSGD_CUDA_CHECK((map_attr<R, kMapApplyParam>()));
  PSGD_CUDA_CHECK((map_attr<R, kMapUpdFU>()));
  PSGD_CUDA_CHECK((map_attr<R, kMapUpdFV>()));
  PSGD_CUDA_CHECK((map_attr<R, kMapUpdAppU>()));
  PSGD_CUDA_CHECK((map_attr<R, kMapUpdAppV>()));
  PSGD_CUDA_CHECK((map_attr<R, kMapApplyD>()));
  done.set(ctx->device);
  return PSGD_OK;
}

struct Scratch {
  float* partial;      // [grid][max table] floats
  double* G;           // reduced table
  SmallState* st;
  float* a; float* b; float* nd;   // N-vectors (update only)
};

// n_vecs N-vectors of scratch: 3 (a, b, nablaD) for the three-sweep update, 1 (nablaD) for the fused forms
static int carve(psgd_ctx* ctx, int64_t n, int r, int grid, int n_vecs, Scratch* s, int64_t extra_vec = 0) {
  const size_t table = (size_t)(2 * r + 2) * (2 * r + 2);      // >= every partial record (Gram tables, update+apply sums)
  size_t bytes = WsCarver::padded(sizeof(float) * table * grid) + WsCarver::padded(sizeof(double) * table) +
                 WsCarver::padded(sizeof(SmallState)) + n_vecs * WsCarver::padded(sizeof(float) * (size_t)n) +
                 WsCarver::padded(sizeof(float) * (size_t)extra_vec);
  PSGD_RETURN_IF(ctx->reserve(bytes));
  WsCarver c(ctx->ws);
  s->partial = c.take<float>(table * grid);
  s->G = c.take<double>(table);
  s->st = c.take<SmallState>(1);
  s->a = s->b = s->nd = nullptr;
  if (n_vecs >= 1) s->nd = c.take<float>(n);
  if (n_vecs >= 3) { s->a = c.take<float>(n); s->b = c.take<float>(n); }
  if (n_vecs == 0 && extra_vec) s->a = c.take<float>(extra_vec);
  return PSGD_OK;
}

// Everything between two sweeps: reduce the `nblocks` partial records of `count` floats into s.G (float64, fixed order),
// all-reduce it over the ranks together with n_max maxima at max_buf, run the r x r algebra `kind`.  One launch
// (uvd_mid_kernel) unless the torch.distributed hook is the exchange (host callback) or "uvd_mid" is off.
static int mid_step(psgd_ctx* ctx, const Scratch& s, int nblocks, int count, int kind, int r, int fused, int update_U,
                    float step, float tiny, float* max_buf, int n_max) {
  cudaStream_t st = ctx->stream;
  const bool hook_only = ctx->comm_world <= 0 && ctx->allreduce != nullptr;
  if (!ctx->opt_uvd_mid || hook_only) {
    reduce_partials_kernel<<<1, kReduceThreads, 0, st>>>(s.partial, nblocks, count, s.G);
    PSGD_LAUNCH_CHECK(ctx);
    PSGD_RETURN_IF(cross_rank_reduce(ctx, s.G, count, max_buf, n_max));
    if (kind == kMidSmall1) uvd_small1_kernel<<<1, 32, 0, st>>>(s.G, r, fused, update_U, step, tiny, s.st);
    else if (kind == kMidSmallUA) uvd_small_ua_kernel<<<1, 32, 0, st>>>(s.G, r, update_U, step, tiny, s.st);
    else if (kind == kMidSmallApply) uvd_small_apply_kernel<<<1, 32, 0, st>>>(s.G, r, s.st);
    else if (kind == kMidSmall2) uvd_small2_kernel<<<1, 32, 0, st>>>(s.G, r, update_U, step, tiny, s.st);
    if (kind != kMidNone) PSGD_LAUNCH_CHECK(ctx);
    return PSGD_OK;
  }
  comm::DeviceArgs da;
  PSGD_RETURN_IF(comm::device_args(ctx, count, n_max, &da));
  MidArgs a{};
  a.partial = s.partial; a.nblocks = nblocks; a.count = count; a.G = s.G; a.ticket = &s.st->ticket; a.st = s.st;
  a.kind = kind; a.r = r; a.fused = fused; a.update_U = update_U; a.step = step; a.tiny = tiny;
  a.peers = da.peers; a.rank = da.rank; a.world = da.world; a.max_buf = max_buf; a.n_max = n_max;
  a.host_err = da.host_err; a.timeout_ns = da.timeout_ns;
  PSGD_REQUIRE(count <= kMidMaxCount, PSGD_ERR_BAD_SHAPE, "mid kernel: record of %d entries exceeds %d", count, kMidMaxCount);
  const int ctas = count < kMidCtas ? count : kMidCtas;
  ProfScope prof(ctx, PSGD_K_UVD_MID);
  uvd_mid_kernel<<<ctas, kMidThreads, 0, st>>>(a);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

// psgd.py:562-567
static int balance_pass(psgd_ctx* ctx, float* U, float* V, int64_t count, SmallState* sst) {
  cudaStream_t st = ctx->stream;
  zero_small_kernel<<<1, 1, 0, st>>>(sst);
  PSGD_LAUNCH_CHECK(ctx);
  maxabs2_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(U, V, count, sst);
  PSGD_LAUNCH_CHECK(ctx);
  PSGD_RETURN_IF(cross_rank_reduce(ctx, nullptr, 0, &sst->maxU, 2));
  balance_rho_kernel<<<1, 1, 0, st>>>(sst);
  PSGD_LAUNCH_CHECK(ctx);
  balance_scale_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(U, V, count, sst);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

// Sweep 1 of every update form + the r x r algebra that follows it.
template <int R>
static int update_head(psgd_ctx* ctx, const SweepArgs& a1, const Scratch& s, int grid, int fused, int update_U,
                       float step, float tiny) {
  PSGD_RETURN_IF((launch_gram<R, kUpdate>(ctx, a1, s.partial, grid)));
  constexpr int table = GramPlan<R, kUpdate>::TABLE;
  return mid_step(ctx, s, grid, table, kMidSmall1, R, fused, update_U, step, tiny, nullptr, 0);
}

template <int R>
static int update_impl(psgd_ctx* ctx, float* U, float* V, float* d, const float* v, const float* h, int64_t n,
                       float step, float tiny, int balance, int update_U) {
  PSGD_RETURN_IF(ensure_attrs<R>(ctx));
  const int grid = grid_for(ctx, n);
  const int fused = ctx->opt_uvd_fused;
  Scratch s;
  PSGD_RETURN_IF(carve(ctx, n, R, grid, fused ? 1 : 3, &s));
  cudaStream_t st = ctx->stream;

  if (balance) PSGD_RETURN_IF(balance_pass(ctx, U, V, n * R, s.st));

  // sweep 1
  SweepArgs a1{};
  a1.zero_word = &s.st->ticket;
  a1.mat[0] = U; a1.mat[1] = V; a1.vec[0] = d; a1.vec[1] = h; a1.vec[2] = v; a1.n = n; a1.direct = ctx->opt_direct;
  PSGD_RETURN_IF((update_head<R>(ctx, a1, s, grid, fused, update_U, step, tiny)));

  if (fused) {
    // sweep 2: a, b, nablaD per row + the rank-2 update of U (or V), whose coefficients came out of the Gram table
    MapOut o2{};
    o2.o2 = s.nd; o2.st = s.st; o2.mat_out = update_U ? U : V;
    if (update_U) PSGD_RETURN_IF((launch_map<R, kMapUpdFU>(ctx, a1, o2, 1, grid)));
    else PSGD_RETURN_IF((launch_map<R, kMapUpdFV>(ctx, a1, o2, 0, grid)));
    PSGD_RETURN_IF(cross_rank_reduce(ctx, nullptr, 0, &s.st->max_nabla, 1));
    // pass 3: only d waits for max|nablaD|
    {
      ProfScope prof(ctx, PSGD_K_UVD_D_UPDATE);
      d_update_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(d, s.nd, n, step, tiny, s.st);
      PSGD_LAUNCH_CHECK(ctx);
    }
    return PSGD_OK;
  }

  // three-sweep form (psgd_set_option("uvd_fused", 0)): the reductions over a, b are taken directly in sweep 2
  MapOut o2{};
  o2.o0 = s.a; o2.o1 = s.b; o2.o2 = s.nd; o2.partial = s.partial; o2.st = s.st;
  PSGD_RETURN_IF((launch_map<R, kMapUpd2>(ctx, a1, o2, update_U, grid)));
  constexpr int cnt2 = 3 + 2 * R;
  PSGD_RETURN_IF(mid_step(ctx, s, grid, cnt2, kMidSmall2, R, 0, update_U, step, tiny, &s.st->max_nabla, 1));

  // sweep 3
  SweepArgs a3{};
  a3.zero_word = &s.st->ticket;
  a3.mat[0] = update_U ? U : V; a3.vec[0] = s.a; a3.vec[1] = s.b; a3.vec[2] = s.nd; a3.vec[3] = d;
  a3.n = n; a3.direct = ctx->opt_direct;
  MapOut o3{};
  o3.mat_out = update_U ? U : V; o3.vec_out = d; o3.st = s.st;
  if (update_U) PSGD_RETURN_IF((launch_map<R, kMapUpd3U>(ctx, a3, o3, 1, grid)));
  else PSGD_RETURN_IF((launch_map<R, kMapUpd3V>(ctx, a3, o3, 0, grid)));
  return PSGD_OK;
}

// update_precond_UVd_math_ followed by precond_grad_UVd_math on the updated state, as THREE sweeps:
//   1  Gram table of (U, V, d h, v/d)                                              reads U V d h v
//   2  a, b, nablaD per row, rank-2 update of U (or V) written back, and the apply's sums over the UPDATED rows
//      [U' V']^T [d g | d nablaD g] (U'^T U' is expanded through the table)        reads U V d h v g, writes U' (V'), nablaD
//   3  d' = d - mu_d d nablaD and out = d' (I + V'U'^T)(I + U'V'^T)(d' g)           reads U' V' d nablaD g, writes d', out
// Same results as the two separate calls; each of U, V is read three times per step instead of five.
template <int R>
static int update_apply_impl(psgd_ctx* ctx, float* U, float* V, float* d, const float* v, const float* h,
                             const float* g, float* out, int64_t n, float step, float tiny, int balance, int update_U) {
  PSGD_RETURN_IF(ensure_attrs<R>(ctx));
  const int grid = grid_for(ctx, n);
  Scratch s;
  PSGD_RETURN_IF(carve(ctx, n, R, grid, 1, &s));
  cudaStream_t st = ctx->stream;
  if (balance) PSGD_RETURN_IF(balance_pass(ctx, U, V, n * R, s.st));

  SweepArgs a1{};
  a1.zero_word = &s.st->ticket;
  a1.mat[0] = U; a1.mat[1] = V; a1.vec[0] = d; a1.vec[1] = h; a1.vec[2] = v; a1.vec[3] = g; a1.n = n;
  a1.direct = ctx->opt_direct;
  PSGD_RETURN_IF((update_head<R>(ctx, a1, s, grid, 1, update_U, step, tiny)));

  MapOut o2{};
  o2.o2 = s.nd; o2.st = s.st; o2.mat_out = update_U ? U : V; o2.partial = s.partial;
  if (update_U) PSGD_RETURN_IF((launch_map<R, kMapUpdAppU>(ctx, a1, o2, 1, grid)));
  else PSGD_RETURN_IF((launch_map<R, kMapUpdAppV>(ctx, a1, o2, 0, grid)));
  constexpr int cnt = updapp_count(R);
  PSGD_RETURN_IF(mid_step(ctx, s, grid, cnt, kMidSmallUA, R, 1, update_U, step, tiny, &s.st->max_nabla, 1));

  SweepArgs a3{};
  a3.zero_word = &s.st->ticket;
  a3.mat[0] = U; a3.mat[1] = V; a3.vec[0] = d; a3.vec[1] = s.nd; a3.vec[2] = g; a3.n = n; a3.direct = ctx->opt_direct;
  MapOut o3{};
  o3.o0 = out; o3.o1 = d; o3.st = s.st;
  PSGD_RETURN_IF((launch_map<R, kMapApplyD>(ctx, a3, o3, 0, grid)));
  return PSGD_OK;
}

template <int R>
static int apply_impl(psgd_ctx* ctx, const float* U, const float* V, const float* d, const float* g, float* out,
                      int64_t n) {
  PSGD_RETURN_IF(ensure_attrs<R>(ctx));
  const int grid = grid_for(ctx, n);
  Scratch s;
  PSGD_RETURN_IF(carve(ctx, n, R, grid, 0, &s));
  cudaStream_t st = ctx->stream;
  SweepArgs a{};
  a.zero_word = &s.st->ticket;
  a.mat[0] = U; a.mat[1] = V; a.vec[0] = d; a.vec[1] = g; a.n = n; a.direct = ctx->opt_direct;
  PSGD_RETURN_IF((launch_gram<R, kApply>(ctx, a, s.partial, grid)));
  constexpr int table = GramPlan<R, kApply>::TABLE;
  PSGD_RETURN_IF(mid_step(ctx, s, grid, table, kMidSmallApply, R, 0, 0, 0.f, 0.f, nullptr, 0));
  MapOut o{};
  o.o0 = out; o.st = s.st;
  PSGD_RETURN_IF((launch_map<R, kMapApply2>(ctx, a, o, 0, grid)));
  return PSGD_OK;
}

// param -= lr_params * min(max_norm / (||pre||_2 + tiny), 1) * pre (+ v)         psgd.py:750-762
__global__ void __launch_bounds__(256) clip_update_kernel(float* __restrict__ param, const float* __restrict__ pre,
                                                          const float* __restrict__ v, int64_t n, float lr_params,
                                                          float max_norm, float tiny, const double* __restrict__ sumsq) {
  const float grad_norm = sqrtf((float)sumsq[0]) + tiny;
  const float lr = lr_params * fminf(max_norm / grad_norm, 1.0f);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n4 = n / 4;
  float4* p4 = reinterpret_cast<float4*>(param);
  const float4* g4 = reinterpret_cast<const float4*>(pre);
  const float4* v4 = reinterpret_cast<const float4*>(v);
  for (int64_t i = tid; i < n4; i += stride) {
    float4 P = p4[i];
    const float4 G = g4[i];
    float4 u = make_float4(lr * G.x, lr * G.y, lr * G.z, lr * G.w);
    if (v) { const float4 X = v4[i]; u.x += X.x; u.y += X.y; u.z += X.z; u.w += X.w; }
    P.x -= u.x; P.y -= u.y; P.z -= u.z; P.w -= u.w;
    p4[i] = P;
  }
  for (int64_t i = n4 * 4 + tid; i < n; i += stride) {
    float u = lr * pre[i];
    if (v) u += v[i];
    param[i] -= u;
  }
}

// Tail of UVd.step (psgd.py:747-762): pre = P g, optional clipping by ||pre||_2, param -= lr pre (+ v).
//   no clipping (max_norm = inf): the parameter update is fused into the second apply sweep -- pre_grad is never
//                                 written unless the caller asks for it (pre_out);
//   clipping: the second sweep writes pre_grad and reduces its sum of squares (all-reduced when sharded), a streaming
//             kernel then applies the clipped learning rate computed on the device (no host sync).
template <int R>
static int step_tail_impl(psgd_ctx* ctx, const float* U, const float* V, const float* d, const float* g, float* param,
                          const float* v, float* pre_out, int64_t n, float lr_params, float max_norm, float tiny) {
  PSGD_RETURN_IF(ensure_attrs<R>(ctx));
  const int grid = grid_for(ctx, n);
  const bool clip = !isinf(max_norm);
  Scratch s;
  PSGD_RETURN_IF(carve(ctx, n, R, grid, 0, &s, (clip && !pre_out) ? n : 0));
  cudaStream_t st = ctx->stream;
  SweepArgs a{};
  a.zero_word = &s.st->ticket;
  a.mat[0] = U; a.mat[1] = V; a.vec[0] = d; a.vec[1] = g; a.n = n; a.direct = ctx->opt_direct;
  PSGD_RETURN_IF((launch_gram<R, kApply>(ctx, a, s.partial, grid)));
  constexpr int table = GramPlan<R, kApply>::TABLE;
  PSGD_RETURN_IF(mid_step(ctx, s, grid, table, kMidSmallApply, R, 0, 0, 0.f, 0.f, nullptr, 0));
  MapOut o{};
  o.st = s.st;
  if (!clip) {
    a.vec[2] = param; a.vec[3] = v ? v : param;
    o.o0 = param; o.o1 = pre_out; o.lr = lr_params; o.has_v = v ? 1 : 0;
    return launch_map<R, kMapApplyParam>(ctx, a, o, 0, grid);
  }
  float* pre = pre_out ? pre_out : s.a;
  o.o0 = pre; o.partial = s.partial;
  PSGD_RETURN_IF((launch_map<R, kMapApplyNorm>(ctx, a, o, 0, grid)));
  PSGD_RETURN_IF(mid_step(ctx, s, grid, 1, kMidNone, R, 0, 0, 0.f, 0.f, nullptr, 0));
  clip_update_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(param, pre, v, n, lr_params, max_norm, tiny, s.G);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

// IpUVtmatvec for one column: out = x + U (V^T x)
template <int R>
static int matvec_impl(psgd_ctx* ctx, const float* U, const float* V, const float* x, float* out, int64_t n) {
  PSGD_RETURN_IF(ensure_attrs<R>(ctx));
  const int grid = grid_for(ctx, n);
  Scratch s;
  PSGD_RETURN_IF(carve(ctx, n, R, grid, 0, &s));
  cudaStream_t st = ctx->stream;
  SweepArgs a{};
  a.zero_word = &s.st->ticket;
  a.mat[0] = U; a.mat[1] = V; a.vec[0] = x; a.n = n; a.direct = ctx->opt_direct;
  PSGD_RETURN_IF((launch_gram<R, kMatvec>(ctx, a, s.partial, grid)));
  constexpr int table = GramPlan<R, kMatvec>::TABLE;
  PSGD_RETURN_IF(mid_step(ctx, s, grid, table, kMidSmallApply, R, 0, 0, 0.f, 0.f, nullptr, 0));   // p = V^T x (t unused)
  SweepArgs a2{};
  a2.zero_word = &s.st->ticket;
  a2.mat[0] = U; a2.vec[0] = x; a2.n = n; a2.direct = ctx->opt_direct;
  MapOut o{};
  o.o0 = out; o.st = s.st;
  PSGD_RETURN_IF((launch_map<R, kMapMatvec2>(ctx, a2, o, 0, grid)));
  return PSGD_OK;
}

#define PSGD_RANK_SWITCH(r, CALL)                                                              \
  switch (r) {                                                                                 \
    case 1: return CALL(1);   case 2: return CALL(2);   case 3: return CALL(3);                \
    case 4: return CALL(4);   case 5: return CALL(5);   case 6: return CALL(6);                \
    case 7: return CALL(7);   case 8: return CALL(8);   case 9: return CALL(9);                \
    case 10: return CALL(10); case 11: return CALL(11); case 12: return CALL(12);              \
    case 13: return CALL(13); case 14: return CALL(14); case 15: return CALL(15);              \
    case 16: return CALL(16);                                                                  \
    default:                                                                                   \
      ::psgd::set_error("UVd rank %d not supported (1..%d)", (int)(r), kMaxRank);              \
      return PSGD_ERR_BAD_SHAPE;                                                               \
  }

int update(psgd_ctx* ctx, float* U, float* V, float* d, const float* v, const float* h, int64_t n, int r,
           float step, float tiny, int balance, int update_U) {
#define CALL(R) update_impl<R>(ctx, U, V, d, v, h, n, step, tiny, balance, update_U)
  PSGD_RANK_SWITCH(r, CALL)
#undef CALL
}
int update_apply(psgd_ctx* ctx, float* U, float* V, float* d, const float* v, const float* h, const float* g,
                 float* out, int64_t n, int r, float step, float tiny, int balance, int update_U) {
#define CALL(R) update_apply_impl<R>(ctx, U, V, d, v, h, g, out, n, step, tiny, balance, update_U)
  PSGD_RANK_SWITCH(r, CALL)
#undef CALL
}
int apply(psgd_ctx* ctx, const float* U, const float* V, const float* d, const float* g, float* out, int64_t n,
          int r) {
#define CALL(R) apply_impl<R>(ctx, U, V, d, g, out, n)
  PSGD_RANK_SWITCH(r, CALL)
#undef CALL
}
int step_tail(psgd_ctx* ctx, const float* U, const float* V, const float* d, const float* g, float* param, const float* v,
              float* pre_out, int64_t n, int r, float lr_params, float max_norm, float tiny) {
#define CALL(R) step_tail_impl<R>(ctx, U, V, d, g, param, v, pre_out, n, lr_params, max_norm, tiny)
  PSGD_RANK_SWITCH(r, CALL)
#undef CALL
}
int matvec(psgd_ctx* ctx, const float* U, const float* V, const float* x, float* out, int64_t n, int r) {
#define CALL(R) matvec_impl<R>(ctx, U, V, x, out, n)
  PSGD_RANK_SWITCH(r, CALL)
#undef CALL
}

}  // namespace uvd
}  // namespace psgd

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
using namespace psgd;

static int check_uvd_ptrs(const void* const* ptrs, int count) {
  for (int i = 0; i < count; ++i) {
    PSGD_REQUIRE(ptrs[i] != nullptr, PSGD_ERR_BAD_POINTER, "UVd: null device pointer (arg %d)", i);
    PSGD_REQUIRE(aligned16(ptrs[i]), PSGD_ERR_BAD_POINTER, "UVd: device pointer %d is not 16-byte aligned", i);
  }
  return PSGD_OK;
}

extern "C" int psgd_uvd_update(psgd_ctx* ctx, float* U, float* V, float* d, const float* v, const float* h,
                               int64_t n, int r, float step, float tiny, int balance, int update_U) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_REQUIRE(n >= 0 && r >= 1, PSGD_ERR_BAD_SHAPE, "UVd update: bad sizes n=%lld r=%d", (long long)n, r);
  // an empty shard still takes part in the cross-rank reductions when a hook is installed
  if (n == 0 && !is_sharded(ctx)) return PSGD_OK;
  const void* ptrs[] = {U, V, d, v, h};
  if (n > 0) PSGD_RETURN_IF(check_uvd_ptrs(ptrs, 5));
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  return uvd::update(ctx, U, V, d, v, h, n, r, step, tiny, balance, update_U);
}

extern "C" int psgd_uvd_update_apply(psgd_ctx* ctx, float* U, float* V, float* d, const float* v, const float* h,
                                     const float* g, float* out, int64_t n, int r, float step, float tiny, int balance,
                                     int update_U) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_REQUIRE(n >= 0 && r >= 1, PSGD_ERR_BAD_SHAPE, "UVd update+apply: bad sizes n=%lld r=%d", (long long)n, r);
  if (n == 0 && !is_sharded(ctx)) return PSGD_OK;
  const void* ptrs[] = {U, V, d, v, h, g, out};
  if (n > 0) PSGD_RETURN_IF(check_uvd_ptrs(ptrs, 7));
  PSGD_REQUIRE(n == 0 || (out != d && out != g), PSGD_ERR_BAD_POINTER, "UVd update+apply: out must not alias d or g");
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  return uvd::update_apply(ctx, U, V, d, v, h, g, out, n, r, step, tiny, balance, update_U);
}

extern "C" int psgd_uvd_apply(psgd_ctx* ctx, const float* U, const float* V, const float* d, const float* g,
                              float* out, int64_t n, int r) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_REQUIRE(n >= 0 && r >= 1, PSGD_ERR_BAD_SHAPE, "UVd apply: bad sizes n=%lld r=%d", (long long)n, r);
  if (n == 0 && !is_sharded(ctx)) return PSGD_OK;
  const void* ptrs[] = {U, V, d, g, out};
  if (n > 0) PSGD_RETURN_IF(check_uvd_ptrs(ptrs, 5));
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  return uvd::apply(ctx, U, V, d, g, out, n, r);
}

extern "C" int psgd_uvd_step_tail(psgd_ctx* ctx, const float* U, const float* V, const float* d, const float* g,
                                  float* param, const float* v, float* pre_out, int64_t n, int r, float lr_params,
                                  float grad_clip_max_norm, float tiny) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_REQUIRE(n >= 0 && r >= 1, PSGD_ERR_BAD_SHAPE, "UVd step tail: bad sizes n=%lld r=%d", (long long)n, r);
  PSGD_REQUIRE(grad_clip_max_norm > 0.f, PSGD_ERR_BAD_SHAPE, "UVd step tail: grad_clip_max_norm must be > 0 (inf = none)");
  if (n == 0 && !is_sharded(ctx)) return PSGD_OK;
  const void* ptrs[] = {U, V, d, g, param};
  if (n > 0) {
    PSGD_RETURN_IF(check_uvd_ptrs(ptrs, 5));
    PSGD_REQUIRE((!v || aligned16(v)) && (!pre_out || aligned16(pre_out)), PSGD_ERR_BAD_POINTER,
                 "UVd step tail: v / pre_out must be 16-byte aligned");
  }
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  return uvd::step_tail(ctx, U, V, d, g, param, v, pre_out, n, r, lr_params, grad_clip_max_norm, tiny);
}

// strided column gather/scatter for k > 1 right-hand sides
__global__ void psgd_col_copy_kernel(const float* __restrict__ src, int64_t sstride, float* __restrict__ dst,
                                     int64_t dstride, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i * dstride] = src[i * sstride];
}

extern "C" int psgd_ipuvt_matvec(psgd_ctx* ctx, const float* U, const float* V, const float* x, float* out,
                                 int64_t n, int r, int k) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_REQUIRE(n >= 0 && r >= 1 && k >= 1, PSGD_ERR_BAD_SHAPE, "IpUVtmatvec: bad sizes n=%lld r=%d k=%d",
               (long long)n, r, k);
  if (n == 0) return PSGD_OK;
  const void* ptrs[] = {U, V, x, out};
  PSGD_RETURN_IF(check_uvd_ptrs(ptrs, 4));
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  if (k == 1) return uvd::matvec(ctx, U, V, x, out, n, r);
  // k columns: gather each column to a contiguous vector, run the single-column path, scatter back
  float *xc = nullptr, *oc = nullptr;
  PSGD_CUDA_CHECK(cudaMallocAsync(&xc, sizeof(float) * n, ctx->stream));
  PSGD_CUDA_CHECK(cudaMallocAsync(&oc, sizeof(float) * n, ctx->stream));
  int rc = PSGD_OK;
  for (int c = 0; c < k && rc == PSGD_OK; ++c) {
    psgd_col_copy_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(x + c, k, xc, 1, n);
    ctx->launches++;
    rc = uvd::matvec(ctx, U, V, xc, oc, n, r);
    psgd_col_copy_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(oc, 1, out + c, k, n);
    ctx->launches++;
  }
  cudaFreeAsync(xc, ctx->stream);
  cudaFreeAsync(oc, ctx->stream);
  return rc;
}

// Pipeline geometry of the update Gram sweep for rank r (diagnostics / documentation): rows per stage, stages, threads.
extern "C" int psgd_uvd_plan_info(int r, int* tile_rows, int* stages, int* threads, int* roles) {
#define CALL(R)                                                                              \
  ([&]() {                                                                                   \
    using P = uvd::GramPlan<R, uvd::kUpdate>;                                                \
    using L = uvd::TileLayout<R, 2, 3, P::TILE>;                                             \
    *tile_rows = P::TILE; *stages = L::kStages; *threads = P::THREADS; *roles = P::NROLES;   \
    return (int)PSGD_OK;                                                                     \
  })()
  PSGD_REQUIRE(tile_rows && stages && threads && roles, PSGD_ERR_BAD_POINTER, "psgd_uvd_plan_info: null output");
  using uvd::kMaxRank;
  PSGD_RANK_SWITCH(r, CALL)
#undef CALL
}
