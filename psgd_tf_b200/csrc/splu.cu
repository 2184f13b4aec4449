// Sparse-LU (SPLU) preconditioner  Q = L U,  L = [L1 0; L2 diag(l3)],  U = [U1 U2; 0 diag(u3)]   (SURVEY.md section 8f).
//
// Replaces the TensorFlow op sequences of update_precond_splu (psgd.py:396-477) and precond_grad_splu (psgd.py:483-524)
// on the already concatenated vectors dx, dg, g: [n].  L12 = [L1; L2] is [n, r] row-major, U12 = [U1 U2] is [r, n]
// row-major, l3, u3: [n - r].
//
// Same shape of computation as UVd: every big operand is indexed by the parameter index j, so a call is a short
// chain of streaming passes over (L2, U2, l3, u3, dx2, dg2) separated by tiny r x r kernels (one warp):
//
//   update:  balance maxima -> small0 (rho, balanced corners)
//            pass 1 (reduce)      U2 dg2                                                         psgd.py:430
//            small 1              Ug1, Qg1, iUtx1                                                :430-436
//            pass 2 (map+reduce)  Qg2, iQtx2 per j;  L2^T iQtx2, L2^T Qg2                        :431-442
//            small 2              iQtx1, LtQg1, Pg1, iLiQtx1                                     :440-448
//            pass 3 (map+reduce)  Pg2, iPx2 per j;  U2 iPx2;  max|grad2|, max|grad3| of both factors  :443-474
//            small 3              iPx1, grad1 (both), steps, new L1 / U1, rank-2 coefficient vectors   :452-476
//            pass 4 (map)         new L2, l3, U2, u3                                             :464-478
//   apply:   pass A1 (reduce) U2 g2 -> small -> pass A2 (map+reduce) Qg2, L2^T Qg2 -> small -> pass A3 (map) result.
//
// The reference re-balances the factors first (L /= rho, U *= rho, :411-417); the passes apply rho on the fly, so the
// balanced factors are never materialised.  Cross-block reductions are two-stage and fixed-order (float64 final stage).
// r <= 32 (the reference's demo uses r = 10, demo_usage_of_all_preconditioners.py:45).
#include <math.h>

#include <mutex>
#include <set>
#include <utility>

#include "common.cuh"

namespace psgd {
namespace splu {

constexpr int kMaxR = 32;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct State {
  float rho;
  float stepL, stepU;
  float maxL, maxU;            // atomic max targets (grad2 / grad3 parts), non-negative
  float L1[kMaxR * kMaxR];     // balanced corners, row-major r x r
  float U1[kMaxR * kMaxR];
  float dx1[kMaxR], dg1[kMaxR];
  float Ug1[kMaxR], Qg1[kMaxR], iUtx1[kMaxR], iQtx1[kMaxR], LtQg1[kMaxR], Pg1[kMaxR], iLiQtx1[kMaxR];
  float cL1[kMaxR], cL2[kMaxR], cU1[kMaxR], cU2[kMaxR];
};

struct Args {
  const float* L12; const float* l3; const float* U12; const float* u3;   // inputs (unbalanced)
  const float* dx; const float* dg;                                        // [n]
  float* L12_out; float* l3_out; float* U12_out; float* u3_out;
  float* v0; float* v1; float* v2; float* v3;                              // [m] scratch vectors
  float* partial;                                                          // [grid][2 * kMaxR]
  State* st;
  long long n;
  int r;
  float step, tiny;
};

// block-wide fixed-order reduction of per-thread accumulators acc[0..cnt) -> partial[blockIdx.x][off + k]
template <int RM>
__device__ __forceinline__ void block_reduce_to(const float (&acc)[RM], int cnt, float* partial, int off) {
  __shared__ float red[kWarps][RM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < RM; ++k) {
    if (k < cnt) {
      const float s = warp_sum(acc[k]);
      if (lane == 0) red[warp][k] = s;
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < cnt) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w][threadIdx.x];
    partial[(size_t)blockIdx.x * 2 * kMaxR + off + threadIdx.x] = s;
  }
}

// out[k] = sum over blocks of partial[b][off + k], k < r, by one warp: lanes stride the blocks (independent loads), then a
// fixed butterfly in float64 -- the order depends only on (nblocks, lane count), so results are deterministic.
__device__ __forceinline__ void warp_sum_partials(const float* partial, int nblocks, int off, int r, float* out, int lane) {
  for (int k = 0; k < r; ++k) {
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += (double)partial[(size_t)b * 2 * kMaxR + off + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[k] = (float)s;
  }
  __syncwarp();
}

// ---- balance maxima (no abs, as the reference: psgd.py:411-412) ---------------------------------------------------
__global__ void __launch_bounds__(kThreads) max_kernel(Args a, float* __restrict__ pmax) {
  __shared__ float red[kWarps][2];
  const long long m = a.n - a.r;
  float ml = -INFINITY, mu = -INFINITY;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (long long)gridDim.x * blockDim.x) {
    ml = fmaxf(ml, a.l3[j]);
    mu = fmaxf(mu, a.u3[j]);
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < a.r) {
    ml = fmaxf(ml, a.L12[(size_t)threadIdx.x * a.r + threadIdx.x]);        // diag of L12 [n, r]
    mu = fmaxf(mu, a.U12[(size_t)threadIdx.x * a.n + threadIdx.x]);        // diag of U12 [r, n]
  }
  ml = warp_max(ml); mu = warp_max(mu);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = ml; red[threadIdx.x >> 5][1] = mu; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kWarps; ++w) { red[0][0] = fmaxf(red[0][0], red[w][0]); red[0][1] = fmaxf(red[0][1], red[w][1]); }
    pmax[2 * blockIdx.x] = red[0][0];
    pmax[2 * blockIdx.x + 1] = red[0][1];
  }
}

// rho, balanced corners, leading r entries of dx / dg.  update: balance = 1; apply: balance = 0 (rho = 1, v1 = g)
__global__ void __launch_bounds__(32) small0_kernel(Args a, const float* __restrict__ pmax, int nblocks, int balance) {
  State* st = a.st;
  const int lane = threadIdx.x, r = a.r;
  float rho = 1.0f;
  if (balance) {
    float ml = -INFINITY, mu = -INFINITY;
    for (int b = lane; b < nblocks; b += 32) { ml = fmaxf(ml, pmax[2 * b]); mu = fmaxf(mu, pmax[2 * b + 1]); }
    ml = warp_max(ml); mu = warp_max(mu);
    rho = sqrtf(ml / mu);                                                  // psgd.py:413
  }
  for (int e = lane; e < r * r; e += 32) {
    const int i = e / r, j = e % r;
    st->L1[e] = balance ? a.L12[(size_t)i * r + j] / rho : a.L12[(size_t)i * r + j];        // :414
    st->U1[e] = balance ? rho * a.U12[(size_t)i * a.n + j] : a.U12[(size_t)i * a.n + j];    // :416
  }
  if (lane < r) {
    st->dx1[lane] = a.dx ? a.dx[lane] : 0.f;
    st->dg1[lane] = a.dg[lane];
  }
  if (lane == 0) { st->rho = rho; st->maxL = 0.f; st->maxU = 0.f; }
}

// Everything a pass reads per parameter index j -- the row L2[j, :], the column U2[:, j], and up to NV vectors -- goes
// through ONE multi-stage cp.async ring in shared memory (SASS LDGSTS: no registers, no use-stall): tile t + S - 1 is
// streaming in while tile t is consumed, so the passes run at the rate the loads can be ISSUED, not at one DRAM latency
// per tile (round 1: direct loads, 16 resident warps per SM each stalled on its own 14 loads -- 30-53 % of HBM).
//   * L2 is [m, r] row-major: a thread that walked its own row would touch r scattered words (one 128-byte line per ~3
//     rows); the tile's 256 consecutive rows are one contiguous run of 256 r floats, copied fully coalesced (16 bytes at a
//     time when the run is aligned) and read back row-wise from shared memory (pass 4 also rewrites them there and
//     stores the run contiguously);
//   * U2 is [r, n] row-major: r coalesced 4-byte copies per thread, laid out [k][thread];
//   * vectors: one 4-byte copy each, laid out [v][thread].
// Values are staged raw; the reader applies the balancing factor.  One block barrier per tile: after it every thread
// has finished tile t - 1, whose stage is exactly the one the copies of tile t + S - 1 go to.
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(float* smem_dst, const float* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int RM>
constexpr int ring_stages() { return RM <= 16 ? 3 : 2; }

template <int RM, int NV>
struct Ring {
  static constexpr int S = ring_stages<RM>();
  static constexpr int kStageFloats = kThreads * (2 * RM + NV);
  static constexpr size_t kBytes = (size_t)S * kStageFloats * sizeof(float);
  float* base;
  const float* L2;            // nullptr: the pass does not read L2
  const float* U2;            // nullptr: the pass does not read U2
  const float* vec[NV];       // already offset so that index j addresses parameter r + j; nullptr entries are skipped
  long long m, n, first, stride;
  int r, t;
  bool l2_aligned, pair_ok;

  __device__ __forceinline__ float* stage(int s) const { return base + s * kStageFloats; }
  __device__ __forceinline__ int rows_at(long long j0) const { return (int)(m - j0 < kThreads ? m - j0 : kThreads); }
  __device__ __forceinline__ void issue(int s, long long j0) {
    if (j0 < m) {
      float* st = stage(s);
      const int rows = rows_at(j0);
      if (L2) {
        const float* p = L2 + (size_t)j0 * r;
        const int cnt = rows * r;                       // <= kThreads * RM
        if (l2_aligned && (cnt & 3) == 0) {
#pragma unroll
          for (int i = 0; i < (RM + 3) / 4; ++i) {
            const int e = 4 * ((int)threadIdx.x + i * kThreads);
            if (e < cnt) cp_async16(st + e, p + e);
          }
        } else {
#pragma unroll
          for (int i = 0; i < RM; ++i) {
            const int e = threadIdx.x + i * kThreads;
            if (e < cnt) cp_async4(st + e, p + e);
          }
        }
      }
      if (pair_ok && (rows & 1) == 0) {
        // two parameter indices per copy: half the copy instructions (r even, n even, 8-byte aligned arrays)
        if (2 * (int)threadIdx.x < rows) {
          const long long j = j0 + 2 * threadIdx.x;
          if (U2) {
            float* uc = st + kThreads * RM + 2 * threadIdx.x;
#pragma unroll
            for (int k = 0; k < RM; ++k)
              if (k < r) cp_async8(uc + k * kThreads, U2 + (size_t)k * n + j);
          }
          float* vv = st + kThreads * 2 * RM + 2 * threadIdx.x;
#pragma unroll
          for (int v = 0; v < NV; ++v)
            if (vec[v]) cp_async8(vv + v * kThreads, vec[v] + j);
        }
      } else if ((int)threadIdx.x < rows) {
        const long long j = j0 + threadIdx.x;
        if (U2) {
          float* uc = st + kThreads * RM + threadIdx.x;
#pragma unroll
          for (int k = 0; k < RM; ++k)
            if (k < r) cp_async4(uc + k * kThreads, U2 + (size_t)k * n + j);
        }
        float* vv = st + kThreads * 2 * RM + threadIdx.x;
#pragma unroll
        for (int v = 0; v < NV; ++v)
          if (vec[v]) cp_async4(vv + v * kThreads, vec[v] + j);
      }
    }
    cp_async_commit();                                  // one group per slot, empty or not: the wait counts groups
  }
  __device__ __forceinline__ void begin(float* smem, long long m_, long long n_, int r_) {
    base = smem; m = m_; n = n_; r = r_; t = 0;
    first = (long long)blockIdx.x * kThreads;
    stride = (long long)gridDim.x * kThreads;
    // a full tile is kThreads * r floats, a multiple of 4: every tile start is 16-byte aligned iff the array is
    l2_aligned = L2 && (reinterpret_cast<uintptr_t>(L2) & 15u) == 0;
    // tile starts are even; U2 row k starts at U2 + k n: pairs are 8-byte aligned iff the bases are and n is even
    pair_ok = (n & 1) == 0 && (!U2 || (reinterpret_cast<uintptr_t>(U2) & 7u) == 0);
    for (int v = 0; v < NV; ++v) pair_ok = pair_ok && (!vec[v] || (reinterpret_cast<uintptr_t>(vec[v]) & 7u) == 0);
    for (int s = 0; s < S - 1; ++s) issue(s, first + s * stride);
  }
  // Tile j0 (the t-th of this block) becomes readable; EVERY thread of the block calls this exactly once per tile.
  __device__ __forceinline__ float* acquire(long long j0) {
    cp_async_wait<S - 2>();
    __syncthreads();
    issue((t + S - 1) % S, j0 + (S - 1) * stride);
    float* st = stage(t % S);
    ++t;
    return st;
  }
  // accessors into a stage
  __device__ __forceinline__ static float* lrow(float* st, int r_) { return st + threadIdx.x * r_; }
  __device__ __forceinline__ static const float* ucol(const float* st) { return st + kThreads * RM + threadIdx.x; }   // [k * kThreads]
  __device__ __forceinline__ static float vecv(const float* st, int v) { return st[kThreads * 2 * RM + v * kThreads + threadIdx.x]; }
};

// ---- pass 1 / A1:  partial[b][k] = sum_j (rho U2[k, j]) w[j],  w = dg2 (update) or g2 (apply) ------------------------
template <int RM>
__global__ void __launch_bounds__(kThreads) pass1_kernel(Args a) {
  extern __shared__ __align__(16) float ring_mem[];
  const long long m = a.n - a.r;
  const int r = a.r;
  const float rho = a.st->rho;
  Ring<RM, 1> ring;
  ring.L2 = nullptr; ring.U2 = a.U12 + r; ring.vec[0] = a.dg + r;
  ring.begin(ring_mem, m, a.n, r);
  float acc[RM];
#pragma unroll
  for (int k = 0; k < RM; ++k) acc[k] = 0.f;
  for (long long j0 = ring.first; j0 < m; j0 += ring.stride) {
    const int rows = ring.rows_at(j0);
    const float* st = ring.acquire(j0);
    if ((int)threadIdx.x < rows) {
      const float wj = ring.vecv(st, 0);
      const float* uc = ring.ucol(st);
#pragma unroll
      for (int k = 0; k < RM; ++k)
        if (k < r) acc[k] = fmaf(rho * uc[k * kThreads], wj, acc[k]);
    }
  }
  cp_async_wait<0>();
  block_reduce_to<RM>(acc, r, a.partial, 0);
}

// ---- small 1: Ug1 = U1 dg1 + U2 dg2, Qg1 = L1 Ug1, iUtx1 = U1^-T dx1 (update only)          psgd.py:430-436 -----------
__global__ void __launch_bounds__(32) small1_kernel(Args a, int nblocks, int update) {
  State* st = a.st;
  __shared__ float sUg1[kMaxR], sp[kMaxR];
  const int lane = threadIdx.x, r = a.r;
  warp_sum_partials(a.partial, nblocks, 0, r, sp, lane);
  if (lane < r) {
    float s = 0.f;
    for (int k = 0; k < r; ++k) s = fmaf(st->U1[lane * r + k], st->dg1[k], s);
    s = s + sp[lane];
    sUg1[lane] = s;
    st->Ug1[lane] = s;
  }
  __syncwarp();
  if (lane < r) {
    float s = 0.f;
    for (int k = 0; k < r; ++k) s = fmaf(st->L1[lane * r + k], sUg1[k], s);
    st->Qg1[lane] = s;
  }
  if (update && lane == 0) {
    // U1^T x = dx1, reading only the upper triangle of U1 (U1^T lower -> forward substitution)
    float x[kMaxR];
    for (int i = 0; i < r; ++i) {
      float s = st->dx1[i];
      for (int k = 0; k < i; ++k) s -= st->U1[k * r + i] * x[k];
      x[i] = s / st->U1[i * r + i];
      st->iUtx1[i] = x[i];
    }
  }
}

// ---- pass 2: per j  Qg2, iQtx2 (stored);  partial sums L2^T iQtx2 (off 0) and L2^T Qg2 (off kMaxR)   psgd.py:431-442 ----
// apply (update = 0): Qg2 = L2 Ug1 + l3 u3 g2 (stored), partial L2^T Qg2                       psgd.py:507-512
template <int RM>
__global__ void __launch_bounds__(kThreads, RM <= 12 ? 2 : 1) pass2_kernel(Args a, int update) {
  extern __shared__ __align__(16) float ring_mem[];
  const long long m = a.n - a.r;
  const int r = a.r;
  const State* st = a.st;
  const float rho = st->rho;
  Ring<RM, 4> ring;
  ring.L2 = a.L12 + (size_t)r * r;
  ring.U2 = update ? a.U12 + r : nullptr;
  ring.vec[0] = a.l3; ring.vec[1] = a.u3; ring.vec[2] = a.dg + r; ring.vec[3] = update ? a.dx + r : nullptr;
  ring.begin(ring_mem, m, a.n, r);
  float ug1[RM], iu1[RM], pa[RM], pb[RM];
#pragma unroll
  for (int k = 0; k < RM; ++k) {
    ug1[k] = k < r ? st->Ug1[k] : 0.f;
    iu1[k] = (update && k < r) ? st->iUtx1[k] : 0.f;
    pa[k] = 0.f; pb[k] = 0.f;
  }
  for (long long j0 = ring.first; j0 < m; j0 += ring.stride) {
    const int rows = ring.rows_at(j0);
    float* sg = ring.acquire(j0);
    if ((int)threadIdx.x < rows) {
      const long long j = j0 + threadIdx.x;
      const float* tile = ring.lrow(sg, r);
      const float* uc = ring.ucol(sg);
      const float l3raw = ring.vecv(sg, 0), u3raw = ring.vecv(sg, 1), dgj = ring.vecv(sg, 2);
      const float dxj = update ? ring.vecv(sg, 3) : 0.f;
      const float l3 = update ? l3raw / rho : l3raw;
      const float u3 = update ? rho * u3raw : u3raw;
      float lrow[RM];
      float dotL = 0.f, dotU = 0.f;
#pragma unroll
      for (int k = 0; k < RM; ++k) {
        if (k < r) {
          lrow[k] = update ? tile[k] / rho : tile[k];
          dotL = fmaf(lrow[k], ug1[k], dotL);
          if (update) dotU = fmaf(rho * uc[k * kThreads], iu1[k], dotU);
        }
      }
      const float Ug2 = u3 * dgj;                                            // :431 / :507
      const float Qg2 = dotL + l3 * Ug2;                                     // :434 / :510
      a.v0[j] = Qg2;
      float iQtx2 = 0.f;
      if (update) {
        const float iUtx2 = (dxj - dotU) / u3;                               // :437
        iQtx2 = iUtx2 / l3;                                                  // :439
        a.v1[j] = iQtx2;
      }
#pragma unroll
      for (int k = 0; k < RM; ++k)
        if (k < r) { pa[k] = fmaf(lrow[k], iQtx2, pa[k]); pb[k] = fmaf(lrow[k], Qg2, pb[k]); }
    }
  }
  cp_async_wait<0>();
  block_reduce_to<RM>(pa, update ? r : 0, a.partial, 0);
  block_reduce_to<RM>(pb, r, a.partial, kMaxR);
}

// ---- small 2: iQtx1, LtQg1, Pg1, iLiQtx1                                                     psgd.py:440-448 ------------
// apply: LtQg1 = L1^T Qg1 + L2^T Qg2; out[:r] = U1^T LtQg1                                     psgd.py:512-515
__global__ void __launch_bounds__(32) small2_kernel(Args a, int nblocks, int update, float* __restrict__ out) {
  State* st = a.st;
  __shared__ float sLt[kMaxR], spa[kMaxR], spb[kMaxR];
  const int lane = threadIdx.x, r = a.r;
  if (update) warp_sum_partials(a.partial, nblocks, 0, r, spa, lane);
  warp_sum_partials(a.partial, nblocks, kMaxR, r, spb, lane);
  if (lane < r) {
    float s = 0.f;
    for (int k = 0; k < r; ++k) s = fmaf(st->L1[k * r + lane], st->Qg1[k], s);       // (L1^T Qg1)[lane]
    s = s + spb[lane];
    sLt[lane] = s;
    st->LtQg1[lane] = s;
  }
  __syncwarp();
  if (lane < r) {
    float s = 0.f;
    for (int k = 0; k < r; ++k) s = fmaf(st->U1[k * r + lane], sLt[k], s);           // (U1^T LtQg1)[lane]
    st->Pg1[lane] = s;
    if (!update) out[lane] = s;
  }
  if (update && lane == 0) {
    float x[kMaxR], y[kMaxR];
    // L1^T x = iUtx1 - L2^T iQtx2, lower triangle of L1 only (L1^T upper -> back substitution)      :440
    for (int i = r - 1; i >= 0; --i) {
      float s = st->iUtx1[i] - spa[i];
      for (int k = i + 1; k < r; ++k) s -= st->L1[k * r + i] * x[k];
      x[i] = s / st->L1[i * r + i];
      st->iQtx1[i] = x[i];
    }
    // L1 y = iQtx1 : forward substitution                                                            :448
    for (int i = 0; i < r; ++i) {
      float s = x[i];
      for (int k = 0; k < i; ++k) s -= st->L1[i * r + k] * y[k];
      y[i] = s / st->L1[i * r + i];
      st->iLiQtx1[i] = y[i];
    }
  }
}

// ---- pass 3: Pg2, iPx2 per j (stored); partial U2 iPx2; max |grad2|, |grad3| of both factors   psgd.py:443-474 ----------
// apply (update = 0): out[r + j] = U2[:, j] . LtQg1 + u3 l3 Qg2                                   psgd.py:513-516
template <int RM>
__global__ void __launch_bounds__(kThreads, RM <= 12 ? 2 : 1) pass3_kernel(Args a, int update, float* __restrict__ out) {
  extern __shared__ __align__(16) float ring_mem[];
  __shared__ float redm[kWarps][2];
  const long long m = a.n - a.r;
  const int r = a.r;
  State* st = a.st;
  const float rho = st->rho;
  float lt1[RM], il1[RM], qg1[RM], iq1[RM], pg1[RM], dx1[RM], pc[RM];
#pragma unroll
  for (int k = 0; k < RM; ++k) {
    const bool ok = k < r;
    lt1[k] = ok ? st->LtQg1[k] : 0.f;
    il1[k] = (ok && update) ? st->iLiQtx1[k] : 0.f;
    qg1[k] = (ok && update) ? st->Qg1[k] : 0.f;
    iq1[k] = (ok && update) ? st->iQtx1[k] : 0.f;
    pg1[k] = (ok && update) ? st->Pg1[k] : 0.f;
    dx1[k] = (ok && update) ? st->dx1[k] : 0.f;
    pc[k] = 0.f;
  }
  Ring<RM, 6> ring;
  ring.L2 = update ? a.L12 + (size_t)r * r : nullptr;           // the apply's third pass does not read L2
  ring.U2 = a.U12 + r;
  ring.vec[0] = a.l3; ring.vec[1] = a.u3; ring.vec[2] = a.v0;
  ring.vec[3] = update ? a.v1 : nullptr; ring.vec[4] = update ? a.dg + r : nullptr; ring.vec[5] = update ? a.dx + r : nullptr;
  ring.begin(ring_mem, m, a.n, r);
  float mxL = 0.f, mxU = 0.f;
  for (long long j0 = ring.first; j0 < m; j0 += ring.stride) {
    const int rows = ring.rows_at(j0);
    float* sg = ring.acquire(j0);
    if ((int)threadIdx.x < rows) {
      const long long j = j0 + threadIdx.x;
      const float* tile = ring.lrow(sg, r);
      const float* uc = ring.ucol(sg);
      const float l3raw = ring.vecv(sg, 0), u3raw = ring.vecv(sg, 1), Qg2 = ring.vecv(sg, 2);
      const float iQtx2 = update ? ring.vecv(sg, 3) : 0.f;
      const float dgj = update ? ring.vecv(sg, 4) : 0.f, dxj = update ? ring.vecv(sg, 5) : 0.f;
      const float l3 = update ? l3raw / rho : l3raw;
      const float u3 = update ? rho * u3raw : u3raw;
      float urow[RM];
      float dotU = 0.f, dotL = 0.f;
#pragma unroll
      for (int k = 0; k < RM; ++k) {
        if (k < r) {
          const float u = uc[k * kThreads];
          urow[k] = update ? rho * u : u;
          dotU = fmaf(urow[k], lt1[k], dotU);
          if (update) dotL = fmaf(tile[k] / rho, il1[k], dotL);
        }
      }
      const float LtQg2 = l3 * Qg2;                                           // :443 / :513
      const float Pg2 = dotU + u3 * LtQg2;                                    // :446 / :516
      if (!update) {
        out[r + j] = Pg2;
      } else {
        const float iLiQtx2 = (iQtx2 - dotL) / l3;                            // :449
        const float iPx2 = iLiQtx2 / u3;                                      // :451
        a.v2[j] = Pg2;
        a.v3[j] = iPx2;
        mxL = fmaxf(mxL, fabsf(Qg2 * Qg2 - iQtx2 * iQtx2));                   // grad3 of L   :458
        mxU = fmaxf(mxU, fabsf(Pg2 * dgj - dxj * iPx2));                      // grad3 of U   :471
#pragma unroll
        for (int k = 0; k < RM; ++k) {
          if (k < r) {
            pc[k] = fmaf(urow[k], iPx2, pc[k]);
            mxL = fmaxf(mxL, fabsf(Qg2 * qg1[k] - iQtx2 * iq1[k]));           // grad2 of L   :457
            mxU = fmaxf(mxU, fabsf(pg1[k] * dgj - dx1[k] * iPx2));            // grad2 of U   :470
          }
        }
      }
    }
  }
  cp_async_wait<0>();
  if (!update) return;
  block_reduce_to<RM>(pc, r, a.partial, 0);
  mxL = warp_max(mxL); mxU = warp_max(mxU);
  if ((threadIdx.x & 31) == 0) { redm[threadIdx.x >> 5][0] = mxL; redm[threadIdx.x >> 5][1] = mxU; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kWarps; ++w) { redm[0][0] = fmaxf(redm[0][0], redm[w][0]); redm[0][1] = fmaxf(redm[0][1], redm[w][1]); }
    atomic_max_nonneg(&st->maxL, redm[0][0]);
    atomic_max_nonneg(&st->maxU, redm[0][1]);
  }
}

// ---- small 3: iPx1, grad1 of both factors, steps, new L1 / U1, coefficient vectors            psgd.py:452-476 ------------
__global__ void __launch_bounds__(32) small3_kernel(Args a, int nblocks) {
  State* st = a.st;
  __shared__ float g1[kMaxR][kMaxR + 1];
  __shared__ float sx[kMaxR];
  __shared__ float spc[kMaxR];
  const int lane = threadIdx.x, r = a.r;
  warp_sum_partials(a.partial, nblocks, 0, r, spc, lane);
  if (lane == 0) {
    // U1 x = iLiQtx1 - U2 iPx2 : back substitution                                    :452
    for (int i = r - 1; i >= 0; --i) {
      float s = st->iLiQtx1[i] - spc[i];
      for (int k = i + 1; k < r; ++k) s -= st->U1[i * r + k] * sx[k];
      sx[i] = s / st->U1[i * r + i];
    }
  }
  __syncwarp();
  // ---- L: grad1 = tril(Qg1 Qg1^T - iQtx1 iQtx1^T)                                   :455-456
  float mx = 0.f;
  for (int e = lane; e < r * r; e += 32) {
    const int i = e / r, j = e % r;
    const float g = (j <= i) ? (st->Qg1[i] * st->Qg1[j] - st->iQtx1[i] * st->iQtx1[j]) : 0.f;
    g1[i][j] = g;
    mx = fmaxf(mx, fabsf(g));
  }
  mx = fmaxf(warp_max(mx), st->maxL);
  const float stepL = a.step / (mx + a.tiny);                                          // :462
  __syncwarp();
  for (int e = lane; e < r * r; e += 32) {                                            // newL1 = L1 - (step0 grad1) L1   :463
    const int i = e / r, j = e % r;
    float s = 0.f;
    for (int k = 0; k < r; ++k) s = fmaf(stepL * g1[i][k], st->L1[k * r + j], s);
    a.L12_out[(size_t)i * r + j] = st->L1[e] - s;
  }
  if (lane < r) {                                                                      // row vectors Qg1^T L1, iQtx1^T L1
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < r; ++k) { s1 = fmaf(st->Qg1[k], st->L1[k * r + lane], s1); s2 = fmaf(st->iQtx1[k], st->L1[k * r + lane], s2); }
    st->cL1[lane] = s1; st->cL2[lane] = s2;
  }
  __syncwarp();
  // ---- U: grad1 = triu(Pg1 dg1^T - dx1 iPx1^T)                                      :468-469
  mx = 0.f;
  for (int e = lane; e < r * r; e += 32) {
    const int i = e / r, j = e % r;
    const float g = (j >= i) ? (st->Pg1[i] * st->dg1[j] - st->dx1[i] * sx[j]) : 0.f;
    g1[i][j] = g;
    mx = fmaxf(mx, fabsf(g));
  }
  mx = fmaxf(warp_max(mx), st->maxU);
  const float stepU = a.step / (mx + a.tiny);                                          // :475
  __syncwarp();
  for (int e = lane; e < r * r; e += 32) {                                            // newU1 = U1 - U1 (step0 grad1)   :476
    const int i = e / r, j = e % r;
    float s = 0.f;
    for (int k = 0; k < r; ++k) s = fmaf(st->U1[i * r + k], stepU * g1[k][j], s);
    a.U12_out[(size_t)i * a.n + j] = st->U1[e] - s;
  }
  if (lane < r) {                                                                      // column vectors U1 Pg1, U1 dx1
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < r; ++k) { s1 = fmaf(st->U1[lane * r + k], st->Pg1[k], s1); s2 = fmaf(st->U1[lane * r + k], st->dx1[k], s2); }
    st->cU1[lane] = s1; st->cU2[lane] = s2;
  }
  if (lane == 0) { st->stepL = stepL; st->stepU = stepU; }
}

// ---- pass 4: new L2, l3, U2, u3                                                               psgd.py:464-478 ------------
template <int RM>
__global__ void __launch_bounds__(kThreads, RM <= 12 ? 2 : 1) pass4_kernel(Args a) {
  extern __shared__ __align__(16) float ring_mem[];
  const long long m = a.n - a.r;
  const int r = a.r;
  const State* st = a.st;
  const float rho = st->rho, stepL = st->stepL, stepU = st->stepU;
  float* L2o = a.L12_out + (size_t)r * r;
  float* U2o = a.U12_out + r;
  float cL1[RM], cL2[RM], cU1[RM], cU2[RM];
#pragma unroll
  for (int k = 0; k < RM; ++k) {
    const bool ok = k < r;
    cL1[k] = ok ? st->cL1[k] : 0.f; cL2[k] = ok ? st->cL2[k] : 0.f;
    cU1[k] = ok ? st->cU1[k] : 0.f; cU2[k] = ok ? st->cU2[k] : 0.f;
  }
  Ring<RM, 8> ring;
  ring.L2 = a.L12 + (size_t)r * r;
  ring.U2 = a.U12 + r;
  ring.vec[0] = a.l3; ring.vec[1] = a.u3; ring.vec[2] = a.v0; ring.vec[3] = a.v1; ring.vec[4] = a.v2; ring.vec[5] = a.v3;
  ring.vec[6] = a.dg + r; ring.vec[7] = a.dx + r;
  ring.begin(ring_mem, m, a.n, r);
  for (long long j0 = ring.first; j0 < m; j0 += ring.stride) {
    const int rows = ring.rows_at(j0);
    float* sg = ring.acquire(j0);
    if ((int)threadIdx.x < rows) {
      const long long j = j0 + threadIdx.x;
      float* tile = ring.lrow(sg, r);
      const float* uc = ring.ucol(sg);
      const float l3raw = ring.vecv(sg, 0), u3raw = ring.vecv(sg, 1);
      const float Qg2 = ring.vecv(sg, 2), iQtx2 = ring.vecv(sg, 3), Pg2 = ring.vecv(sg, 4), iPx2 = ring.vecv(sg, 5);
      const float dgj = ring.vecv(sg, 6), dxj = ring.vecv(sg, 7);
      const float l3 = l3raw / rho, u3 = rho * u3raw;
      const float g3L = Qg2 * Qg2 - iQtx2 * iQtx2;                           // :458
      const float g3U = Pg2 * dgj - dxj * iPx2;                              // :471
#pragma unroll
      for (int k = 0; k < RM; ++k) {
        if (k < r) {
          const float l = tile[k] / rho;
          tile[k] = l - stepL * (Qg2 * cL1[k] - iQtx2 * cL2[k]) - stepL * g3L * l;                          // :464
          const float u = rho * uc[k * kThreads];
          U2o[(size_t)k * a.n + j] = u - stepU * (cU1[k] * dgj - cU2[k] * iPx2) - stepU * g3U * u;        // :477
        }
      }
      a.l3_out[j] = l3 - stepL * g3L * l3;                                   // :465
      a.u3_out[j] = u3 - stepU * g3U * u3;                                   // :478
    }
    __syncthreads();
    {                                   // the updated rows leave as one contiguous, coalesced run; the stage is
      float* po = L2o + (size_t)j0 * r; // not refilled before the barrier of the next acquire()
      const int cnt = rows * r;
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        const int e = threadIdx.x + i * kThreads;
        if (e < cnt) po[e] = sg[e];
      }
    }
  }
  cp_async_wait<0>();
}

template <typename K>
static int ensure_smem(psgd_ctx* ctx, K kernel, size_t bytes) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  std::lock_guard<std::mutex> lock(mu);
  const std::pair<int, const void*> key(ctx->device, reinterpret_cast<const void*>(kernel));
  if (done.count(key)) return PSGD_OK;
  PSGD_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  done.insert(key);
  return PSGD_OK;
}

static int grid_for(const psgd_ctx* ctx, long long m) {
  long long b = (m + kThreads - 1) / kThreads;
  const long long cap = (long long)ctx->num_sms * 4;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

#define SPLU_LAUNCH(RM, NV, KERN, ...)                                                              \
  do {                                                                                              \
    PSGD_RETURN_IF(splu::ensure_smem(ctx, KERN<RM>, splu::Ring<RM, NV>::kBytes));                   \
    KERN<RM><<<grid, splu::kThreads, splu::Ring<RM, NV>::kBytes, st>>>(__VA_ARGS__);                \
  } while (0)
#define SPLU_DISPATCH(r, NV, KERN, ...)                               \
  do {                                                                \
    if ((r) <= 8) SPLU_LAUNCH(8, NV, KERN, __VA_ARGS__);              \
    else if ((r) <= 10) SPLU_LAUNCH(10, NV, KERN, __VA_ARGS__);       \
    else if ((r) <= 12) SPLU_LAUNCH(12, NV, KERN, __VA_ARGS__);       \
    else if ((r) <= 16) SPLU_LAUNCH(16, NV, KERN, __VA_ARGS__);       \
    else SPLU_LAUNCH(32, NV, KERN, __VA_ARGS__);                      \
  } while (0)

static int check_common(psgd_ctx* ctx, long long n, int r) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_REQUIRE(r >= 1 && r <= kMaxR && n >= r, PSGD_ERR_BAD_SHAPE, "SPLU: need 1 <= r <= %d and n >= r (n=%lld, r=%d)",
               kMaxR, n, r);
  return PSGD_OK;
}

}  // namespace splu
}  // namespace psgd

using namespace psgd;

extern "C" int psgd_splu_update(psgd_ctx* ctx, const float* L12, const float* l3, const float* U12, const float* u3,
                                const float* dx, const float* dg, float* L12_out, float* l3_out, float* U12_out,
                                float* u3_out, int64_t n, int r, float step, float tiny) {
  PSGD_RETURN_IF(splu::check_common(ctx, n, r));
  const long long m = n - r;
  PSGD_REQUIRE(L12 && U12 && dx && dg && L12_out && U12_out && (m == 0 || (l3 && u3 && l3_out && u3_out)),
               PSGD_ERR_BAD_POINTER, "SPLU update: null device pointer");
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  const int grid = splu::grid_for(ctx, m);
  const size_t bytes = WsCarver::padded(sizeof(splu::State)) + 4 * WsCarver::padded(sizeof(float) * (size_t)(m > 0 ? m : 1)) +
                       WsCarver::padded(sizeof(float) * (size_t)grid * 2 * splu::kMaxR) + WsCarver::padded(sizeof(float) * 2 * grid);
  PSGD_RETURN_IF(ctx->reserve(bytes));
  WsCarver c(ctx->ws);
  splu::Args a{};
  a.L12 = L12; a.l3 = l3; a.U12 = U12; a.u3 = u3; a.dx = dx; a.dg = dg;
  a.L12_out = L12_out; a.l3_out = l3_out; a.U12_out = U12_out; a.u3_out = u3_out;
  a.st = c.take<splu::State>(1);
  a.v0 = c.take<float>(m > 0 ? m : 1); a.v1 = c.take<float>(m > 0 ? m : 1);
  a.v2 = c.take<float>(m > 0 ? m : 1); a.v3 = c.take<float>(m > 0 ? m : 1);
  a.partial = c.take<float>((size_t)grid * 2 * splu::kMaxR);
  float* pmax = c.take<float>(2 * grid);
  a.n = n; a.r = r; a.step = step; a.tiny = tiny;
  cudaStream_t st = ctx->stream;
  splu::max_kernel<<<grid, splu::kThreads, 0, st>>>(a, pmax);
  PSGD_LAUNCH_CHECK(ctx);
  splu::small0_kernel<<<1, 32, 0, st>>>(a, pmax, grid, 1);
  PSGD_LAUNCH_CHECK(ctx);
  SPLU_DISPATCH(r, 1, splu::pass1_kernel, a);
  PSGD_LAUNCH_CHECK(ctx);
  splu::small1_kernel<<<1, 32, 0, st>>>(a, grid, 1);
  PSGD_LAUNCH_CHECK(ctx);
  SPLU_DISPATCH(r, 4, splu::pass2_kernel, a, 1);
  PSGD_LAUNCH_CHECK(ctx);
  splu::small2_kernel<<<1, 32, 0, st>>>(a, grid, 1, nullptr);
  PSGD_LAUNCH_CHECK(ctx);
  SPLU_DISPATCH(r, 6, splu::pass3_kernel, a, 1, (float*)nullptr);
  PSGD_LAUNCH_CHECK(ctx);
  splu::small3_kernel<<<1, 32, 0, st>>>(a, grid);
  PSGD_LAUNCH_CHECK(ctx);
  SPLU_DISPATCH(r, 8, splu::pass4_kernel, a);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

extern "C" int psgd_splu_apply(psgd_ctx* ctx, const float* L12, const float* l3, const float* U12, const float* u3,
                               const float* g, float* out, int64_t n, int r) {
  PSGD_RETURN_IF(splu::check_common(ctx, n, r));
  const long long m = n - r;
  PSGD_REQUIRE(L12 && U12 && g && out && (m == 0 || (l3 && u3)), PSGD_ERR_BAD_POINTER, "SPLU apply: null device pointer");
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  const int grid = splu::grid_for(ctx, m);
  const size_t bytes = WsCarver::padded(sizeof(splu::State)) + WsCarver::padded(sizeof(float) * (size_t)(m > 0 ? m : 1)) +
                       WsCarver::padded(sizeof(float) * (size_t)grid * 2 * splu::kMaxR);
  PSGD_RETURN_IF(ctx->reserve(bytes));
  WsCarver c(ctx->ws);
  splu::Args a{};
  a.L12 = L12; a.l3 = l3; a.U12 = U12; a.u3 = u3; a.dx = nullptr; a.dg = g;
  a.st = c.take<splu::State>(1);
  a.v0 = c.take<float>(m > 0 ? m : 1);
  a.partial = c.take<float>((size_t)grid * 2 * splu::kMaxR);
  a.n = n; a.r = r;
  cudaStream_t st = ctx->stream;
  splu::small0_kernel<<<1, 32, 0, st>>>(a, nullptr, 0, 0);
  PSGD_LAUNCH_CHECK(ctx);
  SPLU_DISPATCH(r, 1, splu::pass1_kernel, a);
  PSGD_LAUNCH_CHECK(ctx);
  splu::small1_kernel<<<1, 32, 0, st>>>(a, grid, 0);
  PSGD_LAUNCH_CHECK(ctx);
  SPLU_DISPATCH(r, 4, splu::pass2_kernel, a, 0);
  PSGD_LAUNCH_CHECK(ctx);
  splu::small2_kernel<<<1, 32, 0, st>>>(a, grid, 0, out);
  PSGD_LAUNCH_CHECK(ctx);
  SPLU_DISPATCH(r, 6, splu::pass3_kernel, a, 0, out);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}
