// Bandwidth-bound fused kernels for the factor formats that contain no dense matrix:
//
//   (normalization, scaling) Kronecker pair      update psgd.py:328-369   apply psgd.py:372-391
//   full-matrix apply  Q^T (Q g)                 psgd.py:45-63   (and the GEMV a = Q dg of psgd.py:38)
//
// The reference runs ~12 TensorFlow element-wise / reduction ops per call, each a full pass over [M, N] with a
// temporary.  Here every [M, N] operand is read the algorithmic number of times:
//   update  dX twice + dG once = 12 MN bytes  (Bt's last row is a column reduction over dX that the row statistics need)
//   apply   G once, result once = 8 MN bytes
//   Q^T(Qg) Q twice             = 8 n^2 bytes
// and nothing of size MN is ever written except the apply's result.
//
// All kernels share one tiling: a CTA owns a kRows x kCols tile, warp w owns kRows/8 rows, lane l owns columns
// c0 + l + 32 k (k < 8) -- every load instruction is one fully coalesced 128-byte line, with no alignment requirement
// on N (NMT shapes such as [1025, 4935] are odd).  Row statistics are reduced with warp butterflies, column statistics
// per lane and then across the CTA's warps in shared memory; cross-CTA partials are combined by tiny finish kernels
// in fixed order, so results are deterministic.
#include <math.h>

#include "kron_stream.cuh"

namespace psgd {
namespace ks {

constexpr int kRows = 64, kCols = 256, kThreads = 256, kWarps = 8;
constexpr int kRowsPerWarp = kRows / kWarps;   // 8
constexpr int kColsPerLane = kCols / 32;       // 8

static dim3 tile_grid(int M, int N) { return dim3((N + kCols - 1) / kCols, (M + kRows - 1) / kRows); }
int col_tiles(int N) { return (N + kCols - 1) / kCols; }
int row_tiles(int M) { return (M + kRows - 1) / kRows; }

// The reducing kernels walk a CTA down a column strip over `rows_per_cta` rows (a multiple of kRows) instead of one
// 64-row tile: the per-CTA prologue (per-column constants, last rows) and epilogue (shared-memory reduction, partial
// stores) cost as much as streaming one 128 KB tile (ncu: 3.2 TB/s with one tile per CTA), and the cross-CTA partial
// tables shrink from M/64 records to a few.  The strips are sized so that the whole grid is ONE resident wave
// (`ctas_per_sm` = what the kernel's registers allow): with long-running CTAs a second, partly filled wave costs as much
// as a full one (ncu: 608 CTAs on 592 slots ran no faster than one tile per CTA).
static int rows_per_cta(const psgd_ctx* ctx, int M, int N, int ctas_per_sm) {
  const int rt = row_tiles(M);
  int chunks = (ctx->num_sms * ctas_per_sm) / col_tiles(N);
  if (chunks > rt) chunks = rt;
  if (chunks < 1) chunks = 1;
  return ((rt + chunks - 1) / chunks) * kRows;
}
static int row_chunks(int M, int rpc) { return (M + rpc - 1) / rpc; }

// ---------------------------------------------------------------------------------------------
// weighted column sums:  partial[row_tile][j] = sum_{i in tile} w(i) X[i, j]
//   mode 0: w = ql1[i] / (ql0[i] * ql0[M-1])        Bt's last-row correction       psgd.py:231-232
//   mode 1: w = wvec[i]                               Q^T t                          psgd.py:55 (second product)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) col_wsum_kernel(int mode, const float* __restrict__ ql,
                                                            const float* __restrict__ wvec, const float* __restrict__ X,
                                                            int ldx, int M, int N, int rpc, float* __restrict__ partial) {
  __shared__ float red[kWarps][kCols];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * kCols + lane;
  const int rbeg = blockIdx.y * rpc, rend = min(M, rbeg + rpc);
  float acc[kColsPerLane];
#pragma unroll
  for (int k = 0; k < kColsPerLane; ++k) acc[k] = 0.f;
  const float qlast = mode == 0 ? ql[M - 1] : 0.f;
  for (int rt = rbeg; rt < rend; rt += kRows) {
  const int r0 = rt + warp * kRowsPerWarp;
  for (int rr = 0; rr < kRowsPerWarp; rr += 4) {            // four rows of loads in flight per warp
    float xv[4][kColsPerLane], w[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int i = r0 + rr + h;
      const bool rok = i < M;
      const int ic = rok ? i : 0;
      w[h] = !rok ? 0.f : (mode == 0 ? ql[M + ic] / (ql[ic] * qlast) : wvec[ic]);
      const float* xr = X + (size_t)ic * ldx;
#pragma unroll
      for (int k = 0; k < kColsPerLane; ++k) {
        const int j = c0 + 32 * k;
        xv[h][k] = (rok && j < N) ? xr[j] : 0.f;
      }
    }
#pragma unroll
    for (int h = 0; h < 4; ++h)
#pragma unroll
      for (int k = 0; k < kColsPerLane; ++k) acc[k] = fmaf(w[h], xv[h][k], acc[k]);
  }
  }
#pragma unroll
  for (int k = 0; k < kColsPerLane; ++k) red[warp][lane + 32 * k] = acc[k];
  __syncthreads();
  const int j = blockIdx.x * kCols + threadIdx.x;
  if (j < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w][threadIdx.x];
    partial[(size_t)blockIdx.y * N + j] = s;
  }
}

// out[j] = sum over row tiles (fixed order) of partial[t][j]
__global__ void __launch_bounds__(256) col_finish_kernel(const float* __restrict__ partial, int tiles, int N,
                                                         float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  float s = 0.f;
#pragma unroll 8
  for (int t = 0; t < tiles; ++t) s += partial[(size_t)t * N + j];
  out[j] = s;
}

int col_wsum_partials(psgd_ctx* ctx, int mode, const float* ql, const float* wvec, const float* X, int ldx, int M, int N,
                      float* partial, int* chunks_out) {
  const int rpc = rows_per_cta(ctx, M, N, 4), chunks = row_chunks(M, rpc);       // 64 registers: 4 CTAs per SM
  col_wsum_kernel<<<dim3(col_tiles(N), chunks), kThreads, 0, ctx->stream>>>(mode, ql, wvec, X, ldx, M, N, rpc, partial);
  PSGD_LAUNCH_CHECK(ctx);
  *chunks_out = chunks;
  return PSGD_OK;
}

int col_wsum(psgd_ctx* ctx, int mode, const float* ql, const float* wvec, const float* X, int ldx, int M, int N,
             float* partial, float* out) {
  const int rpc = rows_per_cta(ctx, M, N, 4), chunks = row_chunks(M, rpc);       // 64 registers: 4 CTAs per SM
  col_wsum_kernel<<<dim3(col_tiles(N), chunks), kThreads, 0, ctx->stream>>>(mode, ql, wvec, X, ldx, M, N, rpc, partial);
  PSGD_LAUNCH_CHECK(ctx);
  col_finish_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(partial, chunks, N, out);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

// ---------------------------------------------------------------------------------------------
// row dots (GEMV):  out[i] = sum_j X[i, j] w[j]        t = Q g, a = Q dg        psgd.py:38, :55
// one warp per row, lanes stride the columns (coalesced), fixed butterfly
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float dot4(const float4 a, const float4 b, float s) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, fmaf(a.x, b.x, s))));
}
// vec != 0: rows and w are 16-byte aligned (ldx % 4 == 0): 128-bit loads, four per lane in flight
__global__ void __launch_bounds__(kThreads) row_dot_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ w,
                                                           int M, int N, int vec, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = blockIdx.x * kWarps + warp; i < M; i += gridDim.x * kWarps) {
    const float* xr = X + (size_t)i * ldx;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (vec) {
      const float4* x4 = reinterpret_cast<const float4*>(xr);
      const float4* w4 = reinterpret_cast<const float4*>(w);
      const int n4 = N >> 2;
      int q = lane;
      for (; q + 96 < n4; q += 128) {
        const float4 a0 = x4[q], a1 = x4[q + 32], a2 = x4[q + 64], a3 = x4[q + 96];
        s0 = dot4(a0, w4[q], s0); s1 = dot4(a1, w4[q + 32], s1);
        s2 = dot4(a2, w4[q + 64], s2); s3 = dot4(a3, w4[q + 96], s3);
      }
      for (; q < n4; q += 32) s0 = dot4(x4[q], w4[q], s0);
      for (int j = (n4 << 2) + lane; j < N; j += 32) s1 = fmaf(xr[j], w[j], s1);
      const float s = warp_sum((s0 + s1) + (s2 + s3));
      if (lane == 0) out[i] = s;
      continue;
    }
    int j = lane;
    for (; j + 96 < N; j += 128) {
      s0 = fmaf(xr[j], w[j], s0);
      s1 = fmaf(xr[j + 32], w[j + 32], s1);
      s2 = fmaf(xr[j + 64], w[j + 64], s2);
      s3 = fmaf(xr[j + 96], w[j + 96], s3);
    }
    for (; j < N; j += 32) s0 = fmaf(xr[j], w[j], s0);
    const float s = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) out[i] = s;
  }
}

int row_dot(psgd_ctx* ctx, const float* X, int ldx, const float* w, int M, int N, float* out) {
  int grid = (M + kWarps - 1) / kWarps;
  const int cap = ctx->num_sms * 8;
  if (grid > cap) grid = cap;
  const int vec = (ldx % 4 == 0) && aligned16(X) && aligned16(w);
  row_dot_kernel<<<grid, kThreads, 0, ctx->stream>>>(X, ldx, w, M, N, vec, out);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

// ---------------------------------------------------------------------------------------------
// (normalization, scaling) update, second pass: A and Bt are formed in registers from dG, dX and reduced at once
//   A [i,j] = (ql0[i] dG[i,j] + ql1[i] dG[M-1,j]) qr[j]                          psgd.py:349-351
//   Bt[i,j] = ((1/ql0[i]) dX[i,j] - [i == M-1] cvec[j]) (1/qr[j])                 psgd.py:353-356
//   row i : sum_j A^2, sum_j Bt^2, sum_j A A[M-1], sum_j Bt Bt[M-1]                psgd.py:358-359
//   col j : sum_i A^2, sum_i Bt^2                                                  psgd.py:366
// rowpart[col_tile][i][4], colpart[row_tile][2][N]
// ---------------------------------------------------------------------------------------------
// cpart/cchunks: the col_wsum partials of cvec (Bt's last-row correction), summed here in fixed order by every CTA for
// its own columns -- one launch less than finishing them separately, which matters at NMT sizes where the whole update
// is launch-latency bound.
__global__ void __launch_bounds__(kThreads) ns_stats_kernel(const float* __restrict__ ql, const float* __restrict__ qr,
                                                            const float* __restrict__ cpart, int cchunks,
                                                            const float* __restrict__ dX,
                                                            const float* __restrict__ dG, int M, int N, int rpc,
                                                            float* __restrict__ rowpart, float* __restrict__ colpart) {
  __shared__ float red[kWarps][2][kCols];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * kCols + lane;
  const int rbeg = blockIdx.y * rpc, rend = min(M, rbeg + rpc);
  const float* ql1 = ql + M;
  const float qlast0 = ql[M - 1], qlast1 = ql1[M - 1];
  const float rqlast0 = 1.0f / qlast0;
  // per-column constants of this lane: qr, 1/qr, A[M-1, j], Bt[M-1, j]
  float cq[kColsPerLane], crq[kColsPerLane], alast[kColsPerLane], blast[kColsPerLane], cv[kColsPerLane];
  const float* gl = dG + (size_t)(M - 1) * N;
  const float* xl = dX + (size_t)(M - 1) * N;
#pragma unroll
  for (int k = 0; k < kColsPerLane; ++k) {
    const int j = c0 + 32 * k;
    const bool ok = j < N;
    cq[k] = ok ? qr[j] : 0.f;
    crq[k] = ok ? 1.0f / qr[j] : 0.f;
    float c = 0.f;
    if (ok)
      for (int t = 0; t < cchunks; ++t) c += cpart[(size_t)t * N + j];
    cv[k] = c;
    const float g = ok ? gl[j] : 0.f, x = ok ? xl[j] : 0.f;
    alast[k] = (qlast0 * g + qlast1 * g) * cq[k];
    blast[k] = (rqlast0 * x - cv[k]) * crq[k];
  }
  float ca[kColsPerLane], cb[kColsPerLane], glast[kColsPerLane];
#pragma unroll
  for (int k = 0; k < kColsPerLane; ++k) {
    ca[k] = 0.f; cb[k] = 0.f;
    const int j = c0 + 32 * k;
    glast[k] = j < N ? gl[j] : 0.f;
  }
  // two rows per iteration, all 32 loads issued before any arithmetic: the kernel lives on memory-level parallelism
  for (int rt = rbeg; rt < rend; rt += kRows) {
  const int r0 = rt + warp * kRowsPerWarp;
  for (int rr = 0; rr < kRowsPerWarp; rr += 2) {
    const int i0 = r0 + rr;
    if (i0 >= M) break;
    const bool two = i0 + 1 < M;
    const int i1 = two ? i0 + 1 : i0;
    const float* g0 = dG + (size_t)i0 * N;
    const float* x0 = dX + (size_t)i0 * N;
    const float* g1 = dG + (size_t)i1 * N;
    const float* x1 = dX + (size_t)i1 * N;
    float gv0[kColsPerLane], xv0[kColsPerLane], gv1[kColsPerLane], xv1[kColsPerLane];
#pragma unroll
    for (int k = 0; k < kColsPerLane; ++k) {
      const int j = c0 + 32 * k;
      const bool ok = j < N;
      gv0[k] = ok ? g0[j] : 0.f; xv0[k] = ok ? x0[j] : 0.f;
      gv1[k] = ok ? g1[j] : 0.f; xv1[k] = ok ? x1[j] : 0.f;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      const int i = h ? i1 : i0;
      const float q0 = ql[i], q1 = ql1[i], rq0 = 1.0f / q0;
      const bool last = i == M - 1;
      float sa = 0.f, sb = 0.f, da = 0.f, db = 0.f;
#pragma unroll
      for (int k = 0; k < kColsPerLane; ++k) {
        if (c0 + 32 * k < N) {
          float a = q0 * (h ? gv1[k] : gv0[k]);
          a = a + q1 * glast[k];
          a = a * cq[k];
          float b = rq0 * (h ? xv1[k] : xv0[k]);
          if (last) b = b - cv[k];
          b = b * crq[k];
          sa = fmaf(a, a, sa); sb = fmaf(b, b, sb);
          da = fmaf(a, alast[k], da); db = fmaf(b, blast[k], db);
          ca[k] = fmaf(a, a, ca[k]); cb[k] = fmaf(b, b, cb[k]);
        }
      }
      sa = warp_sum(sa); sb = warp_sum(sb); da = warp_sum(da); db = warp_sum(db);
      if (lane == 0) {
        float* rp = rowpart + ((size_t)blockIdx.x * M + i) * 4;
        rp[0] = sa; rp[1] = sb; rp[2] = da; rp[3] = db;
      }
    }
  }
  }
#pragma unroll
  for (int k = 0; k < kColsPerLane; ++k) { red[warp][0][lane + 32 * k] = ca[k]; red[warp][1][lane + 32 * k] = cb[k]; }
  __syncthreads();
  const int j = blockIdx.x * kCols + threadIdx.x;
  if (j < N) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { s0 += red[w][0][threadIdx.x]; s1 += red[w][1][threadIdx.x]; }
    colpart[((size_t)blockIdx.y * 2 + 0) * N + j] = s0;
    colpart[((size_t)blockIdx.y * 2 + 1) * N + j] = s1;
  }
}

// grad1_diag[i], grad1_bias[i] (0 for the last row) and their max-abs                 psgd.py:358-362
__global__ void __launch_bounds__(256) ns_finish_rows_kernel(const float* __restrict__ rowpart, int ctiles, int M,
                                                             float* __restrict__ g1d, float* __restrict__ g1b,
                                                             float* __restrict__ max1) {
  float mx = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll 4
    for (int c = 0; c < ctiles; ++c) {
      const float4 p = *reinterpret_cast<const float4*>(rowpart + ((size_t)c * M + i) * 4);
      t0 += p.x; t1 += p.y; t2 += p.z; t3 += p.w;
    }
    const float d = t0 - t1;
    const float b = (i == M - 1) ? 0.f : (t2 - t3);
    g1d[i] = d; g1b[i] = b;
    mx = fmaxf(mx, fmaxf(fabsf(d), fabsf(b)));
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) atomic_max_nonneg(max1, mx);
}

// grad2[j] = colsum A^2 - colsum Bt^2 and its max-abs                                  psgd.py:366-367
__global__ void __launch_bounds__(256) ns_finish_cols_kernel(const float* __restrict__ colpart, int rtiles, int N,
                                                             float* __restrict__ grad2, float* __restrict__ max2) {
  float mx = 0.f;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
    for (int t = 0; t < rtiles; ++t) { s0 += colpart[((size_t)t * 2) * N + j]; s1 += colpart[((size_t)t * 2 + 1) * N + j]; }
    const float g = s0 - s1;
    grad2[j] = g;
    mx = fmaxf(mx, fabsf(g));
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) atomic_max_nonneg(max2, mx);
}

size_t ns_update_scratch_floats(int M, int N) {
  return (size_t)col_tiles(N) * M * 4 + (size_t)row_tiles(M) * 2 * N + 64;
}

// one CTA: grad1_diag / grad1_bias / grad2, both max-abs normalisers and the new factors               psgd.py:358-369
// (small problems only -- at most kFinishSmallWork partial records to sum, e.g. the NMT shapes; replaces four launches.
// At [8192, 8192] one CTA summing 336 K records was 40 us slower than the multi-CTA finish kernels.)
constexpr long long kFinishSmallWork = 128 * 1024;
__global__ void __launch_bounds__(1024) ns_finish_small_kernel(const float* __restrict__ rowpart, int ctiles,
                                                               const float* __restrict__ colpart, int rchunks,
                                                               const float* __restrict__ ql, const float* __restrict__ qr,
                                                               float* __restrict__ g1d, float* __restrict__ g1b,
                                                               float* __restrict__ grad2, float* __restrict__ ql_out,
                                                               float* __restrict__ qr_out, int M, int N, float step, float tiny) {
  __shared__ float red[32][2];
  float mx1 = 0.f, mx2 = 0.f;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    for (int c = 0; c < ctiles; ++c) {
      const float4 p = *reinterpret_cast<const float4*>(rowpart + ((size_t)c * M + i) * 4);
      t0 += p.x; t1 += p.y; t2 += p.z; t3 += p.w;
    }
    const float d = t0 - t1;
    const float b = (i == M - 1) ? 0.f : (t2 - t3);
    g1d[i] = d; g1b[i] = b;
    mx1 = fmaxf(mx1, fmaxf(fabsf(d), fabsf(b)));
  }
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
    for (int t = 0; t < rchunks; ++t) { s0 += colpart[((size_t)t * 2) * N + j]; s1 += colpart[((size_t)t * 2 + 1) * N + j]; }
    const float g = s0 - s1;
    grad2[j] = g;
    mx2 = fmaxf(mx2, fabsf(g));
  }
  mx1 = warp_max(mx1); mx2 = warp_max(mx2);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = mx1; red[threadIdx.x >> 5][1] = mx2; }
  __syncthreads();                                   // also orders this CTA's g1d / g1b / grad2 stores before the re-reads
  mx1 = 0.f; mx2 = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { mx1 = fmaxf(mx1, red[w][0]); mx2 = fmaxf(mx2, red[w][1]); }
  const float step1 = step / (mx1 + tiny);                                            // psgd.py:362
  const float step2 = step / (mx2 + tiny);                                            // psgd.py:367
  const float qlast = ql[M - 1];
  for (int i = threadIdx.x; i < M; i += blockDim.x) {                                 // psgd.py:363-364
    ql_out[i] = ql[i] - step1 * g1d[i] * ql[i];
    ql_out[M + i] = ql[M + i] - step1 * (g1d[i] * ql[M + i] + qlast * g1b[i]);
  }
  for (int j = threadIdx.x; j < N; j += blockDim.x) qr_out[j] = qr[j] - step2 * grad2[j] * qr[j];      // psgd.py:369
}

bool ns_finish_is_fused(int M, int N) {
  // upper bound of the records: col_tiles(N) per row + (at most row_tiles(M)) per column
  return (long long)M * col_tiles(N) + (long long)N * row_tiles(M) <= kFinishSmallWork;
}

int ns_update_stats(psgd_ctx* ctx, const float* ql, const float* qr, const float* cpart, int cchunks, const float* dX,
                    const float* dG, int M, int N, float* scratch, float* g1d, float* g1b, float* grad2, float* max1,
                    float* max2, float* ql_out, float* qr_out, float step, float tiny) {
  float* rowpart = scratch;
  float* colpart = scratch + (((size_t)col_tiles(N) * M * 4 + 63) / 64) * 64;
  const int rpc = rows_per_cta(ctx, M, N, 2), chunks = row_chunks(M, rpc);       // 128 registers: 2 CTAs per SM
  ns_stats_kernel<<<dim3(col_tiles(N), chunks), kThreads, 0, ctx->stream>>>(ql, qr, cpart, cchunks, dX, dG, M, N, rpc, rowpart,
                                                                            colpart);
  PSGD_LAUNCH_CHECK(ctx);
  if (ns_finish_is_fused(M, N)) {
    ns_finish_small_kernel<<<1, 1024, 0, ctx->stream>>>(rowpart, col_tiles(N), colpart, chunks, ql, qr, g1d, g1b, grad2,
                                                        ql_out, qr_out, M, N, step, tiny);
    PSGD_LAUNCH_CHECK(ctx);
    return PSGD_OK;
  }
  int gr = (M + 255) / 256, gc = (N + 255) / 256;
  if (gr > ctx->num_sms * 4) gr = ctx->num_sms * 4;
  if (gc > ctx->num_sms * 4) gc = ctx->num_sms * 4;
  ns_finish_rows_kernel<<<gr, 256, 0, ctx->stream>>>(rowpart, col_tiles(N), M, g1d, g1b, max1);
  PSGD_LAUNCH_CHECK(ctx);
  ns_finish_cols_kernel<<<gc, 256, 0, ctx->stream>>>(colpart, chunks, N, grad2, max2);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

// ---------------------------------------------------------------------------------------------
// (normalization, scaling) apply in one pass over G                                     psgd.py:383-389
//   P[i,j]   = (ql0[i] G[i,j] + ql1[i] G[M-1,j]) qr[j]^2
//   out[i,j] = ql0[i] P[i,j]                        (+ sum_i ql1[i] P[i,j] on the last row: partials + fix-up kernel)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) ns_apply_kernel(const float* __restrict__ ql, const float* __restrict__ qr,
                                                            const float* __restrict__ G, float* __restrict__ out, int M,
                                                            int N, float* __restrict__ colpart) {
  __shared__ float red[kWarps][kCols];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * kCols + lane, r0 = blockIdx.y * kRows + warp * kRowsPerWarp;
  const float* ql1 = ql + M;
  const float* gl = G + (size_t)(M - 1) * N;
  float cq2[kColsPerLane], glast[kColsPerLane], acc[kColsPerLane];
#pragma unroll
  for (int k = 0; k < kColsPerLane; ++k) {
    const int j = c0 + 32 * k;
    const float q = j < N ? qr[j] : 0.f;
    cq2[k] = q * q;
    glast[k] = j < N ? gl[j] : 0.f;
    acc[k] = 0.f;
  }
  for (int rr = 0; rr < kRowsPerWarp; rr += 4) {
    float gv[4][kColsPerLane];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int i = r0 + rr + h;
      const float* gr = G + (size_t)(i < M ? i : 0) * N;
#pragma unroll
      for (int k = 0; k < kColsPerLane; ++k) {
        const int j = c0 + 32 * k;
        gv[h][k] = (i < M && j < N) ? gr[j] : 0.f;
      }
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int i = r0 + rr + h;
      if (i < M) {
        const float q0 = ql[i], q1 = ql1[i];
        float* orow = out + (size_t)i * N;
#pragma unroll
        for (int k = 0; k < kColsPerLane; ++k) {
          const int j = c0 + 32 * k;
          if (j < N) {
            float p = q0 * gv[h][k];
            p = p + q1 * glast[k];
            p = p * cq2[k];
            acc[k] = fmaf(q1, p, acc[k]);
            orow[j] = q0 * p;          // the last row is completed by ns_apply_last_row_kernel
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kColsPerLane; ++k) red[warp][lane + 32 * k] = acc[k];
  __syncthreads();
  const int j = blockIdx.x * kCols + threadIdx.x;
  if (j < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w][threadIdx.x];
    colpart[(size_t)blockIdx.y * N + j] = s;
  }
}

__global__ void __launch_bounds__(256) ns_apply_last_row_kernel(const float* __restrict__ colpart, int rtiles, int M, int N,
                                                                float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  float s = 0.f;
#pragma unroll 8
  for (int t = 0; t < rtiles; ++t) s += colpart[(size_t)t * N + j];
  out[(size_t)(M - 1) * N + j] += s;                                                  // psgd.py:388-389
}

size_t ns_apply_scratch_floats(int M, int N) { return (size_t)row_tiles(M) * N + 64; }

int ns_apply(psgd_ctx* ctx, const float* ql, const float* qr, const float* G, float* out, int M, int N, float* scratch) {
  ns_apply_kernel<<<tile_grid(M, N), kThreads, 0, ctx->stream>>>(ql, qr, G, out, M, N, scratch);
  PSGD_LAUNCH_CHECK(ctx);
  ns_apply_last_row_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(scratch, row_tiles(M), M, N, out);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

}  // namespace ks
}  // namespace psgd
