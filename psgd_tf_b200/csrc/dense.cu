// Dense (full-matrix) preconditioner: update and apply                       psgd.py:26-63
//
//   a = Q dg,  b = Q^-T dx,  Q' = Q - mu triu(a a^T - b b^T) Q,  mu = step / (max |triu(a a^T - b b^T)| + tiny)
//
// The reference forms the n x n gradient and multiplies it into Q (n^3).  The gradient is the upper triangle of a rank-2
// matrix, so the product is a pair of suffix scans down the columns of Q and needs no matrix product at all:
//
//   (triu(a a^T) Q)[i, k] = a_i * sum_{j >= i} a_j Q[j, k]
//
// i.e. Q'[i, k] = Q[i, k] - mu (a_i Sa[i, k] - b_i Sb[i, k]) with Sa[i, k] = sum_{j >= i} a_j Q[j, k] and Sb alike -- true
// for ANY square Q, triangular or not, which is what tf.matmul computes.  8192^2: 0.8 GB of HBM traffic (Q read twice,
// Q' written once) instead of 1.1 TFLOP of 3xTF32 tensor-core work.  The scans run over chunks of 128 rows: one pass
// leaves the column sums of every chunk, a tiny kernel turns them into the sums of everything BELOW each chunk, and
// the second pass rescans each chunk from registers and writes Q'.  Every sum has a fixed order (rows bottom to top
// inside a warp's 16 rows, warps bottom to top inside a chunk, chunks bottom to top), so results are run-to-run
// identical; against the reference's n^3 product they differ in rounding only (the tests hold both to 1e-5).
// `dense_scan = 0` keeps the n^3 route (SIMT outer product + tcgen05 GEMM) as the cross-check.
#include "common.cuh"
#include "kron_stream.cuh"
#include "linalg.cuh"
#include "gemm_tc.cuh"
#include "../../include/psgd_b200.h"

namespace psgd {
namespace dense {

constexpr int kChunk = 128;              // rows per chunk (8 warps x 16 rows)
constexpr int kStrip = 128;              // columns per CTA (32 lanes x 4, lanes on consecutive columns)
constexpr int kWarpRows = kChunk / 8;

// max over i <= j of |a_i a_j - b_i b_j|: the gradient is never stored.  One warp per row, lanes stride the columns.
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ a, const float* __restrict__ b, int n,
                                                     float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  float mx = 0.f;
  if (i < n) {
    const float ai = a[i], bi = b[i];
    for (int j = i + lane; j < n; j += 32) mx = fmaxf(mx, fabsf(fmaf(-bi, b[j], __fmul_rn(ai, a[j]))));
  }
  mx = warp_max(mx);
  if (lane == 0 && mx > 0.f) atomic_max_nonneg(out, mx);
}

// One 128 x 128 tile of Q in registers: v[r][q] = Q[row0 + 16 warp + r, col0 + lane + 32 q]; rows / columns beyond n are zeros.
struct Tile {
  float v[kWarpRows][4];
};
__device__ __forceinline__ void load_tile(const float* __restrict__ Q, int n, int row0, int col0, int warp, int lane, Tile& t) {
#pragma unroll
  for (int r = 0; r < kWarpRows; ++r) {
    const int row = row0 + warp * kWarpRows + r;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int col = col0 + lane + 32 * q;
      t.v[r][q] = (row < n && col < n) ? Q[(size_t)row * n + col] : 0.f;
    }
  }
}

// Column sums of one warp's 16 rows, bottom row first.
__device__ __forceinline__ void warp_rows_sum(const Tile& t, const float* wa, const float* wb, float (&sa)[4], float (&sb)[4]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) { sa[q] = 0.f; sb[q] = 0.f; }
#pragma unroll
  for (int r = kWarpRows - 1; r >= 0; --r)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      sa[q] = fmaf(wa[r], t.v[r][q], sa[q]);
      sb[q] = fmaf(wb[r], t.v[r][q], sb[q]);
    }
}

// grid = (column strips, chunks).  pass 0: P[0][chunk][col] = sum_{j in chunk} a_j Q[j, col], P[1][...] with b.
// pass 1: P holds the sums over all rows BELOW the chunk (suffix_kernel); rescans the chunk and writes Q'.
template <int PASS>
__global__ void __launch_bounds__(256) scan_kernel(const float* __restrict__ Q, int n, const float* __restrict__ a,
                                                   const float* __restrict__ b, float* P, int ldp, int chunks,
                                                   const float* __restrict__ maxabs, float step, float tiny, float* out) {
  __shared__ float wsum[8][2][kStrip];
  __shared__ float as[kChunk], bs[kChunk];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = blockIdx.x * kStrip, row0 = blockIdx.y * kChunk;
  if (threadIdx.x < kChunk) {
    const int row = row0 + threadIdx.x;
    as[threadIdx.x] = row < n ? a[row] : 0.f;
    bs[threadIdx.x] = row < n ? b[row] : 0.f;
  }
  Tile t;
  load_tile(Q, n, row0, col0, warp, lane, t);
  __syncthreads();
  const float* wa = as + warp * kWarpRows;
  const float* wb = bs + warp * kWarpRows;
  float sa[4], sb[4];
  warp_rows_sum(t, wa, wb, sa, sb);
#pragma unroll
  for (int q = 0; q < 4; ++q) { wsum[warp][0][lane + 32 * q] = sa[q]; wsum[warp][1][lane + 32 * q] = sb[q]; }
  __syncthreads();
  if (PASS == 0) {
    const int vec = threadIdx.x >> 7, c = threadIdx.x & (kStrip - 1);
    float s = 0.f;
#pragma unroll
    for (int w = 7; w >= 0; --w) s += wsum[w][vec][c];
    if (col0 + c < n) P[((size_t)vec * chunks + blockIdx.y) * ldp + col0 + c] = s;
    return;
  }
  const float mu = step / (*maxabs + tiny);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c = lane + 32 * q, col = col0 + c;
    float ra = 0.f, rb = 0.f;                               // everything below this warp's rows
    if (col < n) {
      ra = P[((size_t)0 * chunks + blockIdx.y) * ldp + col];
      rb = P[((size_t)1 * chunks + blockIdx.y) * ldp + col];
    }
    for (int w = 7; w > warp; --w) { ra += wsum[w][0][c]; rb += wsum[w][1][c]; }
    sa[q] = ra; sb[q] = rb;
  }
#pragma unroll
  for (int r = kWarpRows - 1; r >= 0; --r) {
    const int row = row0 + warp * kWarpRows + r;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      sa[q] = fmaf(wa[r], t.v[r][q], sa[q]);
      sb[q] = fmaf(wb[r], t.v[r][q], sb[q]);
      t.v[r][q] = fmaf(-mu, fmaf(-wb[r], sb[q], __fmul_rn(wa[r], sa[q])), t.v[r][q]);
    }
    if (row < n) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = col0 + lane + 32 * q;
        if (col < n) out[(size_t)row * n + col] = t.v[r][q];
      }
    }
  }
}

// P[vec][c][col] <- sum of P[vec][c'][col] over the chunks c' below c (c' > c), bottom chunk first.
__global__ void __launch_bounds__(256) suffix_kernel(float* P, int ldp, int chunks, int n) {
  const int col = blockIdx.x * 256 + threadIdx.x, vec = blockIdx.y;
  if (col >= n) return;
  float* p = P + (size_t)vec * chunks * ldp + col;
  float run = 0.f;
  for (int c = chunks - 1; c >= 0; --c) {
    const float t = p[(size_t)c * ldp];
    p[(size_t)c * ldp] = run;
    run += t;
  }
}

static int chunks_of(int n) { return (n + kChunk - 1) / kChunk; }
static int ldp_of(int n) { return (n + 31) / 32 * 32; }

}  // namespace dense
}  // namespace psgd

using namespace psgd;

extern "C" int psgd_dense_update(psgd_ctx* ctx, const float* Q, const float* dx, const float* dg, float* Q_out,
                                 int64_t n64, float step, float tiny) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_REQUIRE(n64 >= 1 && n64 < (1 << 20), PSGD_ERR_BAD_SHAPE, "dense update: n=%lld", (long long)n64);
  PSGD_REQUIRE(Q && dx && dg && Q_out, PSGD_ERR_BAD_POINTER, "dense update: null device pointer");
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  const int n = (int)n64;
  const bool scan = ctx->opt_dense_scan != 0;
  const int chunks = dense::chunks_of(n), ldp = dense::ldp_of(n);
  const size_t big = scan ? 2 * (size_t)chunks * ldp : (size_t)n * n;
  PSGD_RETURN_IF(ctx->reserve((big + 4 * (size_t)n + la::trsv_ws_floats(n)) * sizeof(float) + 32 * 256));
  WsCarver c(ctx->ws);
  float* mx = c.take<float>(8);
  float* a = c.take<float>(n);
  float* b = c.take<float>(n);
  float* tws = c.take<float>(la::trsv_ws_floats(n) ? la::trsv_ws_floats(n) : 1);
  float* bigbuf = c.take<float>(big);
  PSGD_CUDA_CHECK(cudaMemsetAsync(mx, 0, 8 * sizeof(float), ctx->stream));
  PSGD_RETURN_IF(ks::row_dot(ctx, Q, n, dg, n, n, a));                        // a = Q dg          psgd.py:38
  PSGD_RETURN_IF(la::trsv_left_upper_adjoint(ctx, Q, n, dx, b, n, tws));      // b = Q^-T dx       psgd.py:39
  if (!scan) {
    la::Gemm g2;                                                              // triu(a a^T - b b^T)   psgd.py:40
    g2.M = n; g2.N = n; g2.K = 1; g2.A = a; g2.lda = 1; g2.B = a; g2.ldb = 1; g2.tb = true; g2.C = bigbuf; g2.ldc = n;
    g2.K2 = 1; g2.A2 = b; g2.lda2 = 1; g2.B2 = b; g2.ldb2 = 1; g2.tb2 = true;
    g2.triu = true; g2.maxabs = mx;
    PSGD_RETURN_IF(la::gemm_simt(ctx, g2));
    la::Gemm g3;                                                              // Q - step0 grad Q  psgd.py:41-42
    g3.M = n; g3.N = n; g3.K = n; g3.A = bigbuf; g3.lda = n; g3.B = Q; g3.ldb = n; g3.C = Q_out; g3.ldc = n;
    g3.D = Q; g3.ldd = n; g3.mu_max = mx; g3.step = step; g3.tiny = tiny;
    g3.a_tri = 1;                                                             // the gradient is upper triangular by construction;
    return tc::gemm_many(ctx, &g3, 1, false);                                 // no hint on Q: the caller may pass any matrix
  }
  {
    ProfScope prof(ctx, PSGD_K_DENSE_SCAN, 12.0 * n * n);
    dense::absmax_kernel<<<(n + 7) / 8, 256, 0, ctx->stream>>>(a, b, n, mx);  // psgd.py:41
    PSGD_LAUNCH_CHECK(ctx);
    const dim3 grid((n + dense::kStrip - 1) / dense::kStrip, chunks);
    dense::scan_kernel<0><<<grid, 256, 0, ctx->stream>>>(Q, n, a, b, bigbuf, ldp, chunks, mx, step, tiny, nullptr);
    PSGD_LAUNCH_CHECK(ctx);
    dense::suffix_kernel<<<dim3((n + 255) / 256, 2), 256, 0, ctx->stream>>>(bigbuf, ldp, chunks, n);
    PSGD_LAUNCH_CHECK(ctx);
    dense::scan_kernel<1><<<grid, 256, 0, ctx->stream>>>(Q, n, a, b, bigbuf, ldp, chunks, mx, step, tiny, Q_out);   // psgd.py:42
    PSGD_LAUNCH_CHECK(ctx);
  }
  return PSGD_OK;
}

extern "C" int psgd_dense_apply(psgd_ctx* ctx, const float* Q, const float* g, float* out, int64_t n64) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_REQUIRE(n64 >= 1 && n64 < (1 << 20), PSGD_ERR_BAD_SHAPE, "dense apply: n=%lld", (long long)n64);
  PSGD_REQUIRE(Q && g && out, PSGD_ERR_BAD_POINTER, "dense apply: null device pointer");
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  const int n = (int)n64;
  // two bandwidth-bound GEMVs: Q is read exactly twice (8 n^2 bytes), coalesced both times
  PSGD_RETURN_IF(ctx->reserve(((size_t)n + (size_t)ks::row_tiles(n) * n) * sizeof(float) + 2048));
  WsCarver c(ctx->ws);
  float* t = c.take<float>(n);
  float* part = c.take<float>((size_t)ks::row_tiles(n) * n);
  PSGD_RETURN_IF(ks::row_dot(ctx, Q, n, g, n, n, t));                        // t = Q g            psgd.py:55
  return ks::col_wsum(ctx, 1, nullptr, t, Q, n, n, n, part, out);            // out = Q^T t
}
