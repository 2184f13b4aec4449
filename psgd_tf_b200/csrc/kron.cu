// Kronecker-product and dense (full-matrix) preconditioners: host orchestration + the structured-factor
// kernels ([2,N] normalization and [1,N] scaling formats).
//
// Replaces the TensorFlow op sequences of
//   update_precond_kron / precond_grad_kron dispatch   psgd.py:72-152
//   _update_precond_dense_dense / _precond_grad_...    psgd.py:156-192
//   _update_precond_norm_dense  / _precond_grad_...    psgd.py:198-270
//   _update_precond_dense_scale / _precond_grad_...    psgd.py:276-322
//   _update_precond_norm_scale  / _precond_grad_...    psgd.py:328-391
//   update_precond_dense / precond_grad_dense          psgd.py:26-63
// Large contractions go through the GEMM/TRSM vocabulary of linalg.cuh (SIMT fp32 or tcgen05 3xTF32,
// chosen per layer size); everything touching a [2,N]/[1,N] factor is a bandwidth-bound kernel here.
#include <algorithm>
#include <vector>

#include "linalg.cuh"
#include "gemm_tc.cuh"
#include "kron_stream.cuh"

namespace psgd {
namespace kron {

struct Scal {
  float max_l, max_r, rho;
  float max1, max2;
  float pad[3];
};

// ---------------------------------------------------------------------------------------------
// balance: rho = sqrt(max_l / max_r)            psgd.py:166-168, :211-213, :288-290, :342-344
// kind: 0 dense (diag of [n,n]), 1 scale ([1,n] row), 2 norm (row 0 of [2,n]).  max is taken WITHOUT abs.
// ---------------------------------------------------------------------------------------------
__device__ float factor_max(int kind, const float* Q, int n) {
  __shared__ float red[8];
  float m = -INFINITY;
  const size_t stride = (kind == PSGD_FACTOR_DENSE) ? (size_t)n + 1 : 1;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, Q[(size_t)i * stride]);
  m = warp_max(m);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, red[w]);
  return r;
}

__global__ void __launch_bounds__(256) balance_kernel(int kind_l, const float* Ql, int nl, int kind_r, const float* Qr,
                                                       int nr, Scal* sc) {
  const float ml = factor_max(kind_l, Ql, nl);
  const float mr = factor_max(kind_r, Qr, nr);
  if (threadIdx.x == 0) {
    sc->max_l = ml; sc->max_r = mr;
    sc->rho = sqrtf(ml / mr);
    sc->max1 = 0.f; sc->max2 = 0.f;
  }
}

// ---- grouped forms for dense factors: one launch for all layers of a group --------------------------------------
constexpr int kPrepBatch = 64;          // layers per launch (pointers travel as kernel parameters)
struct PrepBatch {
  const float* Ql[kPrepBatch];
  const float* Qr[kPrepBatch];
  Scal* sc[kPrepBatch];
  int* fl[kPrepBatch];                  // "factor is not upper triangular" flags (nullptr: not scanned)
  int* fr[kPrepBatch];
};

// grid = layers: maxima, rho and zeroed step maxima of every layer (balance_kernel for a whole group).  The rescaled
// copies Ql/rho, rho*Qr are NOT formed for (dense, dense) pairs: rho cancels in A = Ql dG Qr^T and in
// Bt = Ql^-T dX Qr^-1, so it only scales the returned factors and rides in the epilogue of the last product.
__global__ void __launch_bounds__(256) balance_many_kernel(const __grid_constant__ PrepBatch b, int kind_l, int nl,
                                                            int kind_r, int nr) {
  const int l = blockIdx.x;
  const float ml = factor_max(kind_l, b.Ql[l], nl);
  const float mr = factor_max(kind_r, b.Qr[l], nr);
  if (threadIdx.x == 0) {
    Scal* sc = b.sc[l];
    sc->max_l = ml; sc->max_r = mr;
    sc->rho = sqrtf(ml / mr);
    sc->max1 = 0.f; sc->max2 = 0.f;
  }
}

// grid = (row chunks, 2 x layers): flag = 1 when the strictly lower triangle of a dense factor holds anything but
// zeros.  The tensor-core products skip the K blocks that an upper-triangular factor leaves structurally zero; a
// caller may hand over ANY square matrix (tf.matmul multiplies all of it, psgd.py:173), and then the flag cancels
// the hint inside the GEMM kernel.  Reads n^2/2 floats per factor: 0.25 ms for the 48 factors of the 24 x 4096^2 stack.
constexpr int kScanRows = 32;
__global__ void __launch_bounds__(kPrepBatch) zero_flags_kernel(const __grid_constant__ PrepBatch b) {
  if (b.fl[threadIdx.x]) *b.fl[threadIdx.x] = 0;
  if (b.fr[threadIdx.x]) *b.fr[threadIdx.x] = 0;
}
__global__ void __launch_bounds__(256) tri_scan_kernel(const __grid_constant__ PrepBatch b, int nl, int nr) {
  const int l = blockIdx.y >> 1, side = blockIdx.y & 1;
  int* flag = side ? b.fr[l] : b.fl[l];
  if (!flag) return;
  const float* __restrict__ Q = side ? b.Qr[l] : b.Ql[l];
  const int n = side ? nr : nl;
  const int r0 = blockIdx.x * kScanRows;
  if (r0 >= n) return;
  const int r1 = min(n, r0 + kScanRows);
  const bool vec = (n & 3) == 0 && (reinterpret_cast<uintptr_t>(Q) & 15u) == 0;
  int any = 0;
  if (n <= 1024) {
    // small factor: the chunk as ONE flat strided loop, so that its loads are independent and in flight together
    // (a loop over rows is a chain of 32 short, latency-bound passes: 12 us at n = 257)
    const int total = (r1 - r0) * n;
#pragma unroll 4
    for (int e = threadIdx.x; e < total; e += 256) {
      const int i = r0 + e / n, j = e % n;
      if (j < i) any |= (Q[(size_t)i * n + j] != 0.f);
    }
    if (__syncthreads_or(any) && threadIdx.x == 0) *flag = 1;
    return;
  }
  for (int i = max(r0, 1); i < r1; ++i) {
    const float* row = Q + (size_t)i * n;
    int j0 = 0;
    if (vec) {
      const int n4 = i >> 2;
      const float4* row4 = reinterpret_cast<const float4*>(row);
      for (int j = threadIdx.x; j < n4; j += 256) {
        const float4 v = row4[j];
        any |= (v.x != 0.f) | (v.y != 0.f) | (v.z != 0.f) | (v.w != 0.f);
      }
      j0 = n4 << 2;
    }
    for (int j = j0 + threadIdx.x; j < i; j += 256) any |= (row[j] != 0.f);
  }
  if (__syncthreads_or(any) && threadIdx.x == 0) *flag = 1;
}

// Both factors structured (normalization / scaling: a few thousand floats): maxima, rho and both rescaled copies in ONE
// single-CTA launch instead of three (the update of such a pair is launch-latency bound at NMT sizes).
constexpr int64_t kBalanceSmallMax = 65536;
__global__ void __launch_bounds__(1024) balance_rescale_small_kernel(int kind_l, const float* __restrict__ Ql, int nl,
                                                                     int kind_r, const float* __restrict__ Qr, int nr,
                                                                     int64_t cl, int64_t cr, float* __restrict__ Qlb,
                                                                     float* __restrict__ Qrb, Scal* sc) {
  __shared__ float red[32][2];
  float ml = -INFINITY, mr = -INFINITY;            // row 0 of a [2,n] / [1,n] factor; max WITHOUT abs, as the reference
  for (int i = threadIdx.x; i < nl; i += blockDim.x) ml = fmaxf(ml, Ql[i]);
  for (int i = threadIdx.x; i < nr; i += blockDim.x) mr = fmaxf(mr, Qr[i]);
  ml = warp_max(ml); mr = warp_max(mr);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = ml; red[threadIdx.x >> 5][1] = mr; }
  __syncthreads();
  ml = red[0][0]; mr = red[0][1];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { ml = fmaxf(ml, red[w][0]); mr = fmaxf(mr, red[w][1]); }
  const float rho = sqrtf(ml / mr);
  if (threadIdx.x == 0) { sc->max_l = ml; sc->max_r = mr; sc->rho = rho; sc->max1 = 0.f; sc->max2 = 0.f; }
  for (int64_t i = threadIdx.x; i < cl; i += blockDim.x) Qlb[i] = Ql[i] / rho;
  for (int64_t i = threadIdx.x; i < cr; i += blockDim.x) Qrb[i] = rho * Qr[i];
}

// out = in / rho  (left factor)   or   out = rho * in  (right factor)
__global__ void __launch_bounds__(256) rescale_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                       int64_t count, const Scal* __restrict__ sc, int divide) {
  const float rho = sc->rho;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  if ((count & 3) == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {          // dense factors: 128-bit accesses
    const float4* in4 = reinterpret_cast<const float4*>(in);
    float4* out4 = reinterpret_cast<float4*>(out);
    for (int64_t i = tid; i < (count >> 2); i += stride) {
      float4 v = in4[i];
      if (divide) { v.x /= rho; v.y /= rho; v.z /= rho; v.w /= rho; }
      else { v.x *= rho; v.y *= rho; v.z *= rho; v.w *= rho; }
      out4[i] = v;
    }
    return;
  }
  for (int64_t i = tid; i < count; i += stride) out[i] = divide ? in[i] / rho : rho * in[i];
}

// ---------------------------------------------------------------------------------------------
// normalization-format left factor: Ql = diag(ql0) + e_last-column ql1  (ql1[M-1] == 0)
// ---------------------------------------------------------------------------------------------
// out[i,j] = (ql0[i] X[i,j] + ql1[i] X[M-1,j]) * cs(j)           psgd.py:218-219 (+ :351, :385)
// cs: none / qr[j] / qr[j]^2
__global__ void __launch_bounds__(256) norm_left_mul_kernel(const float* __restrict__ ql, const float* __restrict__ X,
                                                             float* __restrict__ out, int M, int N,
                                                             const float* __restrict__ qr, int qr_mode) {
  const int64_t total = (int64_t)M * N;
  const float* ql1 = ql + M;
  const float* Xlast = X + (size_t)(M - 1) * N;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / N), j = (int)(e % N);
    float v = ql[i] * X[e];
    v = v + ql1[i] * Xlast[j];
    if (qr_mode == 1) v = v * qr[j];
    else if (qr_mode == 2) v = v * (qr[j] * qr[j]);
    out[e] = v;
  }
}

// chunked, deterministic weighted column reductions: partial[chunk][which][j]
//   mode 0: sum_i w(i) X[i,j]            w = ql1[i] / (ql0[i] ql0[M-1])       psgd.py:232
//   mode 1: sum_i ql1[i] X[i,j]                                               psgd.py:265
//   mode 2: (sum_i X[i,j]^2, sum_i Y[i,j]^2)                                  psgd.py:304
constexpr int kColChunkRows = 128;
__global__ void __launch_bounds__(128) col_reduce_kernel(int mode, const float* __restrict__ ql,
                                                          const float* __restrict__ X, const float* __restrict__ Y,
                                                          int M, int N, float* __restrict__ partial) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y;
  const int i0 = chunk * kColChunkRows, i1 = min(M, i0 + kColChunkRows);
  if (j >= N) return;
  float s0 = 0.f, s1 = 0.f;
  const float qlast = (mode == 0) ? ql[M - 1] : 0.f;
  for (int i = i0; i < i1; ++i) {
    const float x = X[(size_t)i * N + j];
    if (mode == 0) s0 = fmaf(ql[M + i] / (ql[i] * qlast), x, s0);
    else if (mode == 1) s0 = fmaf(ql[M + i], x, s0);
    else { const float y = Y[(size_t)i * N + j]; s0 = fmaf(x, x, s0); s1 = fmaf(y, y, s1); }
  }
  partial[((size_t)chunk * 2 + 0) * N + j] = s0;
  partial[((size_t)chunk * 2 + 1) * N + j] = s1;
}
// out0[j] = sum_chunks partial[.][0][j]  (and out1 for mode 2)
__global__ void __launch_bounds__(128) col_finish_kernel(const float* __restrict__ partial, int chunks, int N,
                                                          float* __restrict__ out0, float* __restrict__ out1) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  float s0 = 0.f, s1 = 0.f;
  for (int c = 0; c < chunks; ++c) { s0 += partial[((size_t)c * 2) * N + j]; s1 += partial[((size_t)c * 2 + 1) * N + j]; }
  out0[j] = s0;
  if (out1) out1[j] = s1;
}

// Bt = Ql^-T dX (closed form), optionally times 1/qr[j]          psgd.py:230-232 (+ :356)
__global__ void __launch_bounds__(256) norm_left_solve_kernel(const float* __restrict__ ql, const float* __restrict__ X,
                                                               const float* __restrict__ cvec, float* __restrict__ out,
                                                               int M, int N, const float* __restrict__ qr) {
  const int64_t total = (int64_t)M * N;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / N), j = (int)(e % N);
    float v = (1.0f / ql[i]) * X[e];
    if (i == M - 1) v = v - cvec[j];
    if (qr) v = v * (1.0f / qr[j]);
    out[e] = v;
  }
}

// per-row statistics of A, Bt                                    psgd.py:235-239
//   g1d[i] = sum_j A^2 - sum_j Bt^2 ; g1b[i] = A[i].A[last] - Bt[i].Bt[last] (0 for the last row)
__global__ void __launch_bounds__(128) row_stats_kernel(const float* __restrict__ A, const float* __restrict__ Bt,
                                                         int M, int N, float* __restrict__ g1d,
                                                         float* __restrict__ g1b, Scal* __restrict__ sc) {
  __shared__ float red[4][4];
  const float* Al = A + (size_t)(M - 1) * N;
  const float* Bl = Bt + (size_t)(M - 1) * N;
  float mx = 0.f;
  for (int i = blockIdx.x; i < M; i += gridDim.x) {
    float sa = 0.f, sb = 0.f, da = 0.f, db = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      const float a = A[(size_t)i * N + j], b = Bt[(size_t)i * N + j];
      sa = fmaf(a, a, sa); sb = fmaf(b, b, sb);
      da = fmaf(a, Al[j], da); db = fmaf(b, Bl[j], db);
    }
    sa = warp_sum(sa); sb = warp_sum(sb); da = warp_sum(da); db = warp_sum(db);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
      const int w = threadIdx.x >> 5;
      red[w][0] = sa; red[w][1] = sb; red[w][2] = da; red[w][3] = db;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float t[4];
      for (int q = 0; q < 4; ++q) t[q] = red[0][q] + red[1][q] + red[2][q] + red[3][q];
      const float d = t[0] - t[1];
      const float b = (i == M - 1) ? 0.f : (t[2] - t[3]);
      g1d[i] = d; g1b[i] = b;
      mx = fmaxf(mx, fmaxf(fabsf(d), fabsf(b)));
    }
  }
  if (threadIdx.x == 0) atomic_max_nonneg(&sc->max1, mx);
}

// new ql rows                                                     psgd.py:240-241
__global__ void __launch_bounds__(256) norm_new_ql_kernel(const float* __restrict__ ql, const float* __restrict__ g1d,
                                                           const float* __restrict__ g1b, float* __restrict__ out, int M,
                                                           float step, float tiny, const float* __restrict__ maxabs) {
  const float step1 = step / (*maxabs + tiny);
  const float qlast = ql[M - 1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    out[i] = ql[i] - step1 * g1d[i] * ql[i];
    out[M + i] = ql[M + i] - step1 * (g1d[i] * ql[M + i] + qlast * g1b[i]);
  }
}

// grad2 = colsumA2 - colsumB2 ; max |grad2| -> sc->max2           psgd.py:304-305
__global__ void __launch_bounds__(128) scale_grad_kernel(const float* __restrict__ sa, const float* __restrict__ sb, int N,
                                                          float* __restrict__ grad2, float* __restrict__ maxabs) {
  float mx = 0.f;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
    const float g = sa[j] - sb[j];
    grad2[j] = g;
    mx = fmaxf(mx, fabsf(g));
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) atomic_max_nonneg(maxabs, mx);
}
// qr' = qr - step2 grad2 qr                                       psgd.py:307
__global__ void __launch_bounds__(128) scale_new_qr_kernel(const float* __restrict__ qr, const float* __restrict__ grad2,
                                                            float* __restrict__ out, int N, float step, float tiny,
                                                            const float* __restrict__ maxabs) {
  const float step2 = step / (*maxabs + tiny);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x)
    out[j] = qr[j] - step2 * grad2[j] * qr[j];
}

// out[i,j] = ql0[i] P[i,j]  (+ addlast[j] on the last row)        psgd.py:266-268
__global__ void __launch_bounds__(256) norm_left_out_kernel(const float* __restrict__ ql, const float* __restrict__ P,
                                                             const float* __restrict__ addlast, float* __restrict__ out,
                                                             int M, int N) {
  const int64_t total = (int64_t)M * N;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / N), j = (int)(e % N);
    float v = ql[i] * P[e];
    if (i == M - 1) v = v + addlast[j];
    out[e] = v;
  }
}

// out[i,j] = P[i,j] * qr[j]^2                                     psgd.py:322
__global__ void __launch_bounds__(256) col_scale_sq_kernel(const float* __restrict__ P, const float* __restrict__ qr,
                                                            float* __restrict__ out, int M, int N) {
  const int64_t total = (int64_t)M * N;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = P[e] * (qr[e % N] * qr[e % N]);
}

// X[i,j] *= 1/qr[j]                                              psgd.py:299
__global__ void __launch_bounds__(256) col_scale_recip_kernel(float* __restrict__ X, const float* __restrict__ qr,
                                                               int M, int N) {
  const int64_t total = (int64_t)M * N;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    X[e] = X[e] * (1.0f / qr[e % N]);
}

// (dense, normalization) in its own orientation (the reference transposes to (normalization, dense): psgd.py:86, :128).
// The normalization factor q = [q0; q1] then multiplies from the right,  X Qr^T = X diag(q0) + X[:, N-1] q1,  and every
// row of X is on its own: one warp per row.
//   mode 0: out = X Qr^T                                                                  psgd.py:218-219, :258-259
//   mode 1: out = X Qr^-1:  X[i,j] / q0[j], last column minus sum_j X[i,j] q1[j] / (q0[j] q0[N-1])      :230-232
//   mode 2: out[i,j] = X[i,j] q0[j], last column plus sum_j X[i,j] q1[j]                                 :265-268
// out may be X.
__global__ void __launch_bounds__(256) norm_right_kernel(int mode, const float* __restrict__ q, const float* X, float* out,
                                                          int M, int N) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* q1 = q + N;
  const float qlast = q[N - 1];
  for (int i = blockIdx.x * 8 + warp; i < M; i += gridDim.x * 8) {
    const float* x = X + (size_t)i * N;
    float* o = out + (size_t)i * N;
    const float xlast = x[N - 1];
    float s = 0.f;
    if (mode == 1) {
      for (int j = lane; j < N; j += 32) s = fmaf(q1[j] / (q[j] * qlast), x[j], s);
    } else if (mode == 2) {
      for (int j = lane; j < N; j += 32) s = fmaf(q1[j], x[j], s);
    }
    s = warp_sum(s);
    for (int j = lane; j < N; j += 32) {
      float v;
      if (mode == 0) { v = q[j] * x[j]; v = v + q1[j] * xlast; }
      else if (mode == 1) { v = (1.0f / q[j]) * x[j]; if (j == N - 1) v = v - s; }
      else { v = q[j] * x[j]; if (j == N - 1) v = v + s; }
      o[j] = v;
    }
  }
}
// column statistics of A, Bt for the normalization factor on the right           psgd.py:235-237 for the transposed problem
//   partial[chunk][w][j], w = 0..3:  sum_i A^2,  sum_i Bt^2,  sum_i A[i,j] A[i,N-1],  sum_i Bt[i,j] Bt[i,N-1]
__global__ void __launch_bounds__(128) norm_right_stats_kernel(const float* __restrict__ A, const float* __restrict__ Bt,
                                                                int M, int N, float* __restrict__ partial) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y;
  const int i0 = chunk * kColChunkRows, i1 = min(M, i0 + kColChunkRows);
  if (j >= N) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int i = i0; i < i1; ++i) {
    const float a = A[(size_t)i * N + j], b = Bt[(size_t)i * N + j];
    const float al = A[(size_t)i * N + N - 1], bl = Bt[(size_t)i * N + N - 1];
    s0 = fmaf(a, a, s0); s1 = fmaf(b, b, s1);
    s2 = fmaf(a, al, s2); s3 = fmaf(b, bl, s3);
  }
  float* p = partial + (size_t)chunk * 4 * N + j;
  p[0] = s0; p[N] = s1; p[2 * (size_t)N] = s2; p[3 * (size_t)N] = s3;
}
// d[j] = sum A^2 - sum Bt^2,  bias[j] = sum A A_last - sum Bt Bt_last (0 for the last column),  max of their moduli
__global__ void __launch_bounds__(128) norm_right_finish_kernel(const float* __restrict__ partial, int chunks, int N,
                                                                 float* __restrict__ d, float* __restrict__ bias,
                                                                 float* __restrict__ maxabs) {
  float mx = 0.f;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < chunks; ++c)
      for (int w = 0; w < 4; ++w) t[w] += partial[((size_t)c * 4 + w) * N + j];
    const float dj = t[0] - t[1];
    const float bj = (j == N - 1) ? 0.f : (t[2] - t[3]);
    d[j] = dj; bias[j] = bj;
    mx = fmaxf(mx, fmaxf(fabsf(dj), fabsf(bj)));
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) atomic_max_nonneg(maxabs, mx);
}

// (scaling, dense) in its own orientation (the reference transposes to (dense, scaling): psgd.py:102-104).  One warp per
// row i:  A[i,:] = ql[i] T[i,:],  Bt[i,:] = W[i,:] / ql[i]  in place, and the row sums of their squares -- the scaling
// factor's gradient (psgd.py:304 for the transposed problem) -- in the same pass.
__global__ void __launch_bounds__(256) scale_rows_stats_kernel(float* __restrict__ A, float* __restrict__ Bt,
                                                                const float* __restrict__ ql, int M, int N,
                                                                float* __restrict__ sa, float* __restrict__ sb) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + warp; i < M; i += gridDim.x * 8) {
    const float q = ql[i], rq = 1.0f / q;
    float* a = A + (size_t)i * N;
    float* b = Bt + (size_t)i * N;
    float s0 = 0.f, s1 = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float x = a[j] * q, y = b[j] * rq;
      a[j] = x; b[j] = y;
      s0 = fmaf(x, x, s0); s1 = fmaf(y, y, s1);
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1);
    if (lane == 0) { sa[i] = s0; sb[i] = s1; }
  }
}
// X[i,:] *= ql[i]^2                                               psgd.py:321 for the transposed problem
__global__ void __launch_bounds__(256) row_scale_sq_kernel(float* __restrict__ X, const float* __restrict__ ql, int M, int N) {
  const int64_t total = (int64_t)M * N;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const float q = ql[e / N];
    X[e] = X[e] * (q * q);
  }
}

// ---------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------
static int ew_grid(const psgd_ctx* ctx, int64_t total, int threads) {
  int64_t b = (total + threads - 1) / threads;
  const int64_t cap = (int64_t)ctx->num_sms * 8;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

static int col_reduce(psgd_ctx* ctx, int mode, const float* ql, const float* X, const float* Y, int M, int N,
                      float* partial, float* out0, float* out1) {
  const int chunks = (M + kColChunkRows - 1) / kColChunkRows;
  dim3 grid((N + 127) / 128, chunks);
  col_reduce_kernel<<<grid, 128, 0, ctx->stream>>>(mode, ql, X, Y, M, N, partial);
  PSGD_LAUNCH_CHECK(ctx);
  col_finish_kernel<<<(N + 127) / 128, 128, 0, ctx->stream>>>(partial, chunks, N, out0, out1);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}
static size_t col_partial_floats(int M, int N) {
  return (size_t)((M + kColChunkRows - 1) / kColChunkRows) * 2 * N;
}

constexpr int kUpper = 1, kLower = 2;

// Dense Kron/Cholesky factors are upper triangular by construction (identity-initialised and only ever updated as
// Q - triu(.) Q), which lets the tensor-core engine skip structurally zero K blocks.  The hints
// are dropped with psgd_set_option(ctx, "assume_triangular", 0) for callers that feed arbitrary square matrices.
// All layers of a group share shapes, so each op of the reference's sequence is ONE (grouped) launch over the group.
// a_src / b_src say which caller-supplied factor a hint rests on (kFromL / kFromR: gs[i] belongs to (*Ls)[i], whose
// run-time flag cancels the hint inside the kernel when that factor is not triangular); 0 = triangular by construction.
constexpr int kFromL = 1, kFromR = 2;
struct Layer;
static const int* flag_of(const Layer& L, int src);
static int gemm_all(psgd_ctx* ctx, std::vector<la::Gemm>& gs, int a_tri = 0, int b_tri = 0,
                    const std::vector<Layer>* Ls = nullptr, int a_src = 0, int b_src = 0) {
  if (ctx->opt_assume_tri)
    for (size_t i = 0; i < gs.size(); ++i) {
      la::Gemm& g = gs[i];
      g.a_tri = a_tri; g.b_tri = b_tri;
      if (Ls && a_src) g.a_full = flag_of((*Ls)[i], a_src);
      if (Ls && b_src) g.b_full = flag_of((*Ls)[i], b_src);
    }
  return tc::gemm_many(ctx, gs.data(), (int)gs.size(), false);
}

static size_t fsize(int kind, int64_t n) {
  return kind == PSGD_FACTOR_DENSE ? (size_t)n * n : (kind == PSGD_FACTOR_NORM ? 2 * (size_t)n : (size_t)n);
}

// One layer in canonical orientation ((D,D), (N,D), (D,S), (N,S)); mirrored formats arrive here transposed/swapped.
struct Layer {
  const float *Ql, *Qr, *dX, *dG, *G;
  float *Ql_out, *Qr_out, *out;
  // scratch
  Scal* sc;
  float *Qlb, *Qrb, *A, *Bt, *T1, *cvec, *part, *grad1, *grad2, *g1d, *g1b, *gvl, *sa, *sb, *gvec, *zinv, *xwork;
  float *t1, *t2, *t3, *P, *addlast;
  float* nrpart;   // (dense, normalization): partial column statistics, 4 per column and chunk
  float* nspart;   // (normalization, scaling): partial tables of the fused streaming kernels
  int *fl, *fr;    // run-time "dense factor is not upper triangular" flags (nullptr: no scan, hints taken as given)
};

static const int* flag_of(const Layer& L, int src) { return src == kFromL ? L.fl : (src == kFromR ? L.fr : nullptr); }

// Dense factors whose products may run on the tensor cores with triangular K-range hints get a run-time check.
static bool wants_scan(const psgd_ctx* ctx, int kind, int n) {
  // (a factor whose order is not a multiple of 4 never reaches the tensor-core engine: gemm_tc_supported)
  return kind == PSGD_FACTOR_DENSE && ctx->opt_assume_tri && ctx->opt_gemm_path != 1 && (n % 4) == 0 &&
         (ctx->opt_gemm_path == 2 || n >= 256);
}

// flags of one group: zeroed, then set by tri_scan_kernel (one launch per kPrepBatch layers)
static int scan_group(psgd_ctx* ctx, std::vector<Layer>& Ls, int M, int N);

static int scan_group(psgd_ctx* ctx, std::vector<Layer>& Ls, int M, int N) {
  if (Ls.empty() || (!Ls[0].fl && !Ls[0].fr)) return PSGD_OK;
  const int nmax = M > N ? M : N;
  for (size_t t0 = 0; t0 < Ls.size(); t0 += kPrepBatch) {
    const int cnt = (int)std::min<size_t>(kPrepBatch, Ls.size() - t0);
    PrepBatch b{};
    for (int t = 0; t < cnt; ++t) {
      const Layer& L = Ls[t0 + t];
      b.Ql[t] = L.Ql; b.Qr[t] = L.Qr; b.fl[t] = L.fl; b.fr[t] = L.fr;
    }
    zero_flags_kernel<<<1, kPrepBatch, 0, ctx->stream>>>(b);
    PSGD_LAUNCH_CHECK(ctx);
    tri_scan_kernel<<<dim3((nmax + kScanRows - 1) / kScanRows, 2 * cnt), 256, 0, ctx->stream>>>(b, M, N);
    PSGD_LAUNCH_CHECK(ctx);
  }
  return PSGD_OK;
}

static la::Gemm mk(int M, int N, int K, const float* A, int lda, bool ta, const float* B, int ldb, bool tb, float* C,
                   int ldc) {
  la::Gemm g;
  g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.ta = ta; g.B = B; g.ldb = ldb; g.tb = tb; g.C = C; g.ldc = ldc;
  return g;
}

// ---------------------------------------------------------------------------------------------
// canonical update of a group of same-shape layers
// ---------------------------------------------------------------------------------------------
static size_t update_ws_floats(int kl, int kr, int64_t M, int64_t N) {
  const size_t MN = (size_t)M * N;
  const bool dd = kl == PSGD_FACTOR_DENSE && kr == PSGD_FACTOR_DENSE;     // no rescaled factor copies (carve_update)
  size_t f = 64 + (dd ? 0 : fsize(kl, M) + fsize(kr, N)) + 4 * MN + (kl == PSGD_FACTOR_DENSE ? (size_t)M * M : 0) +
             (kr == PSGD_FACTOR_DENSE ? (size_t)N * N : 0) + 8 * (size_t)(M + N) +
             2 * col_partial_floats((int)M, (int)N) + 2 * MN /* mirrored transposes */ + 64 * 64;
  f += tc::trsm_scratch_floats((int)(M > N ? M : N));
  if (kl == PSGD_FACTOR_NORM && kr == PSGD_FACTOR_SCALE) f += ks::ns_update_scratch_floats((int)M, (int)N) + 64;
  if (kl == PSGD_FACTOR_DENSE && kr == PSGD_FACTOR_NORM) f += 2 * col_partial_floats((int)M, (int)N) + 64;
  return f;
}

static void carve_flags(const psgd_ctx* ctx, WsCarver& c, Layer& L, int kl, int kr, int M, int N) {
  int* f = c.take<int>(2);
  L.fl = wants_scan(ctx, kl, M) ? f : nullptr;
  L.fr = wants_scan(ctx, kr, N) ? f + 1 : nullptr;
}

static void carve_update(const psgd_ctx* ctx, WsCarver& c, Layer& L, int kl, int kr, int M, int N) {
  const size_t MN = (size_t)M * N;
  L.sc = c.take<Scal>(1);
  carve_flags(ctx, c, L, kl, kr, M, N);
  if (kl == PSGD_FACTOR_DENSE && kr == PSGD_FACTOR_DENSE) {
    // (dense, dense): no rescaled copies, rho rides in the epilogue of the last products (see balance_many_kernel)
    L.Qlb = const_cast<float*>(L.Ql);
    L.Qrb = const_cast<float*>(L.Qr);
  } else {
    L.Qlb = c.take<float>(fsize(kl, M));
    L.Qrb = c.take<float>(fsize(kr, N));
  }
  L.A = c.take<float>(MN);
  L.Bt = c.take<float>(MN);
  L.T1 = c.take<float>(MN);
  L.cvec = c.take<float>(N);
  L.part = c.take<float>(col_partial_floats(M, N));
  L.grad1 = kl == PSGD_FACTOR_DENSE ? c.take<float>((size_t)M * M) : nullptr;
  L.grad2 = kr == PSGD_FACTOR_DENSE ? c.take<float>((size_t)N * N) : nullptr;
  L.g1d = c.take<float>(M);
  L.g1b = c.take<float>(M);
  L.gvl = c.take<float>(M);
  L.sa = c.take<float>(N);
  L.sb = c.take<float>(N);
  L.gvec = c.take<float>(N);
  L.zinv = c.take<float>(tc::trsm_scratch_floats(M > N ? M : N));
  L.xwork = c.take<float>(MN);
  L.nspart = (kl == PSGD_FACTOR_NORM && kr == PSGD_FACTOR_SCALE) ? c.take<float>(ks::ns_update_scratch_floats(M, N)) : nullptr;
  L.nrpart = (kl == PSGD_FACTOR_DENSE && kr == PSGD_FACTOR_NORM) ? c.take<float>(2 * col_partial_floats(M, N)) : nullptr;
}

// grad2 = triu(A^T A - Bt^T Bt), Qr' = Qr - step2 grad2 Qr (times rho for (dense, dense) pairs)      psgd.py:176-179
static int update_right_dense(psgd_ctx* ctx, std::vector<Layer>& Ls, int M, int N, float step, float tiny, bool dd) {
  std::vector<la::Gemm> gs;
  for (auto& L : Ls) {
    la::Gemm g = mk(N, N, M, L.A, N, true, L.A, N, false, L.grad2, N);
    g.K2 = M; g.A2 = L.Bt; g.lda2 = N; g.ta2 = true; g.B2 = L.Bt; g.ldb2 = N; g.tb2 = false;
    g.triu = true; g.maxabs = &L.sc->max2;
    gs.push_back(g);
  }
  PSGD_RETURN_IF(gemm_all(ctx, gs));
  gs.clear();
  for (auto& L : Ls) {
    la::Gemm g = mk(N, N, N, L.grad2, N, false, L.Qrb, N, false, L.Qr_out, N);
    g.D = L.Qrb; g.ldd = N; g.mu_max = &L.sc->max2; g.step = step; g.tiny = tiny;
    if (dd) { g.rho = &L.sc->rho; g.rho_mode = 2; }                      // rho * Qr                      psgd.py:170
    g.d_tri = true;
    gs.push_back(g);
  }
  return gemm_all(ctx, gs, kUpper, kUpper, &Ls, 0, kFromR);
}

// A small (dense, dense) update is a chain of 13 latency-bound launches, but only half of them depend on each other:
//   A = Ql dG Qr^T  ||  Bt = Ql^-T dX Qr^-1,   then   (grad1, Ql')  ||  (grad2, Qr').
// Below the tensor-core sizes the two halves run on two streams (the slot's own and its `branch`), forked and joined
// with events, so the call stays CUDA-graph capturable and the halves become parallel branches of the graph.
struct Branch {
  psgd_ctx* ctx;
  cudaStream_t main, br;
  int slot;
  bool on;
  int fork() {          // the branch may start once everything enqueued on main so far is done
    if (!on) return PSGD_OK;
    PSGD_CUDA_CHECK(cudaEventRecord(ctx->ev_branch_go[slot], main));
    PSGD_CUDA_CHECK(cudaStreamWaitEvent(br, ctx->ev_branch_go[slot], 0));
    return PSGD_OK;
  }
  int join() {          // main continues once the branch is done
    if (!on) return PSGD_OK;
    PSGD_CUDA_CHECK(cudaEventRecord(ctx->ev_branch_done[slot], br));
    PSGD_CUDA_CHECK(cudaStreamWaitEvent(main, ctx->ev_branch_done[slot], 0));
    return PSGD_OK;
  }
  void to_branch() { if (on) ctx->stream = br; }
  void to_main() { ctx->stream = main; }
};

static int make_branch(psgd_ctx* ctx, int kl, int kr, int M, int N, Branch* b) {
  b->ctx = ctx; b->main = ctx->stream; b->br = ctx->stream; b->slot = ctx->stream_slot(); b->on = false;
  const bool dd = kl == PSGD_FACTOR_DENSE && kr == PSGD_FACTOR_DENSE;
  const bool ds = kl == PSGD_FACTOR_DENSE && kr == PSGD_FACTOR_SCALE;
  const bool tc_sized = M >= 256 && N >= 256 && (M % 4) == 0 && (N % 4) == 0;
  if (!ctx->opt_kron_streams || ctx->opt_profile || ctx->opt_gemm_path == 2) return PSGD_OK;
  // (dense, dense): both halves on the SIMT engine.  (dense, scaling) -- the NMT embeddings, canonical [256, 9416]:
  // the products may be tensor-core launches, but the branch only ever runs the SIMT panel solve (n < 512) and
  // streaming kernels next to them, so the GEMM engine's per-slot scratch is used by one stream at a time.
  const bool sd = kl == PSGD_FACTOR_SCALE && kr == PSGD_FACTOR_DENSE;
  if (!(dd && M < 512 && N < 512 && !tc_sized) && !(ds && M < 512) && !(sd && N < 512)) return PSGD_OK;
  const int s = b->slot;
  if (!ctx->branch[s]) {
    PSGD_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->branch[s], cudaStreamNonBlocking));
    PSGD_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_branch_go[s], cudaEventDisableTiming));
    PSGD_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_branch_done[s], cudaEventDisableTiming));
  }
  b->br = ctx->branch[s];
  b->on = true;
  return PSGD_OK;
}

static int update_group(psgd_ctx* ctx, int kl, int kr, std::vector<Layer>& Ls, int M, int N, float step, float tiny) {
  const size_t MN = (size_t)M * N;
  cudaStream_t st = ctx->stream;
  std::vector<la::Gemm> gs;
  std::vector<tc::Trsm> ts;
  const int64_t cl = (int64_t)fsize(kl, M), cr = (int64_t)fsize(kr, N);

  // ---- balance: Ql /= rho, Qr *= rho                          psgd.py:166-170, :211-215, :288-292, :342-346
  const bool small_pair = kl != PSGD_FACTOR_DENSE && kr != PSGD_FACTOR_DENSE && cl + cr <= kBalanceSmallMax;
  const bool dd = kl == PSGD_FACTOR_DENSE && kr == PSGD_FACTOR_DENSE;
  const bool sd = kl == PSGD_FACTOR_SCALE && kr == PSGD_FACTOR_DENSE;    // run in its own orientation, no transposes
  const bool dn = kl == PSGD_FACTOR_DENSE && kr == PSGD_FACTOR_NORM;     // likewise
  for (size_t t0 = 0; dd && t0 < Ls.size(); t0 += kPrepBatch) {          // (dense, dense): ONE launch per group
    const int cnt = (int)std::min<size_t>(kPrepBatch, Ls.size() - t0);
    PrepBatch b{};
    for (int t = 0; t < cnt; ++t) { b.Ql[t] = Ls[t0 + t].Ql; b.Qr[t] = Ls[t0 + t].Qr; b.sc[t] = Ls[t0 + t].sc; }
    balance_many_kernel<<<cnt, 256, 0, st>>>(b, kl, M, kr, N);
    PSGD_LAUNCH_CHECK(ctx);
  }
  for (auto& L : Ls) {
    if (dd) break;
    if (small_pair) {
      balance_rescale_small_kernel<<<1, 1024, 0, st>>>(kl, L.Ql, M, kr, L.Qr, N, cl, cr, L.Qlb, L.Qrb, L.sc);
      PSGD_LAUNCH_CHECK(ctx);
      continue;
    }
    if (dn) {
      // the transposed problem has the normalization factor first: rho = sqrt(max q0 / max diag(Ql)), q / rho, rho Ql
      balance_kernel<<<1, 256, 0, st>>>(kr, L.Qr, N, kl, L.Ql, M, L.sc);                   // psgd.py:86, :211-215
      PSGD_LAUNCH_CHECK(ctx);
      rescale_kernel<<<ew_grid(ctx, cr, 256), 256, 0, st>>>(L.Qr, L.Qrb, cr, L.sc, 1);
      PSGD_LAUNCH_CHECK(ctx);
      rescale_kernel<<<ew_grid(ctx, cl, 256), 256, 0, st>>>(L.Ql, L.Qlb, cl, L.sc, 0);
      PSGD_LAUNCH_CHECK(ctx);
      continue;
    }
    if (sd) {
      // the reference balances the transposed problem (dense factor first): rho = sqrt(max diag(Qr) / max ql),
      // Qr / rho, ql * rho                                                                  psgd.py:102-104, :288-292
      balance_kernel<<<1, 256, 0, st>>>(kr, L.Qr, N, kl, L.Ql, M, L.sc);
      PSGD_LAUNCH_CHECK(ctx);
      rescale_kernel<<<ew_grid(ctx, cl, 256), 256, 0, st>>>(L.Ql, L.Qlb, cl, L.sc, 0);
      PSGD_LAUNCH_CHECK(ctx);
      rescale_kernel<<<ew_grid(ctx, cr, 256), 256, 0, st>>>(L.Qr, L.Qrb, cr, L.sc, 1);
      PSGD_LAUNCH_CHECK(ctx);
      continue;
    }
    balance_kernel<<<1, 256, 0, st>>>(kl, L.Ql, M, kr, L.Qr, N, L.sc);
    PSGD_LAUNCH_CHECK(ctx);
    rescale_kernel<<<ew_grid(ctx, cl, 256), 256, 0, st>>>(L.Ql, L.Qlb, cl, L.sc, 1);
    PSGD_LAUNCH_CHECK(ctx);
    rescale_kernel<<<ew_grid(ctx, cr, 256), 256, 0, st>>>(L.Qr, L.Qrb, cr, L.sc, 0);
    PSGD_LAUNCH_CHECK(ctx);
  }
  PSGD_RETURN_IF(scan_group(ctx, Ls, M, N));

  // ---- A = Ql dG Qr^T  and  Bt = Ql^-T dX Qr^-1 ------------------------------------------------
  Branch br;
  PSGD_RETURN_IF(make_branch(ctx, kl, kr, M, N, &br));
  if (kl == PSGD_FACTOR_DENSE && kr == PSGD_FACTOR_DENSE) {
    PSGD_RETURN_IF(br.fork());
    gs.clear();                                                          // T1 = dG Qr^T            psgd.py:173
    for (auto& L : Ls) gs.push_back(mk(M, N, N, L.dG, N, false, L.Qrb, N, true, L.T1, N));
    PSGD_RETURN_IF(gemm_all(ctx, gs, 0, kLower, &Ls, 0, kFromR));
    gs.clear();                                                          // A = Ql T1
    for (auto& L : Ls) gs.push_back(mk(M, N, M, L.Qlb, M, false, L.T1, N, false, L.A, N));
    PSGD_RETURN_IF(gemm_all(ctx, gs, kUpper, 0, &Ls, kFromL, 0));
    br.to_branch();
    // forked: W goes where Bt will be and the left solve runs in place (the panel solves allow X == B), because T1 is
    // busy on the other stream
    ts.clear();                                                          // W = dX Qr^-1            psgd.py:174
    for (auto& L : Ls) ts.push_back(tc::Trsm{L.Qrb, L.dX, br.on ? L.Bt : L.T1, L.zinv, L.xwork});
    int rc = tc::trsm_right_many(ctx, ts.data(), (int)ts.size(), N, N, N, M, N);
    ts.clear();                                                          // Bt = Ql^-T W
    for (auto& L : Ls) ts.push_back(tc::Trsm{L.Qlb, br.on ? L.Bt : L.T1, L.Bt, L.zinv, L.xwork});
    if (rc == PSGD_OK) rc = tc::trsm_left_many(ctx, ts.data(), (int)ts.size(), M, N, N, M, N);
    br.to_main();
    PSGD_RETURN_IF(rc);
    PSGD_RETURN_IF(br.join());
  } else if (dn) {
    // (dense, normalization), X [M, N] with the normalization factor on the right; the reference's (normalization, dense)
    // update of the transposed problem (psgd.py:86 -> :198-246) written out for X itself
    // (tests/test_kernel_formulations.py):  A = Ql (dG Qr^T),  Bt = Ql^-T (dX Qr^-1),  grad(Ql) = triu(A A^T - Bt Bt^T),
    // and the normalization factor's gradient from COLUMN statistics of A, Bt.
    const int rows_grid = (M + 7) / 8 < ctx->num_sms * 8 ? (M + 7) / 8 : ctx->num_sms * 8;
    for (auto& L : Ls) {
      norm_right_kernel<<<rows_grid, 256, 0, st>>>(0, L.Qrb, L.dG, L.T1, M, N);             // T = dG Qr^T
      PSGD_LAUNCH_CHECK(ctx);
    }
    gs.clear();                                                          // A = Ql T
    for (auto& L : Ls) gs.push_back(mk(M, N, M, L.Qlb, M, false, L.T1, N, false, L.A, N));
    PSGD_RETURN_IF(gemm_all(ctx, gs, kUpper, 0, &Ls, kFromL, 0));
    for (auto& L : Ls) {
      norm_right_kernel<<<rows_grid, 256, 0, st>>>(1, L.Qrb, L.dX, L.T1, M, N);             // S = dX Qr^-1
      PSGD_LAUNCH_CHECK(ctx);
    }
    ts.clear();                                                          // Bt = Ql^-T S
    for (auto& L : Ls) ts.push_back(tc::Trsm{L.Qlb, L.T1, L.Bt, L.zinv, L.xwork});
    PSGD_RETURN_IF(tc::trsm_left_many(ctx, ts.data(), (int)ts.size(), M, N, N, M, N));
    const int chunks = (M + kColChunkRows - 1) / kColChunkRows;
    for (auto& L : Ls) {                                                 // q' from the column statistics    :235-241
      norm_right_stats_kernel<<<dim3((N + 127) / 128, chunks), 128, 0, st>>>(L.A, L.Bt, M, N, L.nrpart);
      PSGD_LAUNCH_CHECK(ctx);
      norm_right_finish_kernel<<<ew_grid(ctx, N, 128), 128, 0, st>>>(L.nrpart, chunks, N, L.sa, L.sb, &L.sc->max2);
      PSGD_LAUNCH_CHECK(ctx);
      norm_new_ql_kernel<<<ew_grid(ctx, N, 256), 256, 0, st>>>(L.Qrb, L.sa, L.sb, L.Qr_out, N, step, tiny, &L.sc->max2);
      PSGD_LAUNCH_CHECK(ctx);
    }
    gs.clear();                                                          // grad1 = triu(A A^T - Bt Bt^T)     :243
    for (auto& L : Ls) {
      la::Gemm g = mk(M, M, N, L.A, N, false, L.A, N, true, L.grad1, M);
      g.K2 = N; g.A2 = L.Bt; g.lda2 = N; g.ta2 = false; g.B2 = L.Bt; g.ldb2 = N; g.tb2 = true;
      g.triu = true; g.maxabs = &L.sc->max1;
      gs.push_back(g);
    }
    PSGD_RETURN_IF(gemm_all(ctx, gs));
    gs.clear();                                                          // Ql' = Ql - step grad1 Ql          :244-246
    for (auto& L : Ls) {
      la::Gemm g = mk(M, M, M, L.grad1, M, false, L.Qlb, M, false, L.Ql_out, M);
      g.D = L.Qlb; g.ldd = M; g.mu_max = &L.sc->max1; g.step = step; g.tiny = tiny;
      g.d_tri = true;
      gs.push_back(g);
    }
    return gemm_all(ctx, gs, kUpper, kUpper, &Ls, 0, kFromL);
  } else if (sd) {
    // (scaling, dense), X [M, N] with the dense factor on the right.  The reference computes the (dense, scaling) update of
    // the transposed problem (psgd.py:102-104 -> :288-307); written out for X itself that is
    //   A = diag(ql) (dG Qr^T),  Bt = diag(1/ql) (dX Qr^-1),  grad(Qr) = triu(A^T A - Bt^T Bt),  grad(ql)_i = |A_i|^2 - |Bt_i|^2
    // -- the right-factor products of the (dense, dense) pair and one streaming pass, no transposed copies of dX, dG.
    PSGD_RETURN_IF(br.fork());
    gs.clear();                                                          // T = dG Qr^T -> A
    for (auto& L : Ls) gs.push_back(mk(M, N, N, L.dG, N, false, L.Qrb, N, true, L.A, N));
    PSGD_RETURN_IF(gemm_all(ctx, gs, 0, kLower, &Ls, 0, kFromR));
    br.to_branch();
    ts.clear();                                                          // W = dX Qr^-1 -> Bt
    for (auto& L : Ls) ts.push_back(tc::Trsm{L.Qrb, L.dX, L.Bt, L.zinv, L.xwork});
    const int rcw = tc::trsm_right_many(ctx, ts.data(), (int)ts.size(), N, N, N, M, N);
    br.to_main();
    PSGD_RETURN_IF(rcw);
    PSGD_RETURN_IF(br.join());
    const int rows_grid = (M + 7) / 8 < ctx->num_sms * 8 ? (M + 7) / 8 : ctx->num_sms * 8;
    for (auto& L : Ls) {
      scale_rows_stats_kernel<<<rows_grid, 256, 0, st>>>(L.A, L.Bt, L.Qlb, M, N, L.g1d, L.g1b);
      PSGD_LAUNCH_CHECK(ctx);
    }
    PSGD_RETURN_IF(br.fork());
    for (auto& L : Ls) {                                                 // ql' = ql - step grad ql   psgd.py:304-307
      scale_grad_kernel<<<ew_grid(ctx, M, 128), 128, 0, st>>>(L.g1d, L.g1b, M, L.gvl, &L.sc->max1);
      PSGD_LAUNCH_CHECK(ctx);
      scale_new_qr_kernel<<<ew_grid(ctx, M, 128), 128, 0, st>>>(L.Qlb, L.gvl, L.Ql_out, M, step, tiny, &L.sc->max1);
      PSGD_LAUNCH_CHECK(ctx);
    }
    br.to_branch();                                                      // Qr' = Qr - step triu(A^T A - Bt^T Bt) Qr
    const int rcr = update_right_dense(ctx, Ls, M, N, step, tiny, false);
    br.to_main();
    PSGD_RETURN_IF(rcr);
    return br.join();
  } else if (kl == PSGD_FACTOR_NORM && kr == PSGD_FACTOR_DENSE) {
    for (auto& L : Ls) {
      norm_left_mul_kernel<<<ew_grid(ctx, MN, 256), 256, 0, st>>>(L.Qlb, L.dG, L.T1, M, N, nullptr, 0);   // :218-219
      PSGD_LAUNCH_CHECK(ctx);
    }
    gs.clear();                                                          // A = (Ql dG) Qr^T        psgd.py:220
    for (auto& L : Ls) gs.push_back(mk(M, N, N, L.T1, N, false, L.Qrb, N, true, L.A, N));
    PSGD_RETURN_IF(gemm_all(ctx, gs, 0, kLower, &Ls, 0, kFromR));
    for (auto& L : Ls) {
      PSGD_RETURN_IF(ks::col_wsum(ctx, 0, L.Qlb, nullptr, L.dX, N, M, N, L.part, L.cvec));
      norm_left_solve_kernel<<<ew_grid(ctx, MN, 256), 256, 0, st>>>(L.Qlb, L.dX, L.cvec, L.T1, M, N, nullptr);   // :230-232
      PSGD_LAUNCH_CHECK(ctx);
    }
    ts.clear();                                                          // Bt = (.) Qr^-1          psgd.py:233
    for (auto& L : Ls) ts.push_back(tc::Trsm{L.Qrb, L.T1, L.Bt, L.zinv, L.xwork});
    PSGD_RETURN_IF(tc::trsm_right_many(ctx, ts.data(), (int)ts.size(), N, N, N, M, N));
  } else if (kl == PSGD_FACTOR_DENSE && kr == PSGD_FACTOR_SCALE) {
    PSGD_RETURN_IF(br.fork());
    gs.clear();                                                          // A = (Ql dG) * qr        psgd.py:295-296
    for (auto& L : Ls) {
      la::Gemm g = mk(M, N, M, L.Qlb, M, false, L.dG, N, false, L.A, N);
      g.colscale = L.Qrb;
      gs.push_back(g);
    }
    PSGD_RETURN_IF(gemm_all(ctx, gs, kUpper, 0, &Ls, kFromL, 0));
    br.to_branch();
    auto bt_chain = [&]() -> int {
      ts.clear();                                                        // Bt = Ql^-T dX           psgd.py:298
      for (auto& L : Ls) ts.push_back(tc::Trsm{L.Qlb, L.dX, L.Bt, L.zinv, L.xwork});
      PSGD_RETURN_IF(tc::trsm_left_many(ctx, ts.data(), (int)ts.size(), M, N, N, M, N));
      for (auto& L : Ls) {
        col_scale_recip_kernel<<<ew_grid(ctx, MN, 256), 256, 0, ctx->stream>>>(L.Bt, L.Qrb, M, N);       // :299
        PSGD_LAUNCH_CHECK(ctx);
      }
      return PSGD_OK;
    };
    const int rcb = bt_chain();
    br.to_main();
    PSGD_RETURN_IF(rcb);
    PSGD_RETURN_IF(br.join());
  } else {  // (NORM, SCALE): no dense factor at all -- A and Bt are never materialised (kron_stream.cu)
    for (auto& L : Ls) {
      int cchunks = 0;
      PSGD_RETURN_IF(ks::col_wsum_partials(ctx, 0, L.Qlb, nullptr, L.dX, N, M, N, L.part, &cchunks));  // psgd.py:353-355
      PSGD_RETURN_IF(ks::ns_update_stats(ctx, L.Qlb, L.Qrb, L.part, cchunks, L.dX, L.dG, M, N, L.nspart, L.g1d, L.g1b,
                                         L.gvec, &L.sc->max1, &L.sc->max2, L.Ql_out, L.Qr_out, step, tiny));   // :349-369
      if (ks::ns_finish_is_fused(M, N)) continue;
      norm_new_ql_kernel<<<ew_grid(ctx, M, 256), 256, 0, st>>>(L.Qlb, L.g1d, L.g1b, L.Ql_out, M, step, tiny, &L.sc->max1);   // :362-364
      PSGD_LAUNCH_CHECK(ctx);
      scale_new_qr_kernel<<<ew_grid(ctx, N, 128), 128, 0, st>>>(L.Qrb, L.gvec, L.Qr_out, N, step, tiny, &L.sc->max2);        // :367-369
      PSGD_LAUNCH_CHECK(ctx);
    }
    return PSGD_OK;
  }

  // ---- left factor (main stream) and right factor (branch, when forked) -----------------------------
  PSGD_RETURN_IF(br.fork());
  if (kl == PSGD_FACTOR_DENSE) {
    gs.clear();                                                          // grad1 = triu(A A^T - Bt Bt^T)   psgd.py:175
    for (auto& L : Ls) {
      la::Gemm g = mk(M, M, N, L.A, N, false, L.A, N, true, L.grad1, M);
      g.K2 = N; g.A2 = L.Bt; g.lda2 = N; g.ta2 = false; g.B2 = L.Bt; g.ldb2 = N; g.tb2 = true;
      g.triu = true; g.maxabs = &L.sc->max1;
      gs.push_back(g);
    }
    PSGD_RETURN_IF(gemm_all(ctx, gs));
    gs.clear();                                                          // Ql' = Ql - step1 grad1 Ql       psgd.py:177, :179
    for (auto& L : Ls) {
      la::Gemm g = mk(M, M, M, L.grad1, M, false, L.Qlb, M, false, L.Ql_out, M);
      g.D = L.Qlb; g.ldd = M; g.mu_max = &L.sc->max1; g.step = step; g.tiny = tiny;
      if (dd) { g.rho = &L.sc->rho; g.rho_mode = 1; }                    // Ql / rho                      psgd.py:169
      g.d_tri = true;
      gs.push_back(g);
    }
    PSGD_RETURN_IF(gemm_all(ctx, gs, kUpper, kUpper, &Ls, 0, kFromL));
  } else {
    const int rows_grid = M < ctx->num_sms * 8 ? M : ctx->num_sms * 8;
    for (auto& L : Ls) {
      row_stats_kernel<<<rows_grid, 128, 0, st>>>(L.A, L.Bt, M, N, L.g1d, L.g1b, L.sc);                  // :235-239
      PSGD_LAUNCH_CHECK(ctx);
      norm_new_ql_kernel<<<ew_grid(ctx, M, 256), 256, 0, st>>>(L.Qlb, L.g1d, L.g1b, L.Ql_out, M, step, tiny, &L.sc->max1);   // :240-241
      PSGD_LAUNCH_CHECK(ctx);
    }
  }
  // ---- right factor (on the branch stream when forked) -------------------------------------------
  br.to_branch();
  int rc = PSGD_OK;
  if (kr == PSGD_FACTOR_DENSE) {
    rc = update_right_dense(ctx, Ls, M, N, step, tiny, dd);
  } else {
    auto scale_right = [&]() -> int {
      for (auto& L : Ls) {
        PSGD_RETURN_IF(col_reduce(ctx, 2, nullptr, L.A, L.Bt, M, N, L.part, L.sa, L.sb));                // :304 / :366
        scale_grad_kernel<<<ew_grid(ctx, N, 128), 128, 0, ctx->stream>>>(L.sa, L.sb, N, L.gvec, &L.sc->max2);
        PSGD_LAUNCH_CHECK(ctx);
        scale_new_qr_kernel<<<ew_grid(ctx, N, 128), 128, 0, ctx->stream>>>(L.Qrb, L.gvec, L.Qr_out, N, step, tiny, &L.sc->max2);   // :307
        PSGD_LAUNCH_CHECK(ctx);
      }
      return PSGD_OK;
    };
    rc = scale_right();
  }
  br.to_main();
  PSGD_RETURN_IF(rc);
  return br.join();
}

// ---------------------------------------------------------------------------------------------
// canonical apply of a group of same-shape layers
// ---------------------------------------------------------------------------------------------
static size_t apply_ws_floats(int64_t M, int64_t N) {
  const size_t MN = (size_t)M * N;
  return 64 * 64 + 5 * MN + (size_t)M * M + (size_t)N * N + 4 * (size_t)(M + N) + col_partial_floats((int)M, (int)N);
}

static void carve_apply(const psgd_ctx* ctx, WsCarver& c, Layer& L, int kl, int kr, int M, int N) {
  const size_t MN = (size_t)M * N;
  carve_flags(ctx, c, L, kl, kr, M, N);
  L.t1 = c.take<float>(MN);
  L.t2 = c.take<float>(MN);
  L.t3 = c.take<float>(MN);
  const size_t side = (kl == PSGD_FACTOR_DENSE && kr == PSGD_FACTOR_DENSE) ? (size_t)(M < N ? M : N) * (M < N ? M : N)
                      : (kl == PSGD_FACTOR_DENSE ? (size_t)M * M : (kr == PSGD_FACTOR_DENSE ? (size_t)N * N : 0));
  L.P = c.take<float>(side ? side : 1);
  L.addlast = c.take<float>(N);
  L.part = c.take<float>(col_partial_floats(M, N));
}

// The reference forms the small Gram matrix (Q^T Q) first whenever that saves dense flops (psgd.py:189, :260, :318).
// On the tensor-core path the factors' triangularity is exploited instead: X Q^T Q as two triangular products
// executes n^3 + n^3 multiply-adds, against 2/3 n^3 + 2 n^3 through the (dense, symmetric) Gram matrix, so large
// factors always take the chained association.  The two associations differ only in fp32 rounding (~1e-7 relative).
static bool chain_preferred(const psgd_ctx* ctx, int n, int other) {
  return ctx->opt_assume_tri && ctx->opt_gemm_path != 1 && n >= 512 && other >= 256 && (n % 4) == 0 && (other % 4) == 0;
}

// out = X (Qr^T Qr) with the reference's association switch          psgd.py:189-192, :260-263
static int right_dense_apply(psgd_ctx* ctx, std::vector<Layer>& Ls, int M, int N, bool from_t1, bool to_out) {
  std::vector<la::Gemm> gs;
  auto X = [&](Layer& L) { return from_t1 ? L.t1 : L.G; };
  auto O = [&](Layer& L) { return to_out ? L.out : L.t2; };
  if (M < N || chain_preferred(ctx, N, M)) {
    for (auto& L : Ls) gs.push_back(mk(M, N, N, X(L), N, false, L.Qr, N, true, L.t3, N));
    PSGD_RETURN_IF(gemm_all(ctx, gs, 0, kLower, &Ls, 0, kFromR));
    gs.clear();
    for (auto& L : Ls) gs.push_back(mk(M, N, N, L.t3, N, false, L.Qr, N, false, O(L), N));
    return gemm_all(ctx, gs, 0, kUpper, &Ls, 0, kFromR);
  }
  for (auto& L : Ls) gs.push_back(mk(N, N, N, L.Qr, N, true, L.Qr, N, false, L.P, N));
  PSGD_RETURN_IF(gemm_all(ctx, gs, kLower, kUpper, &Ls, kFromR, kFromR));
  gs.clear();
  for (auto& L : Ls) gs.push_back(mk(M, N, N, X(L), N, false, L.P, N, false, O(L), N));
  return gemm_all(ctx, gs);
}

static int apply_group(psgd_ctx* ctx, int kl, int kr, std::vector<Layer>& Ls, int M, int N) {
  const size_t MN = (size_t)M * N;
  cudaStream_t st = ctx->stream;
  std::vector<la::Gemm> gs;
  PSGD_RETURN_IF(scan_group(ctx, Ls, M, N));
  if (kl == PSGD_FACTOR_DENSE && kr == PSGD_FACTOR_DENSE) {
    if (M < N) {                                                          // psgd.py:190
      if (chain_preferred(ctx, M, N)) {                                   // Ql^T (Ql G) instead of (Ql^T Ql) G
        for (auto& L : Ls) gs.push_back(mk(M, N, M, L.Ql, M, false, L.G, N, false, L.t3, N));
        PSGD_RETURN_IF(gemm_all(ctx, gs, kUpper, 0, &Ls, kFromL, 0));
        gs.clear();
        for (auto& L : Ls) gs.push_back(mk(M, N, M, L.Ql, M, true, L.t3, N, false, L.t1, N));
        PSGD_RETURN_IF(gemm_all(ctx, gs, kLower, 0, &Ls, kFromL, 0));
      } else {
        for (auto& L : Ls) gs.push_back(mk(M, M, M, L.Ql, M, true, L.Ql, M, false, L.P, M));
        PSGD_RETURN_IF(gemm_all(ctx, gs, kLower, kUpper, &Ls, kFromL, kFromL));
        gs.clear();
        for (auto& L : Ls) gs.push_back(mk(M, N, M, L.P, M, false, L.G, N, false, L.t1, N));
        PSGD_RETURN_IF(gemm_all(ctx, gs));
      }
      gs.clear();
      for (auto& L : Ls) gs.push_back(mk(M, N, N, L.t1, N, false, L.Qr, N, true, L.t2, N));
      PSGD_RETURN_IF(gemm_all(ctx, gs, 0, kLower, &Ls, 0, kFromR));
      gs.clear();
      for (auto& L : Ls) gs.push_back(mk(M, N, N, L.t2, N, false, L.Qr, N, false, L.out, N));
      return gemm_all(ctx, gs, 0, kUpper, &Ls, 0, kFromR);
    }
    if (chain_preferred(ctx, N, M)) {                                     // (G Qr^T) Qr instead of G (Qr^T Qr)
      for (auto& L : Ls) gs.push_back(mk(M, N, N, L.G, N, false, L.Qr, N, true, L.t3, N));
      PSGD_RETURN_IF(gemm_all(ctx, gs, 0, kLower, &Ls, 0, kFromR));
      gs.clear();
      for (auto& L : Ls) gs.push_back(mk(M, N, N, L.t3, N, false, L.Qr, N, false, L.t1, N));
      PSGD_RETURN_IF(gemm_all(ctx, gs, 0, kUpper, &Ls, 0, kFromR));
    } else {
      for (auto& L : Ls) gs.push_back(mk(N, N, N, L.Qr, N, true, L.Qr, N, false, L.P, N));     // psgd.py:192
      PSGD_RETURN_IF(gemm_all(ctx, gs, kLower, kUpper, &Ls, kFromR, kFromR));
      gs.clear();
      for (auto& L : Ls) gs.push_back(mk(M, N, N, L.G, N, false, L.P, N, false, L.t1, N));
      PSGD_RETURN_IF(gemm_all(ctx, gs));
    }
    gs.clear();
    for (auto& L : Ls) gs.push_back(mk(M, N, M, L.Ql, M, false, L.t1, N, false, L.t2, N));
    PSGD_RETURN_IF(gemm_all(ctx, gs, kUpper, 0, &Ls, kFromL, 0));
    gs.clear();
    for (auto& L : Ls) gs.push_back(mk(M, N, M, L.Ql, M, true, L.t2, N, false, L.out, N));
    return gemm_all(ctx, gs, kLower, 0, &Ls, kFromL, 0);
  }
  if (kl == PSGD_FACTOR_DENSE && kr == PSGD_FACTOR_NORM) {
    // the reference's (normalization, dense) apply of the transposed problem (psgd.py:128 -> :249-270) written for G itself:
    // P = G Qr^T, then Ql^T Ql P with the association the reference picks for P^T, then the transposed out-op
    const int rows_grid = (M + 7) / 8 < ctx->num_sms * 8 ? (M + 7) / 8 : ctx->num_sms * 8;
    for (auto& L : Ls) {
      norm_right_kernel<<<rows_grid, 256, 0, st>>>(0, L.Qr, L.G, L.t1, M, N);
      PSGD_LAUNCH_CHECK(ctx);
    }
    if (N < M || chain_preferred(ctx, M, N)) {                             // psgd.py:261 for the transposed problem
      for (auto& L : Ls) gs.push_back(mk(M, N, M, L.Ql, M, false, L.t1, N, false, L.t2, N));
      PSGD_RETURN_IF(gemm_all(ctx, gs, kUpper, 0, &Ls, kFromL, 0));
      gs.clear();
      for (auto& L : Ls) gs.push_back(mk(M, N, M, L.Ql, M, true, L.t2, N, false, L.t3, N));
      PSGD_RETURN_IF(gemm_all(ctx, gs, kLower, 0, &Ls, kFromL, 0));
    } else {                                                               // :263
      for (auto& L : Ls) gs.push_back(mk(M, M, M, L.Ql, M, true, L.Ql, M, false, L.P, M));
      PSGD_RETURN_IF(gemm_all(ctx, gs, kLower, kUpper, &Ls, kFromL, kFromL));
      gs.clear();
      for (auto& L : Ls) gs.push_back(mk(M, N, M, L.P, M, false, L.t1, N, false, L.t3, N));
      PSGD_RETURN_IF(gemm_all(ctx, gs));
    }
    for (auto& L : Ls) {
      norm_right_kernel<<<rows_grid, 256, 0, st>>>(2, L.Qr, L.t3, L.out, M, N);
      PSGD_LAUNCH_CHECK(ctx);
    }
    return PSGD_OK;
  }
  if (kl == PSGD_FACTOR_SCALE && kr == PSGD_FACTOR_DENSE) {
    // the reference's (dense, scaling) apply of the transposed problem (psgd.py:144-146 -> :318-322) written for G itself:
    // out = diag(ql^2) G Qr^T Qr, the small Gram matrix first when that is what the reference does for G^T (N < M)
    if (N < M && !chain_preferred(ctx, N, M)) {
      for (auto& L : Ls) gs.push_back(mk(N, N, N, L.Qr, N, true, L.Qr, N, false, L.P, N));
      PSGD_RETURN_IF(gemm_all(ctx, gs, kLower, kUpper, &Ls, kFromR, kFromR));
      gs.clear();
      for (auto& L : Ls) gs.push_back(mk(M, N, N, L.G, N, false, L.P, N, false, L.out, N));
      PSGD_RETURN_IF(gemm_all(ctx, gs));
    } else {
      for (auto& L : Ls) gs.push_back(mk(M, N, N, L.G, N, false, L.Qr, N, true, L.t1, N));
      PSGD_RETURN_IF(gemm_all(ctx, gs, 0, kLower, &Ls, 0, kFromR));
      gs.clear();
      for (auto& L : Ls) gs.push_back(mk(M, N, N, L.t1, N, false, L.Qr, N, false, L.out, N));
      PSGD_RETURN_IF(gemm_all(ctx, gs, 0, kUpper, &Ls, 0, kFromR));
    }
    for (auto& L : Ls) {
      row_scale_sq_kernel<<<ew_grid(ctx, MN, 256), 256, 0, st>>>(L.out, L.Ql, M, N);
      PSGD_LAUNCH_CHECK(ctx);
    }
    return PSGD_OK;
  }
  if (kl == PSGD_FACTOR_DENSE && kr == PSGD_FACTOR_SCALE) {               // psgd.py:318-322
    if (M < N && !chain_preferred(ctx, M, N)) {
      for (auto& L : Ls) gs.push_back(mk(M, M, M, L.Ql, M, true, L.Ql, M, false, L.P, M));
      PSGD_RETURN_IF(gemm_all(ctx, gs, kLower, kUpper, &Ls, kFromL, kFromL));
      gs.clear();
      for (auto& L : Ls) {
        la::Gemm g = mk(M, N, M, L.P, M, false, L.G, N, false, L.out, N);
        g.colscale = L.Qr; g.colscale_sq = true;
        gs.push_back(g);
      }
      return gemm_all(ctx, gs);
    }
    for (auto& L : Ls) gs.push_back(mk(M, N, M, L.Ql, M, false, L.G, N, false, L.t1, N));
    PSGD_RETURN_IF(gemm_all(ctx, gs, kUpper, 0, &Ls, kFromL, 0));
    gs.clear();
    for (auto& L : Ls) {
      la::Gemm g = mk(M, N, M, L.Ql, M, true, L.t1, N, false, L.out, N);
      g.colscale = L.Qr; g.colscale_sq = true;
      gs.push_back(g);
    }
    return gemm_all(ctx, gs, kLower, 0, &Ls, kFromL, 0);
  }
  // normalization-format left factor                                      psgd.py:258-270, :383-391
  if (kr == PSGD_FACTOR_SCALE) {                                           // one pass over G (kron_stream.cu)
    for (auto& L : Ls) PSGD_RETURN_IF(ks::ns_apply(ctx, L.Ql, L.Qr, L.G, L.out, M, N, L.part));
    return PSGD_OK;
  }
  for (auto& L : Ls) {
    norm_left_mul_kernel<<<ew_grid(ctx, MN, 256), 256, 0, st>>>(L.Ql, L.G, L.t1, M, N, nullptr, 0);
    PSGD_LAUNCH_CHECK(ctx);
  }
  PSGD_RETURN_IF(right_dense_apply(ctx, Ls, M, N, true, false));          // -> t2
  for (auto& L : Ls) {
    PSGD_RETURN_IF(ks::col_wsum(ctx, 1, nullptr, L.Ql + M, L.t2, N, M, N, L.part, L.addlast));   // psgd.py:265
    norm_left_out_kernel<<<ew_grid(ctx, MN, 256), 256, 0, st>>>(L.Ql, L.t2, L.addlast, L.out, M, N);
    PSGD_LAUNCH_CHECK(ctx);
  }
  return PSGD_OK;
}

// ---------------------------------------------------------------------------------------------
// dispatch with the reference's mirroring, grouping of same-shape layers     psgd.py:82-110, :124-152
// ---------------------------------------------------------------------------------------------
static bool is_canonical(int kl, int kr) {
  return (kl == PSGD_FACTOR_DENSE && kr == PSGD_FACTOR_DENSE) || (kl == PSGD_FACTOR_NORM && kr == PSGD_FACTOR_DENSE) ||
         (kl == PSGD_FACTOR_DENSE && kr == PSGD_FACTOR_SCALE) || (kl == PSGD_FACTOR_NORM && kr == PSGD_FACTOR_SCALE);
}
static bool is_mirrored(int kl, int kr) { return is_canonical(kr, kl) && !is_canonical(kl, kr); }

static int check_layer(const char* what, int kl, int kr, int64_t M, int64_t N) {
  PSGD_REQUIRE(M >= 1 && N >= 1 && M < (1LL << 30) && N < (1LL << 30), PSGD_ERR_BAD_SHAPE, "%s: bad shape [%lld,%lld]",
               what, (long long)M, (long long)N);
  PSGD_REQUIRE(kl >= 0 && kl <= 2 && kr >= 0 && kr <= 2, PSGD_ERR_BAD_SHAPE, "%s: bad factor kinds %d,%d", what, kl, kr);
  PSGD_REQUIRE(is_canonical(kl, kr) || is_mirrored(kl, kr), PSGD_ERR_UNSUPPORTED,
               "Unknown Kronecker product preconditioner (left kind %d, right kind %d)", kl, kr);
  return PSGD_OK;
}

struct Key { int kl, kr, M, N; };

static int ensure_side_streams(psgd_ctx* ctx) {
  if (ctx->ev_fork) return PSGD_OK;
  for (int k = 0; k < psgd_ctx::kSideStreams; ++k) {
    PSGD_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->side[k], cudaStreamNonBlocking));
    PSGD_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_side[k], cudaEventDisableTiming));
  }
  PSGD_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  return PSGD_OK;
}

// Runs update (is_update) or apply over a ragged list of layers.  The mirrored format (scale,norm) is brought to
// canonical orientation by transposing dX,dG / G into scratch and swapping the factors (exactly what the reference
// does for all three mirrored formats, psgd.py:86, :102, :104, :128, :144, :146); (scale,dense) and (dense,norm) run in
// their own orientation.
// Layers with equal (kinds, shape) then run as one group so every GEMM/TRSM of the op sequence is a single grouped launch.
static int run_layers(psgd_ctx* ctx, const psgd_kron_layer* in, int count, bool is_update, float step, float tiny) {
  if (count == 0) return PSGD_OK;
  size_t need = 0;
  for (int i = 0; i < count; ++i) {
    const psgd_kron_layer& q = in[i];
    PSGD_RETURN_IF(check_layer(is_update ? "kron update" : "kron apply", q.kind_l, q.kind_r, q.M, q.N));
    if (is_update)
      PSGD_REQUIRE(q.Ql && q.Qr && q.dX && q.dG && q.Ql_out && q.Qr_out, PSGD_ERR_BAD_POINTER,
                   "kron update: null device pointer in layer %d", i);
    else
      PSGD_REQUIRE(q.Ql && q.Qr && q.G && q.out, PSGD_ERR_BAD_POINTER, "kron apply: null device pointer in layer %d", i);
    const size_t f = is_update ? update_ws_floats(q.kind_l, q.kind_r, q.M, q.N)
                               : apply_ws_floats(q.M, q.N) + 2 * (size_t)q.M * q.N;
    need += f * sizeof(float) + 48 * 256;
  }
  // Bound the workspace: process the list in slices whose scratch fits the budget (large uniform stacks still group)
  const size_t budget = (size_t)48 << 30;
  int begin = 0;
  while (begin < count) {
    size_t bytes = 0;
    int end = begin;
    while (end < count) {
      const psgd_kron_layer& q = in[end];
      const size_t f = (is_update ? update_ws_floats(q.kind_l, q.kind_r, q.M, q.N)
                                  : apply_ws_floats(q.M, q.N) + 2 * (size_t)q.M * q.N) * sizeof(float) + 48 * 256;
      if (end > begin && bytes + f > budget) break;
      bytes += f;
      ++end;
    }
    PSGD_RETURN_IF(ctx->reserve(bytes));
    WsCarver c(ctx->ws);
    std::vector<Layer> Ls(end - begin);
    std::vector<Key> keys(end - begin);
    std::vector<float*> untranspose_src(end - begin, nullptr);
    for (int i = begin; i < end; ++i) {
      const psgd_kron_layer& q = in[i];
      Layer& L = Ls[i - begin];
      L = Layer{};
      int kl = q.kind_l, kr = q.kind_r, M = (int)q.M, N = (int)q.N;
      // (scaling, dense) -- the NMT embeddings [9414, 256] -- and (dense, normalization) have kernels for their own
      // orientation (update_group / apply_group); (scaling, normalization) goes through the canonical kernels on
      // transposed copies
      if (is_canonical(kl, kr) || (kl == PSGD_FACTOR_SCALE && kr == PSGD_FACTOR_DENSE) ||
          (kl == PSGD_FACTOR_DENSE && kr == PSGD_FACTOR_NORM)) {
        L.Ql = q.Ql; L.Qr = q.Qr; L.dX = q.dX; L.dG = q.dG; L.G = q.G;
        L.Ql_out = q.Ql_out; L.Qr_out = q.Qr_out; L.out = q.out;
      } else {
        // canonical kernel on (Qr, Ql, X^T); results come back swapped / transposed
        const size_t MN = (size_t)M * N;
        L.Ql = q.Qr; L.Qr = q.Ql; L.Ql_out = q.Qr_out; L.Qr_out = q.Ql_out;
        auto transposed = [&](const float* src, float** out) -> int {
          float* t = c.take<float>(MN);
          PSGD_RETURN_IF(la::transpose(ctx, src, N, t, M, M, N));
          *out = t;
          return PSGD_OK;
        };
        if (is_update) {
          float *dXt = nullptr, *dGt = nullptr;
          PSGD_RETURN_IF(transposed(q.dX, &dXt));
          PSGD_RETURN_IF(transposed(q.dG, &dGt));
          L.dX = dXt; L.dG = dGt;
        } else {
          float* Gt = nullptr;
          PSGD_RETURN_IF(transposed(q.G, &Gt));
          float* Ot = c.take<float>(MN);
          L.G = Gt; L.out = Ot;
          untranspose_src[i - begin] = Ot;
        }
        std::swap(kl, kr);
        std::swap(M, N);                                   // canonical shape
      }
      keys[i - begin] = Key{kl, kr, M, N};
      if (is_update) carve_update(ctx, c, L, kl, kr, M, N);
      else carve_apply(ctx, c, L, kl, kr, M, N);
    }
    // group equal keys (stable)
    struct Group { Key key; std::vector<Layer> layers; double cost; int slot; };
    std::vector<Group> groups;
    std::vector<char> done(end - begin, 0);
    for (int i = 0; i < end - begin; ++i) {
      if (done[i]) continue;
      Group g;
      g.key = keys[i];
      for (int j = i; j < end - begin; ++j)
        if (!done[j] && keys[j].kl == keys[i].kl && keys[j].kr == keys[i].kr && keys[j].M == keys[i].M && keys[j].N == keys[i].N) {
          g.layers.push_back(Ls[j]); done[j] = 1;
        }
      const double M = g.key.M, N = g.key.N;
      g.cost = g.layers.size() * (M * N * ((g.key.kl == PSGD_FACTOR_DENSE ? M : 8.0) + (g.key.kr == PSGD_FACTOR_DENSE ? N : 8.0)) + 2e5);
      g.slot = 0;
      groups.push_back(std::move(g));
    }
    // Groups are independent (layers share nothing, scratch is carved per layer): a ragged list (LeNet5: five shapes,
    // the NMT model: seven) runs its groups CONCURRENTLY on internal side streams, forked from and joined back into the
    // context's stream -- each group is a latency-bound chain of small launches that leaves most SMs idle.  Events
    // only, so the whole call stays capturable into a CUDA graph (where the groups become parallel branches).
    int nslots = 1;
    if (ctx->opt_kron_streams && groups.size() > 1) {
      nslots = (int)std::min<size_t>(groups.size(), psgd_ctx::kSideStreams + 1);
      std::vector<int> order(groups.size());
      for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return groups[a].cost > groups[b].cost; });
      std::vector<double> load(nslots, 0.0);
      for (int gi : order) {                      // longest-processing-time first
        int best = 0;
        for (int s = 1; s < nslots; ++s) if (load[s] < load[best]) best = s;
        groups[gi].slot = best; load[best] += groups[gi].cost;
      }
      PSGD_RETURN_IF(ensure_side_streams(ctx));
      PSGD_CUDA_CHECK(cudaEventRecord(ctx->ev_fork, ctx->stream));
      for (int s = 1; s < nslots; ++s) PSGD_CUDA_CHECK(cudaStreamWaitEvent(ctx->side[s - 1], ctx->ev_fork, 0));
    }
    cudaStream_t main_stream = ctx->stream;
    int status = PSGD_OK;
    for (auto& g : groups) {
      ctx->stream = g.slot == 0 ? main_stream : ctx->side[g.slot - 1];
      status = is_update ? update_group(ctx, g.key.kl, g.key.kr, g.layers, g.key.M, g.key.N, step, tiny)
                         : apply_group(ctx, g.key.kl, g.key.kr, g.layers, g.key.M, g.key.N);
      if (status != PSGD_OK) break;
    }
    ctx->stream = main_stream;
    for (int s = 1; s < nslots; ++s) {            // join (also after an error: never leave a forked stream dangling)
      cudaError_t e1 = cudaEventRecord(ctx->ev_side[s - 1], ctx->side[s - 1]);
      cudaError_t e2 = cudaStreamWaitEvent(main_stream, ctx->ev_side[s - 1], 0);
      if (status == PSGD_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) {
        set_error("kron: joining side stream %d failed: %s", s, cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        status = PSGD_ERR_CUDA;
      }
    }
    PSGD_RETURN_IF(status);
    if (!is_update)
      for (int i = begin; i < end; ++i)
        if (untranspose_src[i - begin])   // canonical result is [N,M]; give the caller [M,N]
          PSGD_RETURN_IF(la::transpose(ctx, untranspose_src[i - begin], (int)in[i].M, in[i].out, (int)in[i].N, (int)in[i].N, (int)in[i].M));
    begin = end;
  }
  (void)need;
  return PSGD_OK;
}

}  // namespace kron
}  // namespace psgd

using namespace psgd;

extern "C" int psgd_kron_update(psgd_ctx* ctx, int kind_l, int kind_r, const float* Ql, const float* Qr,
                                const float* dX, const float* dG, float* Ql_out, float* Qr_out, int64_t M, int64_t N,
                                float step, float tiny) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  psgd_kron_layer L{};
  L.kind_l = kind_l; L.kind_r = kind_r; L.M = M; L.N = N;
  L.Ql = Ql; L.Qr = Qr; L.dX = dX; L.dG = dG; L.Ql_out = Ql_out; L.Qr_out = Qr_out;
  return kron::run_layers(ctx, &L, 1, true, step, tiny);
}

extern "C" int psgd_kron_apply(psgd_ctx* ctx, int kind_l, int kind_r, const float* Ql, const float* Qr,
                               const float* G, float* out, int64_t M, int64_t N) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  psgd_kron_layer L{};
  L.kind_l = kind_l; L.kind_r = kind_r; L.M = M; L.N = N;
  L.Ql = Ql; L.Qr = Qr; L.G = G; L.out = out;
  return kron::run_layers(ctx, &L, 1, false, 0.f, 0.f);
}

// A ragged list of layers on one stream.  Layers are independent (mnist_with_lenet5.py:51-53); layers of equal format
// and shape are processed as a group in which every GEMM / triangular-solve step is ONE launch over the whole group.
extern "C" int psgd_kron_update_batched(psgd_ctx* ctx, const psgd_kron_layer* layers, int count, float step,
                                        float tiny) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_REQUIRE(count >= 0 && (layers || count == 0), PSGD_ERR_BAD_POINTER, "kron batched update: null layer list");
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  return kron::run_layers(ctx, layers, count, true, step, tiny);
}

extern "C" int psgd_kron_apply_batched(psgd_ctx* ctx, const psgd_kron_layer* layers, int count) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_REQUIRE(count >= 0 && (layers || count == 0), PSGD_ERR_BAD_POINTER, "kron batched apply: null layer list");
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  return kron::run_layers(ctx, layers, count, false, 0.f, 0.f);
}
