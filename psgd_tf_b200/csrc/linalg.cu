// SIMT fp32 engine of the device linear-algebra vocabulary (see linalg.cuh).
#include "linalg.cuh"

namespace psgd {
namespace la {

// ---------------------------------------------------------------------------------------------
// GEMM: T x T x 16 tiles (T = 64: 4x4 outputs per thread; T = 32: 2x2), 256 threads, fp32 FMA accumulation.
// The 32-wide tile is for problems whose 64-wide grid would leave most SMs idle (LeNet5 layers: 257 x 120 is 10 CTAs of
// 64 x 64, each a chain of K/16 barrier-separated steps -- 16 us per launch; 36 CTAs of 32 x 32 do a quarter of the work
// per step).  Same accumulation order per output element in both, so results do not depend on the tile choice.
// ---------------------------------------------------------------------------------------------
constexpr int BK = 16, PAD = 4;

// One K step of one product: this thread's elements of the A and B tiles, global -> registers (fetch) and registers ->
// shared memory (stash).  Kept apart so that the loads of step s+1 are in flight while step s is multiplied.
template <int T>
struct TileRegs {
  float a[(T * BK) / 256], b[(T * BK) / 256];
};
template <int T>
__device__ __forceinline__ void fetch_tiles(const float* __restrict__ A, int lda, bool ta, const float* __restrict__ B, int ldb,
                                            bool tb, int M, int N, int K, int m0, int n0, int k0, int tid, TileRegs<T>& r) {
#pragma unroll
  for (int e = 0; e < (T * BK) / 256; ++e) {
    const int idx = tid + 256 * e;
    int m, k;
    if (ta) { m = idx % T; k = idx / T; } else { k = idx % BK; m = idx / BK; }
    const int gm = m0 + m, gk = k0 + k;
    r.a[e] = (gm < M && gk < K) ? (ta ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk]) : 0.f;
  }
#pragma unroll
  for (int e = 0; e < (T * BK) / 256; ++e) {
    const int idx = tid + 256 * e;
    int n, k;
    if (tb) { k = idx % BK; n = idx / BK; } else { n = idx % T; k = idx / T; }
    const int gn = n0 + n, gk = k0 + k;
    r.b[e] = (gn < N && gk < K) ? (tb ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn]) : 0.f;
  }
}
template <int T>
__device__ __forceinline__ void stash_tiles(bool ta, bool tb, int tid, const TileRegs<T>& r, float (*As)[T + PAD],
                                            float (*Bs)[T + PAD]) {
#pragma unroll
  for (int e = 0; e < (T * BK) / 256; ++e) {
    const int idx = tid + 256 * e;
    int m, k;
    if (ta) { m = idx % T; k = idx / T; } else { k = idx % BK; m = idx / BK; }
    As[k][m] = r.a[e];
  }
#pragma unroll
  for (int e = 0; e < (T * BK) / 256; ++e) {
    const int idx = tid + 256 * e;
    int n, k;
    if (tb) { k = idx % BK; n = idx / BK; } else { n = idx % T; k = idx / T; }
    Bs[k][n] = r.b[e];
  }
}

template <int T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(Gemm g) {
  constexpr int R = T / 16;                         // outputs per thread along each dimension
  __shared__ __align__(16) float As[BK][T + PAD];
  __shared__ __align__(16) float Bs[BK][T + PAD];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * T, n0 = blockIdx.x * T;
  if (g.triu && m0 >= n0 + T) {
    // tile entirely below the diagonal: result is zero after masking
    for (int e = tid; e < T * T; e += 256) {
      const int m = m0 + e / T, n = n0 + e % T;
      if (m < g.M && n < g.N) g.C[(size_t)m * g.ldc + n] = 0.f;
    }
    return;
  }
  float acc[R][R];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < R; ++j) acc[i][j] = 0.f;

  // the K steps of product 0, then those of the (subtracted) product 1, as ONE software-pipelined sequence
  const int steps0 = (g.K > 0 && g.A) ? (g.K + BK - 1) / BK : 0;
  const int steps1 = (g.K2 > 0 && g.A2) ? (g.K2 + BK - 1) / BK : 0;
  const int steps = steps0 + steps1;
  auto fetch = [&](int s, TileRegs<T>& r) {
    const bool p1 = s >= steps0;
    const int k0 = (p1 ? s - steps0 : s) * BK;
    if (p1) fetch_tiles<T>(g.A2, g.lda2, g.ta2, g.B2, g.ldb2, g.tb2, g.M, g.N, g.K2, m0, n0, k0, tid, r);
    else fetch_tiles<T>(g.A, g.lda, g.ta, g.B, g.ldb, g.tb, g.M, g.N, g.K, m0, n0, k0, tid, r);
  };
  TileRegs<T> cur, nxt;
  if (steps > 0) {
    fetch(0, cur);
    stash_tiles<T>(steps0 > 0 ? g.ta : g.ta2, steps0 > 0 ? g.tb : g.tb2, tid, cur, As, Bs);
  }
  __syncthreads();
  for (int s = 0; s < steps; ++s) {
    const bool more = s + 1 < steps;
    if (more) fetch(s + 1, nxt);                       // in flight during the multiply below
    const float sign = s >= steps0 ? -1.f : 1.f;
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[R], b[R];
#pragma unroll
      for (int i = 0; i < R; ++i) { a[i] = sign * As[k][ty * R + i]; b[i] = Bs[k][tx * R + i]; }
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    if (more) {
      const bool p1 = s + 1 >= steps0;
      stash_tiles<T>(p1 ? g.ta2 : g.ta, p1 ? g.tb2 : g.tb, tid, nxt, As, Bs);
      __syncthreads();
    }
  }

  float mu = 0.f;
  if (g.D) mu = g.mu_max ? g.step / (*g.mu_max + g.tiny) : 1.0f;
  float mx = 0.f;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int m = m0 + ty * R + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int n = n0 + tx * R + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.colscale) {
        float s = g.colscale[n];
        if (g.colscale_sq) s = s * s;
        v = g.colscale_recip ? v * (1.0f / s) : v * s;
      }
      if (g.triu && m > n) v = 0.f;
      if (g.D) v = g.D[(size_t)m * g.ldd + n] - mu * v;
      if (g.rho_mode == 1) v = v / *g.rho;
      else if (g.rho_mode == 2) v = v * *g.rho;
      mx = fmaxf(mx, fabsf(v));
      g.C[(size_t)m * g.ldc + n] = v;
    }
  }
  if (g.maxabs) {
    mx = warp_max(mx);
    if ((tid & 31) == 0 && mx > 0.f) atomic_max_nonneg(g.maxabs, mx);
  }
}

int gemm_simt(psgd_ctx* ctx, const Gemm& g) {
  if (g.M <= 0 || g.N <= 0) return PSGD_OK;
  const int ctas64 = ((g.N + 63) / 64) * ((g.M + 63) / 64);
  if (ctas64 * 2 <= ctx->num_sms) {
    dim3 grid((g.N + 31) / 32, (g.M + 31) / 32);
    gemm_simt_kernel<32><<<grid, 256, 0, ctx->stream>>>(g);
  } else {
    dim3 grid((g.N + 63) / 64, (g.M + 63) / 64);
    gemm_simt_kernel<64><<<grid, 256, 0, ctx->stream>>>(g);
  }
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

// ---------------------------------------------------------------------------------------------
// triangular solves, left-looking, 32-row blocks: one launch per block step
// ---------------------------------------------------------------------------------------------
constexpr int NB = 32;

// Rows [ib0, ib1) of  X = Q^-T B  for one 32-column slab per CTA, assuming rows < ib0 of the right-hand side have
// already been eliminated (B holds B - Q[0:ib0, :]^T X[0:ib0, :] there).  Left-looking over 32-row steps:
//   X[I,:] = Qii^-T ( B[I,:] - sum_{ib0 <= k < i0} Q[k, I]^T X[k, :] )
// Columns are independent, so the whole row range is solved in one launch.
__global__ void __launch_bounds__(256) trsm_left_block_kernel(const float* __restrict__ Q, int ldq,
                                                              const float* B, int ldb, float* X, int ldx, int m,
                                                              int ib0, int ib1) {
  __shared__ float Qs[NB][NB + 1];   // Qs[k][i] = Q[k0+k, i0+i]
  __shared__ float Xs[NB][NB + 1];   // Xs[k][c] = X[k0+k, c0+c]
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * NB;
  const int lr = tid / NB;          // 0..7
  const int lc = tid % NB;          // 0..31
  for (int i0 = ib0; i0 < ib1; i0 += NB) {
    const int ib = min(NB, ib1 - i0);   // rows in this step
    float acc[4] = {0.f, 0.f, 0.f, 0.f};   // rows lr, lr+8, lr+16, lr+24 ; column lc
    // tiles of the next k0 step are fetched into registers while the current one is multiplied (these solves are
    // latency chains: one global round trip per 32 x 32 tile otherwise)
    float qn[4], xn[4];
    auto fetch = [&](int k0) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = lr + 8 * e;
        qn[e] = (i0 + lc < ib1) ? Q[(size_t)(k0 + k) * ldq + i0 + lc] : 0.f;
        xn[e] = (c0 + lc < m) ? X[(size_t)(k0 + k) * ldx + c0 + lc] : 0.f;
      }
    };
    if (ib0 < i0) fetch(ib0);
    for (int k0 = ib0; k0 < i0; k0 += NB) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = lr + 8 * e;
        Qs[k][lc] = qn[e];
        Xs[k][lc] = xn[e];
      }
      __syncthreads();
      if (k0 + NB < i0) fetch(k0 + NB);
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        const float x = Xs[k][lc];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = fmaf(Qs[k][lr + 8 * e], x, acc[e]);
      }
      __syncthreads();
    }
    // diagonal block: forward substitution with Qii^T (only the upper triangle of Qii is read)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = lr + 8 * e;
      Qs[i][lc] = (i < ib && lc < ib && i <= lc) ? Q[(size_t)(i0 + i) * ldq + i0 + lc] : 0.f;   // Qs[k][i], k<=i
      float b = 0.f;
      if (i < ib && c0 + lc < m) b = B[(size_t)(i0 + i) * ldb + c0 + lc] - acc[e];
      Xs[i][lc] = b;
    }
    __syncthreads();
    // Right-looking over the 32 rows with ALL 256 threads: row i is finished by the thread that owns it, one barrier,
    // then every thread folds x_i into the rows it owns.  Same operations in the same order as the row-by-row loop it
    // replaces (each b_j sees -q_kj x_k for k = 0, 1, ... with one fused rounding each), which one thread per column ran
    // as a serial chain of ~500 dependent shared-memory round trips per block (124 us per solve at n = 257, ncu).
    for (int i = 0; i < ib; ++i) {
      if (lr == (i & 7)) {
        const float x = Xs[i][lc] / Qs[i][i];
        Xs[i][lc] = x;
        if (c0 + lc < m) X[(size_t)(i0 + i) * ldx + c0 + lc] = x;
      }
      __syncthreads();
      const float xi = Xs[i][lc];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = lr + 8 * e;
        if (j > i && j < ib) Xs[j][lc] = fmaf(-Qs[i][j], xi, Xs[j][lc]);
      }
    }
    __syncthreads();   // X rows of this step are visible to the next step's k-loop (same CTA)
  }
}

// Columns [jb0, jb1) of  X = B Q^-1  for one 32-row slab per CTA, assuming columns < jb0 have been eliminated.
//   X[:,J] = ( B[:,J] - sum_{jb0 <= k < j0} X[:, k] Q[k, J] ) Qjj^-1
__global__ void __launch_bounds__(256) trsm_right_block_kernel(const float* __restrict__ Q, int ldq,
                                                               const float* B, int ldb, float* X, int ldx, int m,
                                                               int jb0, int jb1) {
  __shared__ float Qs[NB][NB + 1];   // Qs[k][j] = Q[k0+k, j0+j]
  __shared__ float Xs[NB][NB + 1];   // Xs[r][k] = X[r0+r, k0+k]
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * NB;
  const int lr = tid / NB;          // 0..7
  const int lc = tid % NB;          // 0..31
  for (int j0 = jb0; j0 < jb1; j0 += NB) {
    const int jb = min(NB, jb1 - j0);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};   // rows lr+8e, column lc
    float qn[4], xn[4];                     // next k0 step's tiles, in flight during the multiply
    auto fetch = [&](int k0) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = lr + 8 * e;
        qn[e] = (j0 + lc < jb1) ? Q[(size_t)(k0 + r) * ldq + j0 + lc] : 0.f;
        xn[e] = (r0 + r < m) ? X[(size_t)(r0 + r) * ldx + k0 + lc] : 0.f;
      }
    };
    if (jb0 < j0) fetch(jb0);
    for (int k0 = jb0; k0 < j0; k0 += NB) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = lr + 8 * e;
        Qs[r][lc] = qn[e];
        Xs[r][lc] = xn[e];
      }
      __syncthreads();
      if (k0 + NB < j0) fetch(k0 + NB);
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        const float q = Qs[k][lc];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = fmaf(Xs[lr + 8 * e][k], q, acc[e]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = lr + 8 * e;
      Qs[r][lc] = (r < jb && lc < jb && r <= lc) ? Q[(size_t)(j0 + r) * ldq + j0 + lc] : 0.f;   // Qs[k][j], k<=j
      float b = 0.f;
      if (r0 + r < m && lc < jb) b = B[(size_t)(r0 + r) * ldb + j0 + lc] - acc[e];
      Xs[r][lc] = b;
    }
    __syncthreads();
    // right-looking over the 32 columns with all 256 threads (see trsm_left_block_kernel): column j is finished by the
    // threads that own it, one barrier, then every thread folds it into its own columns; same operations, same order
    for (int j = 0; j < jb; ++j) {
      if (lc == j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = lr + 8 * e;
          const float x = Xs[r][j] / Qs[j][j];
          Xs[r][j] = x;
          if (r0 + r < m) X[(size_t)(r0 + r) * ldx + j0 + j] = x;
        }
      }
      __syncthreads();
      if (lc > j && lc < jb) {
        const float q = Qs[j][lc];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = lr + 8 * e;
          Xs[r][lc] = fmaf(-Xs[r][j], q, Xs[r][lc]);
        }
      }
    }
    __syncthreads();
  }
}

int trsm_left_block(psgd_ctx* ctx, const float* Q, int ldq, const float* B, int ldb, float* X, int ldx, int m, int ib0,
                    int ib1) {
  if (ib1 <= ib0 || m <= 0) return PSGD_OK;
  ProfScope prof(ctx, PSGD_K_TRSM, (double)m * (ib1 - ib0) * (ib1 - ib0));
  trsm_left_block_kernel<<<(m + NB - 1) / NB, 256, 0, ctx->stream>>>(Q, ldq, B, ldb, X, ldx, m, ib0, ib1);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}
int trsm_right_block(psgd_ctx* ctx, const float* Q, int ldq, const float* B, int ldb, float* X, int ldx, int m, int jb0,
                     int jb1) {
  if (jb1 <= jb0 || m <= 0) return PSGD_OK;
  ProfScope prof(ctx, PSGD_K_TRSM, (double)m * (jb1 - jb0) * (jb1 - jb0));
  trsm_right_block_kernel<<<(m + NB - 1) / NB, 256, 0, ctx->stream>>>(Q, ldq, B, ldb, X, ldx, m, jb0, jb1);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

int trsm_left_upper_adjoint(psgd_ctx* ctx, const float* Q, int ldq, const float* B, int ldb, float* X, int ldx,
                            int n, int m) {
  return trsm_left_block(ctx, Q, ldq, B, ldb, X, ldx, m, 0, n);
}

int trsm_right_upper(psgd_ctx* ctx, const float* Q, int ldq, const float* B, int ldb, float* X, int ldx, int m,
                     int n) {
  return trsm_right_block(ctx, Q, ldq, B, ldb, X, ldx, m, 0, n);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int ld_in,
                                                        float* __restrict__ out, int ld_out, int rows, int cols) {
  __shared__ float t[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int r = r0 + ty + 8 * e, c = c0 + tx;
    t[ty + 8 * e][tx] = (r < rows && c < cols) ? in[(size_t)r * ld_in + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = c0 + ty + 8 * e, r = r0 + tx;
    if (r < rows && c < cols) out[(size_t)c * ld_out + r] = t[tx][ty + 8 * e];
  }
}

int transpose(psgd_ctx* ctx, const float* in, int ld_in, float* out, int ld_out, int rows, int cols) {
  if (rows <= 0 || cols <= 0) return PSGD_OK;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  transpose_kernel<<<grid, 256, 0, ctx->stream>>>(in, ld_in, out, ld_out, rows, cols);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

}  // namespace la
}  // namespace psgd
