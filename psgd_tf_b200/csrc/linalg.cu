// SIMT fp32 engine of the device linear-algebra vocabulary (see linalg.cuh).
#include "linalg.cuh"
#include "kron_stream.cuh"

namespace psgd {
namespace la {

// ---------------------------------------------------------------------------------------------
// GEMM, general engine: 64 x 64 x 16 tiles, 4 x 4 outputs per thread, 256 threads, fp32 FMA accumulation in ascending K
// order.  Problems whose 64-wide grid would leave most SMs idle go to gemm_small_kernel below.
// ---------------------------------------------------------------------------------------------
constexpr int PAD = 4;

// One K step of one product: this thread's elements of the A and B tiles, global -> registers (fetch) and registers ->
// shared memory (stash).  Kept apart so that the loads of step s+1 are in flight while step s is multiplied.
template <int T, int BK, int NT>
struct TileRegs {
  float a[(T * BK) / NT], b[(T * BK) / NT];
};
template <int T, int BK, int NT>
__device__ __forceinline__ void fetch_tiles(const float* __restrict__ A, int lda, bool ta, const float* __restrict__ B, int ldb,
                                            bool tb, int M, int N, int K, int m0, int n0, int k0, int tid,
                                            TileRegs<T, BK, NT>& r) {
#pragma unroll
  for (int e = 0; e < (T * BK) / NT; ++e) {
    const int idx = tid + NT * e;
    int m, k;
    if (ta) { m = idx % T; k = idx / T; } else { k = idx % BK; m = idx / BK; }
    const int gm = m0 + m, gk = k0 + k;
    r.a[e] = (gm < M && gk < K) ? (ta ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk]) : 0.f;
  }
#pragma unroll
  for (int e = 0; e < (T * BK) / NT; ++e) {
    const int idx = tid + NT * e;
    int n, k;
    if (tb) { k = idx % BK; n = idx / BK; } else { n = idx % T; k = idx / T; }
    const int gn = n0 + n, gk = k0 + k;
    r.b[e] = (gn < N && gk < K) ? (tb ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn]) : 0.f;
  }
}
template <int T, int BK, int NT>
__device__ __forceinline__ void stash_tiles(bool ta, bool tb, int tid, const TileRegs<T, BK, NT>& r, float (*As)[T + PAD],
                                            float (*Bs)[T + PAD]) {
#pragma unroll
  for (int e = 0; e < (T * BK) / NT; ++e) {
    const int idx = tid + NT * e;
    int m, k;
    if (ta) { m = idx % T; k = idx / T; } else { k = idx % BK; m = idx / BK; }
    As[k][m] = r.a[e];
  }
#pragma unroll
  for (int e = 0; e < (T * BK) / NT; ++e) {
    const int idx = tid + NT * e;
    int n, k;
    if (tb) { k = idx % BK; n = idx / BK; } else { n = idx % T; k = idx / T; }
    Bs[k][n] = r.b[e];
  }
}

template <int T, int BK, int NT>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(Gemm g) {
  constexpr int R = T / 16;                         // outputs per thread along each dimension
  static_assert(NT == 256, "16 x 16 threads");
  __shared__ __align__(16) float tiles[2 * BK * (T + PAD)];
  float (*As)[T + PAD] = reinterpret_cast<float (*)[T + PAD]>(tiles);
  float (*Bs)[T + PAD] = reinterpret_cast<float (*)[T + PAD]>(tiles + BK * (T + PAD));
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * T, n0 = blockIdx.x * T;
  if (g.triu && m0 >= n0 + T) {
    // tile entirely below the diagonal: result is zero after masking
    for (int e = tid; e < T * T; e += NT) {
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
  auto fetch = [&](int s, TileRegs<T, BK, NT>& r) {
    const bool p1 = s >= steps0;
    const int k0 = (p1 ? s - steps0 : s) * BK;
    if (p1) fetch_tiles<T, BK, NT>(g.A2, g.lda2, g.ta2, g.B2, g.ldb2, g.tb2, g.M, g.N, g.K2, m0, n0, k0, tid, r);
    else fetch_tiles<T, BK, NT>(g.A, g.lda, g.ta, g.B, g.ldb, g.tb, g.M, g.N, g.K, m0, n0, k0, tid, r);
  };
  TileRegs<T, BK, NT> cur, nxt;
  if (steps > 0) {
    fetch(0, cur);
    stash_tiles<T, BK, NT>(steps0 > 0 ? g.ta : g.ta2, steps0 > 0 ? g.tb : g.tb2, tid, cur, As, Bs);
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
      stash_tiles<T, BK, NT>(p1 ? g.ta2 : g.ta, p1 ? g.tb2 : g.tb, tid, nxt, As, Bs);
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

// ---------------------------------------------------------------------------------------------
// Small problems (the 64 x 64 grid would leave most SMs idle: LeNet5 layers): one 32 x 32 output tile per CTA of 1024
// threads.  A 32 x 32 x K product on 8 warps is a chain of dependent shared-memory loads and FMAs (2 x 2 outputs per
// thread: 5 instructions per FMA, issued at a sixth of the rate -- 10 us per launch at K = 257, ncu).  Here each
// thread keeps a 4 x 4 block of the tile (two 128-bit loads per 16 FMAs) and 1/16 of the K range: 8 column blocks x 4 K
// slices per warp, 8 row blocks x 4 K slices across the warps.  The K slices meet by warp shuffles, then through shared
// memory in a fixed order, and the epilogue runs one element per thread, coalesced.
// ---------------------------------------------------------------------------------------------
constexpr int ST = 32, SBK = 64, SLD = ST + PAD;
__global__ void __launch_bounds__(1024) gemm_small_kernel(Gemm g) {
  __shared__ __align__(16) float tiles[2 * SBK * SLD];
  float (*As)[SLD] = reinterpret_cast<float (*)[SLD]>(tiles);
  float (*Bs)[SLD] = reinterpret_cast<float (*)[SLD]>(tiles + SBK * SLD);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * ST, n0 = blockIdx.x * ST;
  const int om = tid >> 5, on = tid & 31;              // output role: one element of the tile
  if (g.triu && m0 >= n0 + ST) {                      // tile entirely below the diagonal: zero after masking
    if (m0 + om < g.M && n0 + on < g.N) g.C[(size_t)(m0 + om) * g.ldc + n0 + on] = 0.f;
    return;
  }
  const int ti = w & 7, tj = lane & 7;                // multiply role: rows 4 ti.., columns 4 tj.. of the tile,
  const int kq = ((w >> 3) << 2) + (lane >> 3);       // K rows 4 kq .. 4 kq + 3 of every step
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int steps0 = (g.K > 0 && g.A) ? (g.K + SBK - 1) / SBK : 0;
  const int steps1 = (g.K2 > 0 && g.A2) ? (g.K2 + SBK - 1) / SBK : 0;
  const int steps = steps0 + steps1;
  // Tile traffic of this thread: two elements of A and two of B per step.  Everything that does not change from step
  // to step (global pointer, K offset, validity of the row / column, slot in shared memory) is worked out once per
  // product -- with 1024 threads the address arithmetic of a generic fetch costs more issue slots than the multiply.
  // The subtracted product is negated on its way into shared memory, so the multiply loop never looks at a sign.
  const float* pa[2]; const float* pb[2];
  int ka[2], kb[2], sa[2], sb[2];
  bool oka[2], okb[2];
  size_t stride_a = 0, stride_b = 0;
  int kdim = 0, knext = 0;
  float sgn = 1.f;
  auto setup = [&](const float* A, int lda, bool ta, const float* B, int ldb, bool tb, int K, float sign) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int idx = tid + 1024 * e;
      int m, k;
      if (ta) { m = idx % ST; k = idx / ST; } else { k = idx % SBK; m = idx / SBK; }
      oka[e] = m0 + m < g.M; ka[e] = k; sa[e] = k * SLD + m;
      pa[e] = oka[e] ? (ta ? A + (size_t)k * lda + m0 + m : A + (size_t)(m0 + m) * lda + k) : A;
      int n;
      if (tb) { k = idx % SBK; n = idx / SBK; } else { n = idx % ST; k = idx / ST; }
      okb[e] = n0 + n < g.N; kb[e] = k; sb[e] = SBK * SLD + k * SLD + n;
      pb[e] = okb[e] ? (tb ? B + (size_t)(n0 + n) * ldb + k : B + (size_t)k * ldb + n0 + n) : B;
    }
    stride_a = ta ? (size_t)SBK * lda : (size_t)SBK;
    stride_b = tb ? (size_t)SBK : (size_t)SBK * ldb;
    kdim = K; knext = 0; sgn = sign;
  };
  float ra[2], rb[2];
  int fs = 0;                                             // next step to fetch
  auto fetch = [&]() {
    if (fs == steps0 && steps1 > 0) setup(g.A2, g.lda2, g.ta2, g.B2, g.ldb2, g.tb2, g.K2, -1.f);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      ra[e] = (oka[e] && knext + ka[e] < kdim) ? *pa[e] : 0.f;
      rb[e] = (okb[e] && knext + kb[e] < kdim) ? *pb[e] : 0.f;
      pa[e] += stride_a; pb[e] += stride_b;
    }
    knext += SBK; ++fs;
  };
  if (steps0 > 0) setup(g.A, g.lda, g.ta, g.B, g.ldb, g.tb, g.K, 1.f);
  if (steps > 0) {
    fetch();
#pragma unroll
    for (int e = 0; e < 2; ++e) { tiles[sa[e]] = sgn * ra[e]; tiles[sb[e]] = rb[e]; }
  }
  __syncthreads();
  for (int s = 0; s < steps; ++s) {
    const bool more = s + 1 < steps;
    if (more) fetch();                                    // in flight during the multiply below
    const int da[2] = {sa[0], sa[1]}, db[2] = {sb[0], sb[1]};   // slots and sign of the step just fetched (product switch);
    const float sg = sgn;                                       // the sign goes on at the stash: nothing may depend
                                                                // on the loads before the multiply below has been issued
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = 4 * kq + kk;
      const float4 a = *reinterpret_cast<const float4*>(&As[k][4 * ti]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][4 * tj]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
    if (more) {
#pragma unroll
      for (int e = 0; e < 2; ++e) { tiles[da[e]] = sg * ra[e]; tiles[db[e]] = rb[e]; }
      __syncthreads();
    }
  }
  // K slices: inside the warp (lanes 8 and 16 apart), then the four warps of a row block through shared memory
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[i][j];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      acc[i][j] = v;
    }
  float (*red)[ST][ST + 1] = reinterpret_cast<float (*)[ST][ST + 1]>(tiles);   // the tiles are dead behind the last barrier
  static_assert(4 * ST * (ST + 1) <= 2 * SBK * SLD, "partial sums reuse the tile buffers");
  if (lane < 8) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[w >> 3][4 * ti + i][4 * tj + j] = acc[i][j];
  }
  __syncthreads();
  const int m = m0 + om, n = n0 + on;
  float mx = 0.f;
  if (m < g.M && n < g.N) {
    float v = (red[0][om][on] + red[1][om][on]) + (red[2][om][on] + red[3][om][on]);
    if (g.colscale) {
      float sc = g.colscale[n];
      if (g.colscale_sq) sc = sc * sc;
      v = g.colscale_recip ? v * (1.0f / sc) : v * sc;
    }
    if (g.triu && m > n) v = 0.f;
    if (g.D) {
      const float mu = g.mu_max ? g.step / (*g.mu_max + g.tiny) : 1.0f;
      v = g.D[(size_t)m * g.ldd + n] - mu * v;
    }
    if (g.rho_mode == 1) v = v / *g.rho;
    else if (g.rho_mode == 2) v = v * *g.rho;
    mx = fabsf(v);
    g.C[(size_t)m * g.ldc + n] = v;
  }
  if (g.maxabs) {
    mx = warp_max(mx);
    if (lane == 0 && mx > 0.f) atomic_max_nonneg(g.maxabs, mx);
  }
}

int gemm_simt(psgd_ctx* ctx, const Gemm& g) {
  if (g.M <= 0 || g.N <= 0) return PSGD_OK;
  const int ctas64 = ((g.N + 63) / 64) * ((g.M + 63) / 64);
  if (ctas64 * 2 <= ctx->num_sms) {
    dim3 grid((g.N + ST - 1) / ST, (g.M + ST - 1) / ST);
    gemm_small_kernel<<<grid, 1024, 0, ctx->stream>>>(g);
  } else {
    dim3 grid((g.N + 63) / 64, (g.M + 63) / 64);
    gemm_simt_kernel<64, 16, 256><<<grid, 256, 0, ctx->stream>>>(g);
  }
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

// ---------------------------------------------------------------------------------------------
// Triangular solves for layers below the tensor-core thresholds (LeNet5 / NMT factors, odd sizes, the vector solve of
// the dense preconditioner).  One CTA of 1024 threads per 32-wide slab of the right-hand side (32 columns of B for the
// left solve, 32 rows for the right one) and per PANEL of up to 512 unknowns:
//   * the panel's 32 x 32 diagonal blocks are inverted up front, one warp per block, in shared memory (back
//     substitution on the identity, one column per lane), so that a block step is a 32 x 32 x 32 product and not a
//     32-long chain of divide / barrier / update rounds;
//   * the slab of the right-hand side lives in shared memory for the whole panel and is overwritten by the solution,
//     so a step never waits on a global-memory round trip for values this CTA produced itself;
//   * only the off-diagonal tiles of Q stream in from global memory (64 x 32 at a time, double buffered, the next
//     tile -- also across a step boundary -- in flight while the current one is multiplied);
//   * both solves are the same computation  out[u][s] = sum_k Q[k][u] X[k][s]  (u: unknown inside the block, s: position
//     inside the slab), so one kernel serves both; they differ in how (unknown, slab) maps to memory.  Each thread keeps
//     4 unknowns x 1 slab position and a quarter of the K range (one 128-bit broadcast load of Q and one load of X per 4
//     FMAs; with one output per thread the kernel was bound by shared-memory wavefronts, 2 per FMA: 74 us at n = 257);
//     the four K quarters meet through shared memory in a fixed order.
// ncu, round 2, n = 257, m = 120: the 256-thread row-by-row kernels before these ran 26 k dependent instructions per
// warp at 8 cycles each (2 warps per scheduler): 120 us + 108 us per layer; these take 35 us + 13 us (clock64 stamps,
// tools/trsm_stamps.py: slab load 2.5 k cycles, block inversions 7.8 k, then per block step ~2.0 k per 64-row tile of Q
// and 1.6 k for the reduction + diagonal product).  Panels beyond the first get the contribution of the solved part from
// one SIMT GEMM each (host loop below).
// ---------------------------------------------------------------------------------------------
constexpr int NB = 32;             // diagonal block
constexpr int kPanel = 512;        // unknowns per launch
constexpr int kPanelBlocks = kPanel / NB;
constexpr int KT = 64;             // K rows per staged tile of Q
constexpr int XP = NB + 1;         // pitch of the slab / staging arrays that are read with scalar loads
constexpr size_t kTrsmSmem =
    ((size_t)kPanelBlocks * NB * NB + (size_t)kPanel * XP + 2 * KT * NB + NB * XP + 4 * NB * XP) * sizeof(float);

// In-place inverse of one upper-triangular 32 x 32 block held in shared memory (row-major, pitch 32), by one warp:
// lane c solves U x = e_c from the last row up; row i of U is dead once every lane has used it, so x_i overwrites it.
// The reciprocals of the diagonal are formed by all lanes at once.  Row i is read as 128-bit loads, four at a time, ahead
// of the FMAs that use them, and the dot product runs as four independent chains: with one scalar load in front of
// each FMA (what the compiler emits for the plain loop under the 64-register cap of a 1024-thread CTA) the 496 FMAs cost
// a shared-memory round trip each -- 17.2 k cycles per block, as much as all 16 block steps of a panel (clock64 stamps).
__device__ __forceinline__ void invert_block_warp(float* U, int lane) {
  const float rdiag = 1.0f / U[lane * NB + lane];
  float x[NB];
#pragma unroll
  for (int i = NB - 1; i >= 0; --i) {
    float s0 = (i == lane) ? 1.f : 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {                        // columns 16 h .. 16 h + 15 of row i
      float4 u[4];
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (16 * h + 4 * g + 3 > i) u[g] = *reinterpret_cast<const float4*>(U + i * NB + 16 * h + 4 * g);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int k = 16 * h + 4 * g;
        if (k + 0 > i) s0 = fmaf(-u[g].x, x[k + 0], s0);
        if (k + 1 > i) s1 = fmaf(-u[g].y, x[k + 1], s1);
        if (k + 2 > i) s2 = fmaf(-u[g].z, x[k + 2], s2);
        if (k + 3 > i) s3 = fmaf(-u[g].w, x[k + 3], s3);
      }
    }
    x[i] = ((s0 + s1) + (s2 + s3)) * __shfl_sync(0xffffffffu, rdiag, i);
    __syncwarp();
    U[i * NB + lane] = x[i];
  }
  __syncwarp();
}

// Diagonal blocks of Q[b0 : b1, b0 : b1] -> W (block t at W + t * 1024), upper triangles only; rows / columns beyond b1
// are padded with the identity.  Warps 0 .. blocks-1 work, the others fall through.  The 32 loads of a lane are
// unconditional (clamped addresses, masked afterwards) so that they are all in flight together: as predicated loads
// feeding a store each they went out one L2 round trip at a time, 14 k cycles per block.
__device__ __forceinline__ void load_and_invert_blocks(const float* __restrict__ Q, int ldq, int b0, int b1, float* W,
                                                       int warp, int lane) {
  const int blocks = (b1 - b0 + NB - 1) / NB;
  if (warp < blocks) {
    float* U = W + warp * NB * NB;
    const int d0 = b0 + warp * NB;
    const int col = d0 + lane < b1 ? d0 + lane : b1 - 1;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float v[NB / 2];
#pragma unroll
      for (int j = 0; j < NB / 2; ++j) {
        const int i = h * (NB / 2) + j;
        const int row = d0 + i < b1 ? d0 + i : b1 - 1;
        v[j] = Q[(size_t)row * ldq + col];
      }
#pragma unroll
      for (int j = 0; j < NB / 2; ++j) {
        const int i = h * (NB / 2) + j;
        const bool in = d0 + i < b1 && d0 + lane < b1;
        U[i * NB + lane] = in ? (lane >= i ? v[j] : 0.f) : (i == lane ? 1.f : 0.f);
      }
    }
    __syncwarp();
    invert_block_warp(U, lane);
  }
}

// Unknowns [b0, b1) (at most kPanel) of one slab, assuming the unknowns before b0 have been eliminated already (B holds
// the right-hand side minus their contribution):
//   LEFT :  X = Q^-T B, unknown t = row of X, slab = 32 columns:   X[T,:] = Wtt^T ( B[T,:] - sum_k Q[k,T]^T X[k,:] )
//   RIGHT:  X = B Q^-1, unknown t = column of X, slab = 32 rows:   X[:,T] = ( B[:,T] - sum_k X[:,k] Q[k,T] ) Wtt
// with Wtt = Qtt^-1 and k over the earlier blocks of the panel.
template <bool LEFT>
__global__ void __launch_bounds__(1024) trsm_panel_kernel(const float* __restrict__ Q, int ldq, const float* B, int ldb,
                                                          float* X, int ldx, int m, int b0, int b1, long long* stamps) {
  extern __shared__ __align__(16) float trsm_smem[];
  int nstamp = 0;
  auto stamp = [&]() { if (stamps && threadIdx.x == 0 && blockIdx.x == 0) stamps[nstamp++] = clock64(); };   // tools/trsv_stamps.py
  stamp();
  float* W = trsm_smem;                                   // [blocks][32][32]   inverted diagonal blocks
  float* Xs = W + kPanelBlocks * NB * NB;                 // [unknown][XP]      right-hand side, then the solution
  float* Qs = Xs + kPanel * XP;                           // [2][KT][32]        Qs[k][u] = Q[k0+k, u0+u]
  float* Bs = Qs + 2 * KT * NB;                           // [32][XP]           right-hand side of the block step
  float* Ps = Bs + NB * XP;                               // [4][32][XP]        partial sums of the K quarters
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int s0 = blockIdx.x * NB;
  const int span = ((b1 - b0 + NB - 1) / NB) * NB;
  // (unknown t, slab position s) in global memory
  auto at = [&](int t, int s, int ld) -> size_t {
    return LEFT ? (size_t)t * ld + s0 + s : (size_t)(s0 + s) * ld + t;
  };
  if (LEFT) {
    for (int t = w; t < span; t += 32)
      Xs[t * XP + lane] = (b0 + t < b1 && s0 + lane < m) ? B[at(b0 + t, lane, ldb)] : 0.f;
  } else {
    for (int t = lane; t < span; t += 32)
      Xs[t * XP + w] = (b0 + t < b1 && s0 + w < m) ? B[at(b0 + t, w, ldb)] : 0.f;
  }
  stamp();
  load_and_invert_blocks(Q, ldq, b0, b1, W, w, lane);
  stamp();

  // multiply role: K quarter kg, unknowns 4 ug .. 4 ug + 3 of the block, slab position = lane
  const int kg = w >> 3, ug = w & 7;
  // output role: one (unknown, slab position) per thread, lanes along the contiguous direction of X
  const int ou = LEFT ? w : lane, os = LEFT ? lane : w;
  const bool os_ok = s0 + os < m;

  // off-diagonal tiles in the order they are consumed: (block u0, k0), k0 = b0, b0 + KT, ... < u0
  float q0, q1;
  auto fetch = [&](int u0, int k0) {
    const int col = u0 + lane;
    const int ka = k0 + w, kb = k0 + w + 32;
    q0 = (ka < u0 && col < b1) ? Q[(size_t)ka * ldq + col] : 0.f;
    q1 = (kb < u0 && col < b1) ? Q[(size_t)kb * ldq + col] : 0.f;
  };
  int pu = b0 + NB, pk = b0;                              // the tile in flight
  if (pu < b1) fetch(pu, pk);
  int buf = 0;
  __syncthreads();
  for (int u0 = b0; u0 < b1; u0 += NB) {
    stamp();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = b0; k0 < u0; k0 += KT) {
      float* Qb = Qs + buf * KT * NB;
      Qb[w * NB + lane] = q0;
      Qb[(w + 32) * NB + lane] = q1;
      __syncthreads();
      pk += KT;
      if (pk >= pu) { pu += NB; pk = b0; }
      if (pu < b1) fetch(pu, pk);
      const float* qb = Qb + (kg * (KT / 4)) * NB + 4 * ug;
      const float* xs = Xs + (k0 - b0 + kg * (KT / 4)) * XP + lane;
#pragma unroll
      for (int k = 0; k < KT / 4; ++k) {
        const float4 q = *reinterpret_cast<const float4*>(qb + k * NB);
        const float x = xs[k * XP];
        acc[0] = fmaf(q.x, x, acc[0]); acc[1] = fmaf(q.y, x, acc[1]);
        acc[2] = fmaf(q.z, x, acc[2]); acc[3] = fmaf(q.w, x, acc[3]);
      }
      buf ^= 1;
    }
    stamp();
    const int to = u0 - b0 + ou;                          // this thread's unknown inside the panel (output role)
    float bval = Xs[to * XP + os];
    if (u0 > b0) {
#pragma unroll
      for (int e = 0; e < 4; ++e) Ps[(kg * NB + 4 * ug + e) * XP + lane] = acc[e];
      __syncthreads();
      const float* pp = Ps + ou * XP + os;
      bval -= (pp[0] + pp[NB * XP]) + (pp[2 * NB * XP] + pp[3 * NB * XP]);
    }
    Bs[ou * XP + os] = bval;
    __syncthreads();
    {
      const float* wb = W + ((u0 - b0) / NB) * NB * NB + (kg * (NB / 4)) * NB + 4 * ug;   // Wtt[k][u], zero for k > u
      const float* bs = Bs + (kg * (NB / 4)) * XP + lane;
      float xa[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < NB / 4; ++k) {
        const float4 q = *reinterpret_cast<const float4*>(wb + k * NB);
        const float bb = bs[k * XP];
        xa[0] = fmaf(q.x, bb, xa[0]); xa[1] = fmaf(q.y, bb, xa[1]);
        xa[2] = fmaf(q.z, bb, xa[2]); xa[3] = fmaf(q.w, bb, xa[3]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) Ps[(kg * NB + 4 * ug + e) * XP + lane] = xa[e];
    }
    __syncthreads();
    const float* pp = Ps + ou * XP + os;
    const float x = (pp[0] + pp[NB * XP]) + (pp[2 * NB * XP] + pp[3 * NB * XP]);
    Xs[to * XP + os] = x;
    if (u0 + ou < b1 && os_ok) X[at(u0 + ou, os, ldx)] = x;
    stamp();
    // every later use of Xs / Bs / Ps sits behind the barrier that follows the next tile stash
  }
}

// Vector right-hand side (the dense preconditioner's b = Q^-T dx, psgd.py:39): unknowns [b0, b1) of one panel by ONE CTA,
// right-looking.  Thread t owns unknown b0 + t: once warp 0 has finished a block of 32 unknowns with the inverted diagonal
// block, every thread to the right of the block takes the block's contribution off its own right-hand side,
//   rhs[t] -= sum_{k in block} Q[k, b0 + t] x[k],
// with the 32 rows of Q it needs read a step ahead (they do not depend on the solution) as 32 coalesced loads -- lanes
// are consecutive columns -- so a step is two barriers, one 32 x 32 product on warp 0 and a chain of 32 FMAs.  (The
// left-looking form of the tiled kernel above reduces over the warps every step and issues its predicated loads for all
// 15 possible earlier blocks: 41 us per 512-row panel against 9 us of start-up, ncu.)  The contribution of the unknowns
// before b0 arrives as the vector `sub` = Q[0:b0, b0:b1]^T x[0:b0] (ks::col_wsum: the bandwidth-bound bulk of the solve,
// spread over the whole GPU) and is subtracted here.
constexpr size_t kTrsvSmem = ((size_t)kPanelBlocks * NB * NB + kPanel + NB) * sizeof(float);
__global__ void __launch_bounds__(1024) trsv_panel_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ B,
                                                          float* X, int b0, int b1, const float* __restrict__ sub,
                                                          long long* stamps) {
  extern __shared__ __align__(16) float trsm_smem[];
  int nstamp = 0;
  auto stamp = [&]() { if (stamps && threadIdx.x == 0) stamps[nstamp++] = clock64(); };   // tools/trsv_stamps.py
  stamp();
  float* W = trsm_smem;                                   // [blocks][32][32]
  float* xs = W + kPanelBlocks * NB * NB;                 // [kPanel]  right-hand side of the unknowns still open
  float* xn = xs + kPanel;                                // [32]      the block just solved
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int span = ((b1 - b0 + NB - 1) / NB) * NB;
  // the panel's triangle of Q on its way into L2 before the block steps ask for it
  if (lane < span / NB) {                                  // lane = column block, the warps stride the rows
    const float* qp = Q + (size_t)(b0 + w) * ldq + b0 + lane * NB;
    for (int r = w; r < b1 - b0 && r <= lane * NB + NB - 1; r += 32, qp += (size_t)32 * ldq)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(qp));
  }
  for (int t = tid; t < span; t += 1024) {
    float v = 0.f;
    if (b0 + t < b1) {
      v = B[b0 + t];
      if (sub) v -= sub[t];
    }
    xs[t] = v;
  }
  stamp();
  load_and_invert_blocks(Q, ldq, b0, b1, W, w, lane);
  stamp();
  const bool mine = tid < span && b0 + tid < b1;          // warp w owns block w of the panel
  float qv[NB];
  // Q[u0 + k, b0 + tid], k < 32, for the blocks right of u0.  Warp-uniform branch and a running pointer: with 32 warps in
  // the CTA, per-element predicates and 64-bit address arithmetic executed by every warp (also the idle ones) cost
  // 5.3 k cycles per step (clock64 stamps, tools/trsv_stamps.py) against 0.5 k for everything else.
  auto load_rows = [&](int u0) {
    if (b0 + w * NB < u0 + NB || w * NB >= span) return;
    const int kmax = b1 - u0 < NB ? b1 - u0 : NB;
    const float* q = Q + (size_t)u0 * ldq + b0 + tid;
#pragma unroll
    for (int k = 0; k < NB; ++k, q += ldq) qv[k] = (k < kmax && mine) ? *q : 0.f;
  };
  __syncthreads();
  stamp();
  load_rows(b0);
  for (int u0 = b0; u0 < b1; u0 += NB) {
    stamp();
    if (w == 0) {
      const float* Wb = W + ((u0 - b0) / NB) * NB * NB;
      const float* bs = xs + (u0 - b0);
      float x0 = 0.f, x1 = 0.f;
#pragma unroll
      for (int k = 0; k < NB; k += 2) {                  // Wb[k][lane] = 0 for k > lane
        x0 = fmaf(Wb[k * NB + lane], bs[k], x0);
        x1 = fmaf(Wb[(k + 1) * NB + lane], bs[k + 1], x1);
      }
      const float x = x0 + x1;
      xn[lane] = x;
      if (u0 + lane < b1) X[u0 + lane] = x;
    }
    stamp();
    __syncthreads();
    stamp();
    if (mine && b0 + w * NB >= u0 + NB) {
      float acc = xs[tid];
#pragma unroll
      for (int k = 0; k < NB; ++k) acc = fmaf(-qv[k], xn[k], acc);
      xs[tid] = acc;
    }
    if (u0 + NB < b1) load_rows(u0 + NB);
    stamp();
    __syncthreads();
  }
  stamp();
}

static DeviceOnce trsm_attr_done;
static int ensure_trsm_attrs(psgd_ctx* ctx) {
  if (trsm_attr_done.done(ctx->device)) return PSGD_OK;
  PSGD_CUDA_CHECK(cudaFuncSetAttribute(trsm_panel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrsmSmem));
  PSGD_CUDA_CHECK(cudaFuncSetAttribute(trsm_panel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrsmSmem));
  PSGD_CUDA_CHECK(cudaFuncSetAttribute(trsv_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrsvSmem));
  trsm_attr_done.set(ctx->device);
  return PSGD_OK;
}

int trsm_left_block(psgd_ctx* ctx, const float* Q, int ldq, const float* B, int ldb, float* X, int ldx, int m, int ib0,
                    int ib1) {
  if (ib1 <= ib0 || m <= 0) return PSGD_OK;
  PSGD_REQUIRE(ib1 - ib0 <= kPanel, PSGD_ERR_BAD_SHAPE, "trsm panel of %d rows", ib1 - ib0);
  PSGD_RETURN_IF(ensure_trsm_attrs(ctx));
  ProfScope prof(ctx, PSGD_K_TRSM, (double)m * (ib1 - ib0) * (ib1 - ib0));
  trsm_panel_kernel<true><<<(m + NB - 1) / NB, 1024, kTrsmSmem, ctx->stream>>>(Q, ldq, B, ldb, X, ldx, m, ib0, ib1, ctx->opt_stamp_ptr);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}
int trsm_right_block(psgd_ctx* ctx, const float* Q, int ldq, const float* B, int ldb, float* X, int ldx, int m, int jb0,
                     int jb1) {
  if (jb1 <= jb0 || m <= 0) return PSGD_OK;
  PSGD_REQUIRE(jb1 - jb0 <= kPanel, PSGD_ERR_BAD_SHAPE, "trsm panel of %d columns", jb1 - jb0);
  PSGD_RETURN_IF(ensure_trsm_attrs(ctx));
  ProfScope prof(ctx, PSGD_K_TRSM, (double)m * (jb1 - jb0) * (jb1 - jb0));
  trsm_panel_kernel<false><<<(m + NB - 1) / NB, 1024, kTrsmSmem, ctx->stream>>>(Q, ldq, B, ldb, X, ldx, m, jb0, jb1, ctx->opt_stamp_ptr ? ctx->opt_stamp_ptr + 128 : nullptr);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

size_t trsv_ws_floats(int n) { return n > kPanel ? ((size_t)ks::row_tiles(n) + 1) * kPanel : 0; }

// x = Q^-T b for one vector: panels of 512 unknowns; before each panel but the first, Q[0:p0, p0:p1]^T x[0:p0] as
// column sums over the whole GPU (the n^2 / 2 floats of Q that the solve has to read go by at HBM speed), then the
// panel itself on one CTA.  `ws`: trsv_ws_floats(n) floats.
int trsv_left_upper_adjoint(psgd_ctx* ctx, const float* Q, int ldq, const float* b, float* x, int n, float* ws) {
  if (n <= 0) return PSGD_OK;
  PSGD_RETURN_IF(ensure_trsm_attrs(ctx));
  ProfScope prof(ctx, PSGD_K_TRSM, (double)n * n);
  for (int p0 = 0; p0 < n; p0 += kPanel) {
    const int p1 = p0 + kPanel < n ? p0 + kPanel : n;
    const float* sub = nullptr;
    if (p0 > 0) {
      PSGD_RETURN_IF(ks::col_wsum(ctx, 1, nullptr, x, Q + p0, ldq, p0, p1 - p0, ws + kPanel, ws));
      sub = ws;
    }
    trsv_panel_kernel<<<1, 1024, kTrsvSmem, ctx->stream>>>(Q, ldq, b, x, p0, p1, sub, p0 == 0 ? ctx->opt_stamp_ptr : nullptr);
    PSGD_LAUNCH_CHECK(ctx);
  }
  return PSGD_OK;
}

// Panels after the first: the solved part's contribution comes off the right-hand side in one GEMM, written where the
// solution of the panel will go (X may alias B: the GEMM is elementwise in D and C), then the panel is solved in place.
int trsm_left_upper_adjoint(psgd_ctx* ctx, const float* Q, int ldq, const float* B, int ldb, float* X, int ldx,
                            int n, int m) {
  for (int p0 = 0; p0 < n; p0 += kPanel) {
    const int p1 = p0 + kPanel < n ? p0 + kPanel : n;
    if (p0 == 0) {
      PSGD_RETURN_IF(trsm_left_block(ctx, Q, ldq, B, ldb, X, ldx, m, 0, p1));
      continue;
    }
    Gemm g;                                                // X[P,:] = B[P,:] - Q[0:p0, P]^T X[0:p0, :]
    g.M = p1 - p0; g.N = m; g.K = p0;
    g.A = Q + p0; g.lda = ldq; g.ta = true;
    g.B = X; g.ldb = ldx;
    g.C = X + (size_t)p0 * ldx; g.ldc = ldx;
    g.D = B + (size_t)p0 * ldb; g.ldd = ldb;
    PSGD_RETURN_IF(gemm_simt(ctx, g));
    PSGD_RETURN_IF(trsm_left_block(ctx, Q, ldq, X, ldx, X, ldx, m, p0, p1));
  }
  return PSGD_OK;
}

int trsm_right_upper(psgd_ctx* ctx, const float* Q, int ldq, const float* B, int ldb, float* X, int ldx, int m,
                     int n) {
  for (int p0 = 0; p0 < n; p0 += kPanel) {
    const int p1 = p0 + kPanel < n ? p0 + kPanel : n;
    if (p0 == 0) {
      PSGD_RETURN_IF(trsm_right_block(ctx, Q, ldq, B, ldb, X, ldx, m, 0, p1));
      continue;
    }
    Gemm g;                                                // X[:,P] = B[:,P] - X[:, 0:p0] Q[0:p0, P]
    g.M = m; g.N = p1 - p0; g.K = p0;
    g.A = X; g.lda = ldx;
    g.B = Q + p0; g.ldb = ldq;
    g.C = X + p0; g.ldc = ldx;
    g.D = B + p0; g.ldd = ldb;
    PSGD_RETURN_IF(gemm_simt(ctx, g));
    PSGD_RETURN_IF(trsm_right_block(ctx, Q, ldq, X, ldx, X, ldx, m, p0, p1));
  }
  return PSGD_OK;
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
