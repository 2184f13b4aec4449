// Diagonal (Jacobi) and X-shape preconditioners -- fused streaming kernels.
//
// The reference names these variants (README.md:11-15, :35) but ships no code for them; the math is
// the Lie-group update of every other PSGD variant specialised to the sparsity pattern (SURVEY.md
// appendix B).
//
// Each update is two sweeps because a global max gates the write:
//   sweep 1  read inputs, form the gradient on the pattern, reduce max|grad|      (no store)
//   sweep 2  re-read inputs, recompute the gradient, apply  q -= mu * grad * q     (one store per state)
// X-shape elements i and N-1-i only ever meet each other, so one thread owns the pair and the mirrored
// loads are plain descending coalesced accesses.
#include "common.cuh"

namespace psgd {
namespace ew {

constexpr int kThreads = 256;

struct Scalars {
  float max_abs;
};

__global__ void reset_kernel(Scalars* s) { s->max_abs = 0.f; }

__device__ __forceinline__ void block_max_to(float m, float* dst) {
  __shared__ float red[kThreads / 32];
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) r = fmaxf(r, red[w]);
    atomic_max_nonneg(dst, r);
  }
}

// ---- diagonal -------------------------------------------------------------------------------
__device__ __forceinline__ float diag_nabla(float q, float v, float h) {
  const float Qh = q * h;
  const float iq = v / q;
  return Qh * Qh - iq * iq;
}

template <bool WRITE>
__global__ void __launch_bounds__(kThreads) diag_kernel(float* __restrict__ q, const float* __restrict__ v,
                                                         const float* __restrict__ h, int64_t n, float step, float tiny,
                                                         Scalars* __restrict__ sc) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float mu = 0.f, m = 0.f;
  if (WRITE) mu = step / (sc->max_abs + tiny);
  const int64_t n4 = n / 4;
  float4* q4 = reinterpret_cast<float4*>(q);
  const float4* v4 = reinterpret_cast<const float4*>(v);
  const float4* h4 = reinterpret_cast<const float4*>(h);
  for (int64_t i = tid; i < n4; i += stride) {
    float4 Q = q4[i], V = v4[i], H = h4[i];
    float g0 = diag_nabla(Q.x, V.x, H.x), g1 = diag_nabla(Q.y, V.y, H.y);
    float g2 = diag_nabla(Q.z, V.z, H.z), g3 = diag_nabla(Q.w, V.w, H.w);
    if (WRITE) {
      Q.x -= mu * g0 * Q.x; Q.y -= mu * g1 * Q.y; Q.z -= mu * g2 * Q.z; Q.w -= mu * g3 * Q.w;
      q4[i] = Q;
    } else {
      m = fmaxf(m, fmaxf(fmaxf(fabsf(g0), fabsf(g1)), fmaxf(fabsf(g2), fabsf(g3))));
    }
  }
  for (int64_t i = n4 * 4 + tid; i < n; i += stride) {
    float g = diag_nabla(q[i], v[i], h[i]);
    if (WRITE) q[i] -= mu * g * q[i];
    else m = fmaxf(m, fabsf(g));
  }
  if (!WRITE) block_max_to(m, &sc->max_abs);
}

__global__ void __launch_bounds__(kThreads) diag_apply_kernel(const float* __restrict__ q, const float* __restrict__ g,
                                                               float* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n4 = n / 4;
  const float4* q4 = reinterpret_cast<const float4*>(q);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* o4 = reinterpret_cast<float4*>(out);
  for (int64_t i = tid; i < n4; i += stride) {
    float4 Q = q4[i], G = g4[i];
    o4[i] = make_float4(Q.x * Q.x * G.x, Q.y * Q.y * G.y, Q.z * Q.z * G.z, Q.w * Q.w * G.w);
  }
  for (int64_t i = n4 * 4 + tid; i < n; i += stride) out[i] = q[i] * q[i] * g[i];
}

// ---- X-shape ----------------------------------------------------------------------------------
struct XPair {
  float a_i, a_j, b_i, b_j;     // state at i and its mirror j = n-1-i
  float na_i, na_j, nb;         // gradient on the pattern (nabla_b is the same at i and j)
};

__device__ __forceinline__ void xmat_grad(XPair& p, float v_i, float v_j, float h_i, float h_j, bool centre) {
  const float Qh_i = p.a_i * h_i + p.b_i * h_j;
  const float Qh_j = p.a_j * h_j + p.b_j * h_i;
  const float det = p.a_i * p.a_j - p.b_i * p.b_j;
  const float x_i = (p.a_j * v_i - p.b_j * v_j) / det;
  const float x_j = (p.a_i * v_j - p.b_i * v_i) / det;
  p.na_i = Qh_i * Qh_i - x_i * x_i;
  p.na_j = Qh_j * Qh_j - x_j * x_j;
  p.nb = centre ? 0.f : (Qh_i * Qh_j - x_i * x_j);
}

template <bool WRITE>
__global__ void __launch_bounds__(kThreads) xmat_kernel(float* __restrict__ a, float* __restrict__ b,
                                                         const float* __restrict__ v, const float* __restrict__ h,
                                                         int64_t n, float step, float tiny, Scalars* __restrict__ sc) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t half = (n + 1) / 2;   // pairs, the centre (odd n) pairs with itself
  float mu = 0.f, m = 0.f;
  if (WRITE) mu = step / (sc->max_abs + tiny);
  for (int64_t i = tid; i < half; i += stride) {
    const int64_t j = n - 1 - i;
    XPair p;
    p.a_i = a[i]; p.a_j = a[j]; p.b_i = b[i]; p.b_j = b[j];
    xmat_grad(p, v[i], v[j], h[i], h[j], i == j);
    if (WRITE) {
      const float na_i = p.na_i, na_j = p.na_j, nb = p.nb;
      const float ai = p.a_i - mu * (na_i * p.a_i + nb * p.b_j);
      const float bi = p.b_i - mu * (na_i * p.b_i + nb * p.a_j);
      const float aj = p.a_j - mu * (na_j * p.a_j + nb * p.b_i);
      const float bj = p.b_j - mu * (na_j * p.b_j + nb * p.a_i);
      a[i] = ai; b[i] = bi;
      if (i != j) { a[j] = aj; b[j] = bj; }
    } else {
      m = fmaxf(m, fmaxf(fabsf(p.na_i), fmaxf(fabsf(p.na_j), fabsf(p.nb))));
    }
  }
  if (!WRITE) block_max_to(m, &sc->max_abs);
}

__global__ void __launch_bounds__(kThreads) xmat_apply_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                               const float* __restrict__ g, float* __restrict__ out,
                                                               int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t half = (n + 1) / 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride) {
    const int64_t j = n - 1 - i;
    const float a_i = a[i], a_j = a[j], b_i = b[i], b_j = b[j], g_i = g[i], g_j = g[j];
    const float ab_i = a_i * b_i, ab_j = a_j * b_j;
    // (a^2 + flip(b^2)) g + (ab + flip(ab)) flip(g)
    out[i] = (a_i * a_i + b_j * b_j) * g_i + (ab_i + ab_j) * g_j;
    if (i != j) out[j] = (a_j * a_j + b_i * b_i) * g_j + (ab_j + ab_i) * g_i;
  }
}

// 128-bit variants for n % 8 == 0 (16-byte aligned mirrored blocks): one thread owns elements i..i+3 and their mirror
// images n-1-i .. n-4-i, i.e. one float4 from each half of every array; component e pairs with component 3-e.
__device__ __forceinline__ void unpack(const float4& v, float (&x)[4]) { x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
__device__ __forceinline__ void unpack_rev(const float4& v, float (&x)[4]) { x[0] = v.w; x[1] = v.z; x[2] = v.y; x[3] = v.x; }
__device__ __forceinline__ float4 pack(const float (&x)[4]) { return make_float4(x[0], x[1], x[2], x[3]); }
__device__ __forceinline__ float4 pack_rev(const float (&x)[4]) { return make_float4(x[3], x[2], x[1], x[0]); }

template <bool WRITE>
__global__ void __launch_bounds__(kThreads) xmat_kernel_v4(float* __restrict__ a, float* __restrict__ b,
                                                            const float* __restrict__ v, const float* __restrict__ h,
                                                            int64_t n, float step, float tiny, Scalars* __restrict__ sc) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t groups = n / 8;        // float4 groups in the first half
  float mu = 0.f, m = 0.f;
  if (WRITE) mu = step / (sc->max_abs + tiny);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < groups; q += stride) {
    const int64_t i = 4 * q, jb = n - 4 - i;
    float ai[4], aj[4], bi[4], bj[4], vi[4], vj[4], hi[4], hj[4];
    unpack(*reinterpret_cast<const float4*>(a + i), ai); unpack_rev(*reinterpret_cast<const float4*>(a + jb), aj);
    unpack(*reinterpret_cast<const float4*>(b + i), bi); unpack_rev(*reinterpret_cast<const float4*>(b + jb), bj);
    unpack(*reinterpret_cast<const float4*>(v + i), vi); unpack_rev(*reinterpret_cast<const float4*>(v + jb), vj);
    unpack(*reinterpret_cast<const float4*>(h + i), hi); unpack_rev(*reinterpret_cast<const float4*>(h + jb), hj);
    float nai[4], naj[4], nbi[4], nbj[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      XPair p;
      p.a_i = ai[e]; p.a_j = aj[e]; p.b_i = bi[e]; p.b_j = bj[e];
      xmat_grad(p, vi[e], vj[e], hi[e], hj[e], false);
      if (WRITE) {
        nai[e] = p.a_i - mu * (p.na_i * p.a_i + p.nb * p.b_j);
        nbi[e] = p.b_i - mu * (p.na_i * p.b_i + p.nb * p.a_j);
        naj[e] = p.a_j - mu * (p.na_j * p.a_j + p.nb * p.b_i);
        nbj[e] = p.b_j - mu * (p.na_j * p.b_j + p.nb * p.a_i);
      } else {
        m = fmaxf(m, fmaxf(fabsf(p.na_i), fmaxf(fabsf(p.na_j), fabsf(p.nb))));
      }
    }
    if (WRITE) {
      *reinterpret_cast<float4*>(a + i) = pack(nai); *reinterpret_cast<float4*>(a + jb) = pack_rev(naj);
      *reinterpret_cast<float4*>(b + i) = pack(nbi); *reinterpret_cast<float4*>(b + jb) = pack_rev(nbj);
    }
  }
  if (!WRITE) block_max_to(m, &sc->max_abs);
}

__global__ void __launch_bounds__(kThreads) xmat_apply_kernel_v4(const float* __restrict__ a, const float* __restrict__ b,
                                                                  const float* __restrict__ g, float* __restrict__ out,
                                                                  int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t groups = n / 8;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < groups; q += stride) {
    const int64_t i = 4 * q, jb = n - 4 - i;
    float ai[4], aj[4], bi[4], bj[4], gi[4], gj[4], oi[4], oj[4];
    unpack(*reinterpret_cast<const float4*>(a + i), ai); unpack_rev(*reinterpret_cast<const float4*>(a + jb), aj);
    unpack(*reinterpret_cast<const float4*>(b + i), bi); unpack_rev(*reinterpret_cast<const float4*>(b + jb), bj);
    unpack(*reinterpret_cast<const float4*>(g + i), gi); unpack_rev(*reinterpret_cast<const float4*>(g + jb), gj);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float ab_i = ai[e] * bi[e], ab_j = aj[e] * bj[e];
      oi[e] = (ai[e] * ai[e] + bj[e] * bj[e]) * gi[e] + (ab_i + ab_j) * gj[e];
      oj[e] = (aj[e] * aj[e] + bi[e] * bi[e]) * gj[e] + (ab_j + ab_i) * gi[e];
    }
    *reinterpret_cast<float4*>(out + i) = pack(oi);
    *reinterpret_cast<float4*>(out + jb) = pack_rev(oj);
  }
}

static int grid_for(const psgd_ctx* ctx, int64_t work_items) {
  int64_t blocks = (work_items + kThreads - 1) / kThreads;
  int64_t cap = (int64_t)ctx->num_sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace ew
}  // namespace psgd

using namespace psgd;

static int check_ptrs(const char* what, const void* const* ptrs, int count) {
  for (int i = 0; i < count; ++i) {
    PSGD_REQUIRE(ptrs[i] != nullptr, PSGD_ERR_BAD_POINTER, "%s: null device pointer (arg %d)", what, i);
    PSGD_REQUIRE(aligned16(ptrs[i]), PSGD_ERR_BAD_POINTER, "%s: device pointer %d is not 16-byte aligned", what, i);
  }
  return PSGD_OK;
}

extern "C" int psgd_diag_update(psgd_ctx* ctx, float* q, const float* v, const float* h, int64_t n, float step,
                                float tiny) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_REQUIRE(n >= 0, PSGD_ERR_BAD_SHAPE, "diag update: n=%lld", (long long)n);
  if (n == 0 && !is_sharded(ctx)) return PSGD_OK;      // an empty shard still takes part in the max exchange
  const void* ptrs[] = {q, v, h};
  if (n > 0) PSGD_RETURN_IF(check_ptrs("diag update", ptrs, 3));
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  PSGD_RETURN_IF(ctx->reserve(256));
  ew::Scalars* sc = static_cast<ew::Scalars*>(ctx->ws);
  const int grid = ew::grid_for(ctx, (n + 3) / 4);
  ew::reset_kernel<<<1, 1, 0, ctx->stream>>>(sc);
  PSGD_LAUNCH_CHECK(ctx);
  ew::diag_kernel<false><<<grid, ew::kThreads, 0, ctx->stream>>>(q, v, h, n, step, tiny, sc);
  PSGD_LAUNCH_CHECK(ctx);
  PSGD_RETURN_IF(cross_rank_reduce(ctx, nullptr, 0, &sc->max_abs, 1));
  ew::diag_kernel<true><<<grid, ew::kThreads, 0, ctx->stream>>>(q, v, h, n, step, tiny, sc);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

extern "C" int psgd_diag_apply(psgd_ctx* ctx, const float* q, const float* g, float* out, int64_t n) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_REQUIRE(n >= 0, PSGD_ERR_BAD_SHAPE, "diag apply: n=%lld", (long long)n);
  if (n == 0) return PSGD_OK;
  const void* ptrs[] = {q, g, out};
  PSGD_RETURN_IF(check_ptrs("diag apply", ptrs, 3));
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  ew::diag_apply_kernel<<<ew::grid_for(ctx, (n + 3) / 4), ew::kThreads, 0, ctx->stream>>>(q, g, out, n);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

extern "C" int psgd_xmat_update(psgd_ctx* ctx, float* a, float* b, const float* v, const float* h, int64_t n,
                                float step, float tiny) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_REQUIRE(n >= 0, PSGD_ERR_BAD_SHAPE, "xmat update: n=%lld", (long long)n);
  if (n == 0 && !is_sharded(ctx)) return PSGD_OK;
  const void* ptrs[] = {a, b, v, h};
  if (n > 0) PSGD_RETURN_IF(check_ptrs("xmat update", ptrs, 4));
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  PSGD_RETURN_IF(ctx->reserve(256));
  ew::Scalars* sc = static_cast<ew::Scalars*>(ctx->ws);
  const bool v4 = n > 0 && (n % 8) == 0;                 // mirrored float4 blocks are 16-byte aligned
  const int grid = ew::grid_for(ctx, v4 ? n / 8 : (n + 1) / 2);
  ew::reset_kernel<<<1, 1, 0, ctx->stream>>>(sc);
  PSGD_LAUNCH_CHECK(ctx);
  if (v4) ew::xmat_kernel_v4<false><<<grid, ew::kThreads, 0, ctx->stream>>>(a, b, v, h, n, step, tiny, sc);
  else ew::xmat_kernel<false><<<grid, ew::kThreads, 0, ctx->stream>>>(a, b, v, h, n, step, tiny, sc);
  PSGD_LAUNCH_CHECK(ctx);
  PSGD_RETURN_IF(cross_rank_reduce(ctx, nullptr, 0, &sc->max_abs, 1));
  if (v4) ew::xmat_kernel_v4<true><<<grid, ew::kThreads, 0, ctx->stream>>>(a, b, v, h, n, step, tiny, sc);
  else ew::xmat_kernel<true><<<grid, ew::kThreads, 0, ctx->stream>>>(a, b, v, h, n, step, tiny, sc);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}

extern "C" int psgd_xmat_apply(psgd_ctx* ctx, const float* a, const float* b, const float* g, float* out, int64_t n) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_REQUIRE(n >= 0, PSGD_ERR_BAD_SHAPE, "xmat apply: n=%lld", (long long)n);
  if (n == 0) return PSGD_OK;
  const void* ptrs[] = {a, b, g, out};
  PSGD_RETURN_IF(check_ptrs("xmat apply", ptrs, 4));
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  if ((n % 8) == 0)
    ew::xmat_apply_kernel_v4<<<ew::grid_for(ctx, n / 8), ew::kThreads, 0, ctx->stream>>>(a, b, g, out, n);
  else
    ew::xmat_apply_kernel<<<ew::grid_for(ctx, (n + 1) / 2), ew::kThreads, 0, ctx->stream>>>(a, b, g, out, n);
  PSGD_LAUNCH_CHECK(ctx);
  return PSGD_OK;
}
