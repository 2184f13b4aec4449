// Caller-side pieces of a Kronecker-preconditioned training step that sit directly around the hot path and are pure
// streaming work over a RAGGED LIST of layers:
//
//   psgd_apply_updates   grad_norm = sqrt(sum_l sum(pre_l^2)); lr_adjust = min(clip_thr / grad_norm, 1);
//                        W_l -= lr_adjust * lr * pre_l (+ v_l)            mnist_with_lenet5.py:54-56,
//                                                                         neural_machine_translation_with_attention.py:206, :233
//   psgd_multi_sub       dG_l = perturbed_g_l - g_l                       neural_machine_translation_with_attention.py:200
//
// The reference issues one TensorFlow op per layer per line; here the whole list is one launch per phase ("multi-tensor"
// kernels: blockIdx.y selects the layer, blockIdx.x strides its elements), the norm is reduced on the device in fixed
// order (float64 final stage, all-reduced across ranks when the layers are sharded) and never visits the host.
#include <math.h>

#include "common.cuh"

namespace psgd {
namespace multi {

constexpr int kMaxItems = 64;        // layers per launch (kernel parameter space); longer lists are chunked
constexpr int kBlocksPerItem = 32;
constexpr int kThreads = 256;

struct Items {
  float* w[kMaxItems];
  const float* a[kMaxItems];
  const float* b[kMaxItems];
  long long count[kMaxItems];
};

// partial[(item0 + y) * kBlocksPerItem + x] = sum over this block's slice of a^2
__global__ void __launch_bounds__(kThreads) sumsq_kernel(Items it, int item0, float* __restrict__ partial) {
  __shared__ float red[kThreads / 32];
  const float* a = it.a[blockIdx.y];
  const long long n = it.count[blockIdx.y];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s = fmaf(a[i], a[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += red[w];
    partial[(size_t)(item0 + blockIdx.y) * kBlocksPerItem + blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ partial, int count, double* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < count; i += 256) s += (double)partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

// w -= lr_adjust * lr * a (+ b),  lr_adjust = min(clip / sqrt(sumsq), 1) when sumsq != nullptr
__global__ void __launch_bounds__(kThreads) update_kernel(Items it, float lr, float clip, const double* __restrict__ sumsq) {
  float scale = lr;
  if (sumsq) scale = fminf(clip / sqrtf((float)sumsq[0]), 1.0f) * lr;          // mnist_with_lenet5.py:54-55
  float* w = it.w[blockIdx.y];
  const float* a = it.a[blockIdx.y];
  const float* b = it.b[blockIdx.y];
  const long long n = it.count[blockIdx.y];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float u = scale * a[i];
    if (b) u = u + b[i];                                                        // attention.py:206: lr*g + v
    w[i] = w[i] - u;
  }
}

__global__ void __launch_bounds__(kThreads) sub_kernel(Items it) {
  float* w = it.w[blockIdx.y];
  const float* a = it.a[blockIdx.y];
  const float* b = it.b[blockIdx.y];
  const long long n = it.count[blockIdx.y];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) w[i] = a[i] - b[i];
}

static int blocks_for(const psgd_ctx* ctx, long long max_count) {
  long long b = (max_count + kThreads * 4 - 1) / (kThreads * 4);
  const long long cap = (long long)ctx->num_sms * 4;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace multi
}  // namespace psgd

using namespace psgd;

extern "C" int psgd_apply_updates(psgd_ctx* ctx, const psgd_param_update* items, int count, float lr,
                                  float grad_norm_clip_thr) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_REQUIRE(count >= 0 && (items || count == 0), PSGD_ERR_BAD_POINTER, "psgd_apply_updates: null item list");
  PSGD_REQUIRE(grad_norm_clip_thr > 0.f, PSGD_ERR_BAD_SHAPE, "psgd_apply_updates: clip threshold must be > 0 (INFINITY = none)");
  for (int i = 0; i < count; ++i)
    PSGD_REQUIRE(items[i].count >= 0 && (items[i].count == 0 || (items[i].W && items[i].pre)), PSGD_ERR_BAD_POINTER,
                 "psgd_apply_updates: null pointer in item %d", i);
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  const bool clip = !isinf(grad_norm_clip_thr);
  if (count == 0 && !(clip && is_sharded(ctx))) return PSGD_OK;
  double* sumsq = nullptr;
  if (clip) {
    const size_t np = (size_t)(count > 0 ? count : 1) * multi::kBlocksPerItem;
    PSGD_RETURN_IF(ctx->reserve(WsCarver::padded(np * sizeof(float)) + 512));
    WsCarver c(ctx->ws);
    float* partial = c.take<float>(np);
    sumsq = c.take<double>(1);
    for (int i0 = 0; i0 < count; i0 += multi::kMaxItems) {
      const int m = count - i0 < multi::kMaxItems ? count - i0 : multi::kMaxItems;
      multi::Items it{};
      for (int i = 0; i < m; ++i) { it.a[i] = items[i0 + i].pre; it.count[i] = items[i0 + i].count; }
      multi::sumsq_kernel<<<dim3(multi::kBlocksPerItem, m), multi::kThreads, 0, ctx->stream>>>(it, i0, partial);
      PSGD_LAUNCH_CHECK(ctx);
    }
    multi::sum_partials_kernel<<<1, 256, 0, ctx->stream>>>(partial, count * multi::kBlocksPerItem, sumsq);
    PSGD_LAUNCH_CHECK(ctx);
    PSGD_RETURN_IF(cross_rank_reduce(ctx, sumsq, 1, nullptr, 0));      // layer-sharded stacks: the norm spans all ranks
  }
  for (int i0 = 0; i0 < count; i0 += multi::kMaxItems) {
    const int m = count - i0 < multi::kMaxItems ? count - i0 : multi::kMaxItems;
    multi::Items it{};
    long long mx = 0;
    for (int i = 0; i < m; ++i) {
      it.w[i] = items[i0 + i].W; it.a[i] = items[i0 + i].pre; it.b[i] = items[i0 + i].v; it.count[i] = items[i0 + i].count;
      if (it.count[i] > mx) mx = it.count[i];
    }
    multi::update_kernel<<<dim3(multi::blocks_for(ctx, mx), m), multi::kThreads, 0, ctx->stream>>>(it, lr, grad_norm_clip_thr, sumsq);
    PSGD_LAUNCH_CHECK(ctx);
  }
  return PSGD_OK;
}

extern "C" int psgd_multi_sub(psgd_ctx* ctx, const psgd_diff_item* items, int count) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  PSGD_REQUIRE(count >= 0 && (items || count == 0), PSGD_ERR_BAD_POINTER, "psgd_multi_sub: null item list");
  for (int i = 0; i < count; ++i)
    PSGD_REQUIRE(items[i].count >= 0 && (items[i].count == 0 || (items[i].a && items[i].b && items[i].out)),
                 PSGD_ERR_BAD_POINTER, "psgd_multi_sub: null pointer in item %d", i);
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  for (int i0 = 0; i0 < count; i0 += multi::kMaxItems) {
    const int m = count - i0 < multi::kMaxItems ? count - i0 : multi::kMaxItems;
    multi::Items it{};
    long long mx = 0;
    for (int i = 0; i < m; ++i) {
      it.w[i] = items[i0 + i].out; it.a[i] = items[i0 + i].a; it.b[i] = items[i0 + i].b; it.count[i] = items[i0 + i].count;
      if (it.count[i] > mx) mx = it.count[i];
    }
    multi::sub_kernel<<<dim3(multi::blocks_for(ctx, mx), m), multi::kThreads, 0, ctx->stream>>>(it);
    PSGD_LAUNCH_CHECK(ctx);
  }
  return PSGD_OK;
}
