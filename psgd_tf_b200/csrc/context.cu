// Context, workspace and error plumbing of the C ABI (include/psgd_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace psgd {
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
}  // namespace psgd

int psgd_ctx::reserve(size_t bytes) {
  if (bytes <= ws_bytes) return PSGD_OK;
  // the old block may still be in use by enqueued kernels: free it stream-ordered
  if (ws) PSGD_CUDA_CHECK(cudaFreeAsync(ws, stream));
  ws = nullptr;
  ws_bytes = 0;
  size_t want = bytes + bytes / 8;
  PSGD_CUDA_CHECK(cudaMallocAsync(&ws, want, stream));
  ws_bytes = want;
  return PSGD_OK;
}

int psgd_ctx::reserve_aux(int slot, size_t bytes, float** out) {
  if (bytes > aux_bytes[slot]) {
    if (aux[slot]) PSGD_CUDA_CHECK(cudaFreeAsync(aux[slot], stream));
    aux[slot] = nullptr;
    aux_bytes[slot] = 0;
    PSGD_CUDA_CHECK(cudaMallocAsync(&aux[slot], bytes + bytes / 8, stream));
    aux_bytes[slot] = bytes + bytes / 8;
  }
  *out = static_cast<float*>(aux[slot]);
  return PSGD_OK;
}

extern "C" int psgd_abi_version(void) { return PSGD_B200_ABI_VERSION; }

extern "C" const char* psgd_last_error(void) { return psgd::g_error; }

extern "C" int psgd_create(int device, void* stream, psgd_ctx** out) {
  PSGD_REQUIRE(out != nullptr, PSGD_ERR_BAD_POINTER, "psgd_create: out is null");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    psgd::set_error("psgd_create: no CUDA device available (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return PSGD_ERR_NO_DEVICE;
  }
  PSGD_REQUIRE(device >= 0 && device < count, PSGD_ERR_NO_DEVICE, "psgd_create: device %d out of range [0,%d)",
               device, count);
  PSGD_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  PSGD_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  PSGD_REQUIRE(prop.major == 10, PSGD_ERR_NO_DEVICE,
               "psgd_create: device %d is sm_%d%d; this build targets sm_100a (B200) only", device, prop.major,
               prop.minor);
  psgd_ctx* ctx = new psgd_ctx();
  ctx->device = device;
  ctx->stream = static_cast<cudaStream_t>(stream);
  ctx->num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("PSGD_TRSM_BASE")) {          // tuning knob, same values as psgd_set_option("trsm_base")
    const int v = atoi(e);
    if (v == 128 || v == 256 || v == 512 || v == 1024 || v == 2048) ctx->opt_trsm_base = v;
  }
  *out = ctx;
  return PSGD_OK;
}

extern "C" int psgd_destroy(psgd_ctx* ctx) {
  if (!ctx) return PSGD_OK;
  cudaSetDevice(ctx->device);
  psgd_comm_detach(ctx);
  if (ctx->ws) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->ws);
  }
  for (int k = 0; k <= psgd_ctx::kSideStreams; ++k)
    if (ctx->aux[k]) { cudaDeviceSynchronize(); cudaFree(ctx->aux[k]); }
  for (int k = 0; k < psgd_ctx::kSideStreams; ++k) {
    if (ctx->side[k]) cudaStreamDestroy(ctx->side[k]);
    if (ctx->ev_side[k]) cudaEventDestroy(ctx->ev_side[k]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  for (int k = 0; k <= psgd_ctx::kSideStreams; ++k) {
    if (ctx->branch[k]) cudaStreamDestroy(ctx->branch[k]);
    if (ctx->ev_branch_go[k]) cudaEventDestroy(ctx->ev_branch_go[k]);
    if (ctx->ev_branch_done[k]) cudaEventDestroy(ctx->ev_branch_done[k]);
  }
  delete ctx;
  return PSGD_OK;
}

extern "C" int psgd_set_stream(psgd_ctx* ctx, void* stream) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  if (ctx->stream != static_cast<cudaStream_t>(stream)) {
    // the workspace is stream-ordered: drain the old stream before switching
    PSGD_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = static_cast<cudaStream_t>(stream);
  }
  return PSGD_OK;
}

extern "C" int64_t psgd_launch_count(const psgd_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int64_t psgd_workspace_bytes(const psgd_ctx* ctx) { return ctx ? (int64_t)ctx->ws_bytes : 0; }

extern "C" int psgd_set_option(psgd_ctx* ctx, const char* key, int64_t value) {
  PSGD_REQUIRE(ctx && key, PSGD_ERR_BAD_POINTER, "null context or key");
  if (strcmp(key, "direct") == 0) { ctx->opt_direct = value ? 1 : 0; return PSGD_OK; }
  if (strcmp(key, "uvd_fused") == 0) { ctx->opt_uvd_fused = value ? 1 : 0; return PSGD_OK; }
  if (strcmp(key, "profile") == 0) { ctx->opt_profile = value == 2 ? 2 : (value ? 1 : 0); return PSGD_OK; }
  if (strcmp(key, "assume_triangular") == 0) { ctx->opt_assume_tri = value ? 1 : 0; return PSGD_OK; }
  if (strcmp(key, "tc_bn") == 0) {
    PSGD_REQUIRE(value == 128 || value == 256, PSGD_ERR_BAD_SHAPE, "tc_bn must be 128 or 256");
    ctx->opt_tc_bn = (int)value;
    return PSGD_OK;
  }
  if (strcmp(key, "trsm_base") == 0) {
    PSGD_REQUIRE(value == 128 || value == 256 || value == 512 || value == 1024 || value == 2048, PSGD_ERR_BAD_SHAPE,
                 "trsm_base must be 128, 256, 512, 1024 or 2048");
    ctx->opt_trsm_base = (int)value;
    return PSGD_OK;
  }
  if (strcmp(key, "kron_streams") == 0) { ctx->opt_kron_streams = value ? 1 : 0; return PSGD_OK; }
  if (strcmp(key, "uvd_mid") == 0) { ctx->opt_uvd_mid = value ? 1 : 0; return PSGD_OK; }
  if (strcmp(key, "comm_timeout_ms") == 0) { ctx->opt_comm_timeout_ms = value > 0 ? (int)value : 0; return PSGD_OK; }
  if (strcmp(key, "tc_debug") == 0) { ctx->opt_tc_debug = (int)value; return PSGD_OK; }
  if (strcmp(key, "stamp_ptr") == 0) { ctx->opt_stamp_ptr = reinterpret_cast<long long*>(value); return PSGD_OK; }
  if (strcmp(key, "dense_scan") == 0) { ctx->opt_dense_scan = value ? 1 : 0; return PSGD_OK; }
  if (strcmp(key, "tc_splitk") == 0) { ctx->opt_tc_splitk = value ? 1 : 0; return PSGD_OK; }
  if (strcmp(key, "tc_epi") == 0) { ctx->opt_tc_epi = value < 0 ? 0 : (value > 2 ? 2 : (int)value); return PSGD_OK; }
  if (strcmp(key, "tc_pair_sel") == 0) { ctx->opt_tc_pair_sel = (int)value; ctx->tc_launch_seq = 0; return PSGD_OK; }
  if (strcmp(key, "tc_pair") == 0) { ctx->opt_tc_pair = value ? 1 : 0; return PSGD_OK; }
  if (strcmp(key, "tc_mode") == 0) { ctx->opt_tc_mode = value ? 1 : 0; return PSGD_OK; }
  if (strcmp(key, "gemm_path") == 0) {
    PSGD_REQUIRE(value >= 0 && value <= 2, PSGD_ERR_BAD_SHAPE, "gemm_path must be 0 (auto), 1 (simt) or 2 (tcgen05)");
    ctx->opt_gemm_path = (int)value;
    return PSGD_OK;
  }
  psgd::set_error("unknown option '%s'", key);
  return PSGD_ERR_BAD_SHAPE;
}

extern "C" int psgd_set_allreduce(psgd_ctx* ctx, psgd_allreduce_fn fn, void* user) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  ctx->allreduce = fn;
  ctx->allreduce_user = user;
  return PSGD_OK;
}

extern "C" int psgd_profile_read(psgd_ctx* ctx, int* ids, float* ms, double* work, int cap) {
  if (!ctx) return 0;
  int n = 0;
  for (auto& r : ctx->prof) {
    cudaEventSynchronize(r.e1);
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    if (n < cap && ids && ms) { ids[n] = r.id; ms[n] = t; if (work) work[n] = r.work; ++n; }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  ctx->prof.clear();
  return n;
}
