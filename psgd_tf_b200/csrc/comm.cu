// Peer-memory exchange for the chunk-sharded streaming paths (UVd / diagonal / X-shape) on one NVSwitch box.
//
// The only cross-GPU traffic of those paths is a few hundred partial sums and one or two maxima per dependency
// phase -- latency bound, not bandwidth bound.  Instead of a host-driven collective per
// phase, every rank owns a small "slab" of device memory that all peers map through CUDA IPC; ONE tiny kernel per
// phase
//   1. pushes this rank's partials straight into every peer's slab with plain stores over NVLink (P2P),
//   2. publishes a per-(phase parity, source rank) epoch flag with a system-scope release,
//   3. spins on its own slab's flags (system-scope acquire) until every rank's contribution for this epoch landed,
//   4. reduces the `world` contributions in fixed rank order (float64 sums / float maxima),
// so all ranks obtain bit-identical results, nothing returns to the host between the sweeps, and the whole
// update+apply step is CUDA-graph capturable (the epoch counter lives in device memory).
//
// Two slot banks (epoch parity) suffice: a rank can only start pushing epoch e+2 after it consumed epoch e+1, which
// needs every peer's e+1 push, which each peer issues only after it finished consuming epoch e.
#include <string.h>

#include "comm.cuh"

namespace psgd {
namespace comm {

__global__ void __launch_bounds__(256) exchange_kernel(Peers peers, int rank, int world, double* __restrict__ sum_buf,
                                                       int n_sum, float* __restrict__ max_buf, int n_max,
                                                       unsigned int* __restrict__ host_err, unsigned long long timeout_ns) {
  exchange_device(peers, rank, world, sum_buf, n_sum, max_buf, n_max, host_err, timeout_ns);
}

struct State {
  int rank = 0, world = 0;
  Slab* local = nullptr;
  Peers peers{};
  unsigned int* host_err = nullptr;       // host-mapped (cudaHostAlloc): set by the kernel when a wait timed out
  unsigned int* host_err_dev = nullptr;   // its device alias
};

int device_args(psgd_ctx* ctx, int n_sum, int n_max, DeviceArgs* out) {
  *out = DeviceArgs{};
  if (ctx->comm_world <= 0) return PSGD_OK;
  auto* st = static_cast<State*>(ctx->comm);
  PSGD_REQUIRE(n_sum <= kMaxSum && n_max <= kMaxMax, PSGD_ERR_COMM,
               "peer exchange: %d sums / %d maxima exceed the slab (%d / %d)", n_sum, n_max, kMaxSum, kMaxMax);
  PSGD_REQUIRE(!st->host_err || *static_cast<volatile unsigned int*>(st->host_err) == 0, PSGD_ERR_COMM,
               "peer exchange: an earlier exchange timed out (a rank did not publish its partials within %llu s); "
               "the sharded state is invalid",
               kTimeoutNs / 1000000000ull);
  out->peers = st->peers;
  out->rank = st->rank;
  out->world = st->world;
  out->host_err = st->host_err_dev;
  out->timeout_ns = ctx->opt_comm_timeout_ms > 0 ? (unsigned long long)ctx->opt_comm_timeout_ms * 1000000ull : kTimeoutNs;
  return PSGD_OK;
}

}  // namespace comm

// Cross-rank reduction between two kernels of a sharded sweep: peer-memory exchange when attached, else the
// registered all-reduce hook (torch.distributed / NCCL / gloo), else single-GPU no-op.
int cross_rank_reduce(psgd_ctx* ctx, double* sum_buf, int n_sum, float* max_buf, int n_max) {
  if (ctx->comm_world > 0) {
    auto* st = static_cast<comm::State*>(ctx->comm);
    PSGD_REQUIRE(n_sum <= comm::kMaxSum && n_max <= comm::kMaxMax, PSGD_ERR_COMM,
                 "peer exchange: %d sums / %d maxima exceed the slab (%d / %d)", n_sum, n_max, comm::kMaxSum, comm::kMaxMax);
    PSGD_REQUIRE(!st->host_err || *static_cast<volatile unsigned int*>(st->host_err) == 0, PSGD_ERR_COMM,
                 "peer exchange: an earlier exchange timed out (a rank did not publish its partials within %llu s); "
                 "the sharded state is invalid",
                 comm::kTimeoutNs / 1000000000ull);
    ProfScope prof(ctx, PSGD_K_EXCHANGE);
    comm::exchange_kernel<<<1, 256, 0, ctx->stream>>>(st->peers, st->rank, st->world, sum_buf, n_sum, max_buf, n_max,
                                                      st->host_err_dev,
                                                      ctx->opt_comm_timeout_ms > 0 ? (unsigned long long)ctx->opt_comm_timeout_ms * 1000000ull
                                                                                   : comm::kTimeoutNs);
    PSGD_LAUNCH_CHECK(ctx);
    return PSGD_OK;
  }
  if (!ctx->allreduce) return PSGD_OK;
  if (n_sum > 0) {
    int rc = ctx->allreduce(ctx->allreduce_user, sum_buf, n_sum, 0, (void*)ctx->stream);
    PSGD_REQUIRE(rc == 0, PSGD_ERR_COMM, "all-reduce hook returned %d", rc);
  }
  if (n_max > 0) {
    int rc = ctx->allreduce(ctx->allreduce_user, max_buf, n_max, 1, (void*)ctx->stream);
    PSGD_REQUIRE(rc == 0, PSGD_ERR_COMM, "all-reduce hook returned %d", rc);
  }
  return PSGD_OK;
}

}  // namespace psgd

using namespace psgd;

static_assert(sizeof(cudaIpcMemHandle_t) == PSGD_COMM_HANDLE_BYTES, "cudaIpcMemHandle_t size");

extern "C" int psgd_comm_export(psgd_ctx* ctx, void* handle_out) {
  PSGD_REQUIRE(ctx && handle_out, PSGD_ERR_BAD_POINTER, "psgd_comm_export: null argument");
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  if (!ctx->comm) ctx->comm = new comm::State();
  auto* st = static_cast<comm::State*>(ctx->comm);
  PSGD_REQUIRE(st->world == 0, PSGD_ERR_COMM, "psgd_comm_export: already attached; detach first");
  if (!st->local) {
    // plain cudaMalloc: memory from the stream-ordered pool cannot be exported through legacy CUDA IPC
    PSGD_CUDA_CHECK(cudaMalloc(&st->local, sizeof(comm::Slab)));
  }
  PSGD_CUDA_CHECK(cudaMemset(st->local, 0, sizeof(comm::Slab)));
  if (!st->host_err) {
    PSGD_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&st->host_err), sizeof(unsigned int), cudaHostAllocMapped));
    PSGD_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&st->host_err_dev), st->host_err, 0));
  }
  *st->host_err = 0;
  PSGD_CUDA_CHECK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  PSGD_CUDA_CHECK(cudaIpcGetMemHandle(&h, st->local));
  memcpy(handle_out, &h, sizeof(h));
  return PSGD_OK;
}

extern "C" int psgd_comm_attach(psgd_ctx* ctx, int rank, int world, const void* handles) {
  PSGD_REQUIRE(ctx && handles, PSGD_ERR_BAD_POINTER, "psgd_comm_attach: null argument");
  PSGD_REQUIRE(world >= 1 && world <= comm::kMaxWorld && rank >= 0 && rank < world, PSGD_ERR_BAD_SHAPE,
               "psgd_comm_attach: rank %d of %d (max %d ranks)", rank, world, comm::kMaxWorld);
  auto* st = static_cast<comm::State*>(ctx->comm);
  PSGD_REQUIRE(st && st->local, PSGD_ERR_COMM, "psgd_comm_attach: call psgd_comm_export first");
  PSGD_REQUIRE(st->world == 0, PSGD_ERR_COMM, "psgd_comm_attach: already attached");
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  const char* hs = static_cast<const char*>(handles);
  for (int p = 0; p < world; ++p) {
    if (p == rank) { st->peers.slab[p] = st->local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, hs + (size_t)p * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int q = 0; q < p; ++q)
        if (q != rank) cudaIpcCloseMemHandle(st->peers.slab[q]);
      set_error("psgd_comm_attach: cudaIpcOpenMemHandle(rank %d) failed: %s", p, cudaGetErrorString(e));
      (void)cudaGetLastError();
      return PSGD_ERR_COMM;
    }
    st->peers.slab[p] = static_cast<comm::Slab*>(ptr);
  }
  st->rank = rank;
  st->world = world;
  ctx->comm_world = world;
  return PSGD_OK;
}

extern "C" int psgd_comm_detach(psgd_ctx* ctx) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  auto* st = static_cast<comm::State*>(ctx->comm);
  if (!st) return PSGD_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int p = 0; p < st->world; ++p)
    if (p != st->rank && st->peers.slab[p]) cudaIpcCloseMemHandle(st->peers.slab[p]);
  if (st->local) cudaFree(st->local);
  if (st->host_err) cudaFreeHost(st->host_err);
  delete st;
  ctx->comm = nullptr;
  ctx->comm_world = 0;
  return PSGD_OK;
}

extern "C" int psgd_comm_status(psgd_ctx* ctx, int64_t* epoch_out) {
  PSGD_REQUIRE(ctx, PSGD_ERR_BAD_POINTER, "null context");
  auto* st = static_cast<comm::State*>(ctx->comm);
  PSGD_REQUIRE(st && st->world > 0, PSGD_ERR_COMM, "psgd_comm_status: not attached");
  PSGD_CUDA_CHECK(cudaSetDevice(ctx->device));
  PSGD_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  comm::SlabHead head;
  PSGD_CUDA_CHECK(cudaMemcpy(&head, st->local, sizeof(head), cudaMemcpyDeviceToHost));
  if (epoch_out) *epoch_out = (int64_t)head.epoch;
  PSGD_REQUIRE(head.error == 0, PSGD_ERR_COMM, "peer exchange: a rank did not publish its partials within %llu s",
               comm::kTimeoutNs / 1000000000ull);
  return PSGD_OK;
}
