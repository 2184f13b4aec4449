// Device side of the peer-memory exchange (comm.cu): slab layout and the in-kernel all-reduce, shared with the sweeps'
// "mid" kernels (uvd.cu), which run the exchange inside the launch that also reduces the CTA partials and does the
// r x r algebra, instead of as a launch of its own.
#pragma once

#include "common.cuh"

namespace psgd {
namespace comm {

constexpr int kMaxWorld = 16;
constexpr int kMaxSum = 1280;      // float64 sums per exchange (UVd rank 16: (2r)(2r+2) = 1088)
constexpr int kMaxMax = 8;         // float maxima per exchange
constexpr unsigned long long kTimeoutNs = 20ull * 1000 * 1000 * 1000;

struct SlabHead {
  unsigned long long epoch;                       // local: exchanges completed by this rank
  unsigned int error;                             // local: set when a wait timed out
  unsigned int pad[5];
};
struct Slab {
  SlabHead head;
  unsigned long long flags[2][kMaxWorld];         // [parity][source rank] = epoch of the landed contribution
  double sums[2][kMaxWorld][kMaxSum];
  float maxs[2][kMaxWorld][kMaxMax];
};

struct Peers {
  Slab* slab[kMaxWorld];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// sum_buf: n_sum float64 partial sums (in place), max_buf: n_max non-negative float maxima (in place).
// A peer that does not publish within kTimeoutNs makes the exchange FAIL, not degrade: the error is sticky in the slab
// and mirrored to a host-mapped word (every later library call that needs an exchange returns PSGD_ERR_COMM without
// a synchronisation), and this and all later exchanges hand NaNs to the consuming sweeps, so that no rank can go on
// updating U, V, d or the parameters from stale or partial sums.
__device__ __forceinline__ void exchange_device(Peers peers, int rank, int world, double* __restrict__ sum_buf,
                                                       int n_sum, float* __restrict__ max_buf, int n_max,
                                                       unsigned int* __restrict__ host_err,
                                                       unsigned long long timeout_ns) {
  Slab* mine = peers.slab[rank];
  __shared__ unsigned int s_failed;
  if (threadIdx.x == 0) s_failed = mine->head.error;
  __syncthreads();
  if (s_failed) {                               // sticky: an earlier exchange timed out
    for (int k = threadIdx.x; k < n_sum; k += blockDim.x) sum_buf[k] = __longlong_as_double(0x7ff8000000000000ll);
    for (int k = threadIdx.x; k < n_max; k += blockDim.x) max_buf[k] = __int_as_float(0x7fc00000);
    return;
  }
  const unsigned long long e = mine->head.epoch + 1;   // only this kernel writes it, and kernels of a stream are ordered
  const int par = (int)(e & 1);
  // 1. push
  for (int p = 0; p < world; ++p) {
    Slab* dst = peers.slab[p];
    for (int k = threadIdx.x; k < n_sum; k += blockDim.x) dst->sums[par][rank][k] = sum_buf[k];
    for (int k = threadIdx.x; k < n_max; k += blockDim.x) dst->maxs[par][rank][k] = max_buf[k];
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish, 3. wait
  if (threadIdx.x < world) {
    st_release_sys(&peers.slab[threadIdx.x]->flags[par][rank], e);
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(&mine->flags[par][threadIdx.x]) < e) {
      if (globaltimer_ns() - t0 > timeout_ns) {
        mine->head.error = 1;
        s_failed = 1;
        if (host_err) { *host_err = 1; __threadfence_system(); }
        break;
      }
    }
  }
  __syncthreads();
  if (s_failed) {
    for (int k = threadIdx.x; k < n_sum; k += blockDim.x) sum_buf[k] = __longlong_as_double(0x7ff8000000000000ll);
    for (int k = threadIdx.x; k < n_max; k += blockDim.x) max_buf[k] = __int_as_float(0x7fc00000);
    return;
  }
  // 4. fixed-order reduction: identical bits on every rank
  for (int k = threadIdx.x; k < n_sum; k += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += __ldcg(&mine->sums[par][r][k]);   // L2: where peer stores land
    sum_buf[k] = s;
  }
  for (int k = threadIdx.x; k < n_max; k += blockDim.x) {
    float m = 0.f;
    for (int r = 0; r < world; ++r) m = fmaxf(m, __ldcg(&mine->maxs[par][r][k]));
    max_buf[k] = m;
  }
  __syncthreads();
  if (threadIdx.x == 0) mine->head.epoch = e;
}

// Host side (comm.cu): what a kernel needs to run exchange_device itself.  world = 0 when no exchange is attached.
// Fails with PSGD_ERR_COMM when an earlier exchange timed out (sticky) or the record does not fit the slab.
struct DeviceArgs {
  Peers peers;
  int rank = 0, world = 0;
  unsigned int* host_err = nullptr;
  unsigned long long timeout_ns = kTimeoutNs;
};
int device_args(psgd_ctx* ctx, int n_sum, int n_max, DeviceArgs* out);

}  // namespace comm
}  // namespace psgd
