// Shared host/device plumbing for the PSGD B200 kernels: context, workspace, error reporting,
// PTX wrappers for mbarrier + TMA bulk copies (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/psgd_b200.h"

namespace psgd {

// ---------------------------------------------------------------------------------------------
// error reporting
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define PSGD_CUDA_CHECK(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::psgd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                        __LINE__);                                                         \
      return PSGD_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

#define PSGD_RETURN_IF(status_expr)     \
  do {                                  \
    int _s = (status_expr);             \
    if (_s != PSGD_OK) return _s;       \
  } while (0)

#define PSGD_REQUIRE(cond, code, ...)   \
  do {                                  \
    if (!(cond)) {                      \
      ::psgd::set_error(__VA_ARGS__);   \
      return (code);                    \
    }                                   \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Function attributes (the > 48 KB dynamic shared-memory opt-in) are PER DEVICE, and contexts exist per (thread,
// device): a call site remembers the devices it has configured in a bit mask (atomic: contexts of different threads
// may race here; setting an attribute twice is harmless).
struct DeviceOnce {
  std::atomic<unsigned long long> mask{0};
  bool done(int device) const { return device >= 0 && device < 64 && ((mask.load(std::memory_order_acquire) >> device) & 1ull); }
  void set(int device) { if (device >= 0 && device < 64) mask.fetch_or(1ull << device, std::memory_order_release); }
};

}  // namespace psgd

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct psgd_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 148;
  int64_t launches = 0;
  // growable device workspace (cudaMalloc'ed, owned)
  void* ws = nullptr;
  size_t ws_bytes = 0;
  // small pinned/ device scratch is carved from ws by each op
  int opt_direct = 0;        // streaming kernels: 1 = direct global loads instead of TMA pipeline
  int opt_uvd_fused = 1;     // UVd update: 1 = two sweeps + d pass (rank-2 step from the Gram table), 0 = three sweeps
  int opt_gemm_path = 0;     // dense GEMMs: 0 = auto, 1 = SIMT fp32, 2 = tcgen05 3xTF32
  int opt_tc_bn = 128;       // tcgen05 GEMM tile width (128 or 256)
  int opt_trsm_base = 1024;  // tensor-core triangular solves: width of the diagonal blocks applied via their explicit inverse
  int opt_tc_debug = 0;      // tcgen05 GEMM timing ablations (wrong results; tools/gemm_debug.py only)
  long long* opt_stamp_ptr = nullptr;   // tools only: device buffer that the panel-solve kernels fill with clock64() stamps
  int opt_dense_scan = 1;    // dense update: 1 = column scans (O(n^2)), 0 = the reference's n^3 product (cross-check)
  int opt_tc_splitk = 1;     // tcgen05 GEMM: split K over grouped problems when the output has too few tiles to fill the GPU
  int opt_tc_epi = 2;        // tcgen05 GEMM epilogue stores: 2 = staged through shared memory (full 128-byte lines), 1 = 256-bit, 0 = 128-bit per row
  int opt_tc_pair_sel = -1;  // debugging aid: >= 0 = only the sel-th tensor-core GEMM launch since the option was set uses the pair kernel
  int tc_launch_seq = 0;
  int opt_tc_pair = 0;       // tcgen05 GEMM: 1 = CTA-pair kernel (cta_group::2, 256-row tiles) where applicable, 0 = single-CTA only
  int opt_tc_mode = 1;       // tcgen05 GEMM A operand: 1 = through tensor memory (TS), 0 = from shared memory (SS)
  int opt_kron_streams = 1;  // Kron batched calls: 1 = groups of different shape run concurrently on internal side streams
  static constexpr int kSideStreams = 4;
  cudaStream_t side[kSideStreams] = {};      // created on first use (kron.cu), joined back into `stream` before a call returns
  cudaEvent_t ev_fork = nullptr, ev_side[kSideStreams] = {};
  // second stream per slot: the two independent halves of a small (dense, dense) update run side by side (kron.cu)
  cudaStream_t branch[kSideStreams + 1] = {};
  cudaEvent_t ev_branch_go[kSideStreams + 1] = {}, ev_branch_done[kSideStreams + 1] = {};
  int opt_uvd_mid = 1;       // UVd: 1 = one "mid" launch between sweeps (reduce + exchange + algebra), 0 = three launches
  int opt_comm_timeout_ms = 0;   // peer exchange: wait limit per exchange (0 = the 20 s default)
  int opt_assume_tri = 1;    // dense Kron factors are upper triangular: let GEMMs skip structurally-zero K blocks
  psgd_allreduce_fn allreduce = nullptr;
  void* allreduce_user = nullptr;
  void* comm = nullptr;      // psgd::comm::State of the peer-memory exchange (comm.cu)
  int comm_world = 0;        // > 0 once attached
  // optional per-kernel timing (psgd_set_option("profile", 1)): CUDA event pairs around the large kernels
  int opt_profile = 0;
  struct ProfRec { int id; cudaEvent_t e0, e1; double work; };
  std::vector<ProfRec> prof;

  // Ensure at least `bytes` of workspace; contents are NOT preserved across growth.
  int reserve(size_t bytes);
  // Second scratch area for the GEMM engine's own needs (split-K partials), one per stream slot so that groups running
  // concurrently on the side streams do not share it.  slot 0 = the context's stream, k = side[k - 1].
  void* aux[kSideStreams + 1] = {};
  size_t aux_bytes[kSideStreams + 1] = {};
  cudaStream_t main_stream_of_call = nullptr;      // the caller's stream while a batched call has forked (else nullptr)
  int stream_slot() const {
    for (int k = 0; k < kSideStreams; ++k)
      if (side[k] && stream == side[k]) return k + 1;
    for (int k = 1; k <= kSideStreams; ++k)
      if (branch[k] && stream == branch[k]) return k;
    return 0;
  }
  int reserve_aux(int slot, size_t bytes, float** out);
};

namespace psgd {

// Bump allocator over ctx->ws (256-byte granularity).
struct WsCarver {
  char* base;
  size_t off = 0;
  explicit WsCarver(void* b) : base(static_cast<char*>(b)) {}
  template <typename T>
  T* take(size_t count) {
    T* p = reinterpret_cast<T*>(base + off);
    off += ((count * sizeof(T) + 255) / 256) * 256;
    return p;
  }
  static size_t padded(size_t bytes) { return ((bytes + 255) / 256) * 256; }
};

// Cross-rank reduction between two kernels of a chunk-sharded sweep (comm.cu): n_sum float64 sums and n_max
// non-negative float maxima, reduced in place over all ranks -- by the peer-memory exchange kernel when attached, else
// by the registered all-reduce hook, else a single-GPU no-op.
int cross_rank_reduce(psgd_ctx* ctx, double* sum_buf, int n_sum, float* max_buf, int n_max);
static inline bool is_sharded(const psgd_ctx* ctx) { return ctx->comm_world > 0 || ctx->allreduce != nullptr; }

// Brackets one kernel launch with CUDA events when profiling is on (kernel ids: psgd_b200.h).
struct ProfScope {
  psgd_ctx* c;
  psgd_ctx::ProfRec r;
  bool on;
  // work: algorithmic bytes or flops of the launch (dense count), reported back by psgd_profile_read
  ProfScope(psgd_ctx* ctx, int id, double work = 0.0) : c(ctx), on(ctx->opt_profile != 0) {
    if (!on) return;
    r.id = id;
    r.work = work;
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, c->stream);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(r.e1, c->stream);
    c->prof.push_back(r);
  }
};

#define PSGD_LAUNCH_CHECK(ctx)                \
  do {                                        \
    (ctx)->launches += 1;                     \
    PSGD_CUDA_CHECK(cudaGetLastError());      \
  } while (0)

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// TMA bulk copy (no tensor map): global -> shared, completion counted in bytes on an mbarrier.
// Requires 16-byte aligned src/dst and size % 16 == 0.   SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (TMA store / tcgen05)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Packed fp32 FMA (sm_100+): d.x = a.x*b.x + c.x, d.y = a.y*b.y + c.y, each an IEEE round-to-nearest fma.   SASS: FFMA2.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// atomic max for non-negative floats (bit pattern order == value order); NaN poisons to NaN-ish max.
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
  atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
#endif  // __CUDACC__

}  // namespace psgd
