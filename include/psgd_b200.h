/*
 * psgd_b200.h -- C ABI of the B200-native PSGD preconditioner hot path.
 *
 * Drop-in boundary for lixilinx/psgd_tf's `preconditioned_stochastic_gradient_descent.py`
 * (called psgd.py below; citations are file:line in the reference tree).  The reference has no
 * FFI of its own: its operator interface is a set of Python functions whose arithmetic is delegated
 * to TensorFlow ops.  Each entry point here replaces the TensorFlow op sequence of one such function;
 * the Python mirror in psgd_tf_b200/ binds them with ctypes (see INTEGRATION.md for the stub a
 * reference maintainer would add).
 *
 * Rules of the ABI
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer to float32, row-major,
 *     contiguous, 16-byte aligned (what DLPack hands over for a compact tensor);
 *   - every call enqueues work on the context's CUDA stream and returns without a host sync;
 *   - every call returns 0 on success or a psgd_status code; psgd_last_error() gives the
 *     thread-local message;
 *   - there is NO CPU fallback: a context cannot be created without a CUDA device.
 */
#ifndef PSGD_B200_H_
#define PSGD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSGD_B200_ABI_VERSION 1

typedef enum psgd_status {
  PSGD_OK = 0,
  PSGD_ERR_BAD_SHAPE = 1,     /* sizes inconsistent / unsupported rank                    */
  PSGD_ERR_BAD_POINTER = 2,   /* null or misaligned device pointer                         */
  PSGD_ERR_UNSUPPORTED = 3,   /* unknown Kron factor combination (psgd.py:89-91 semantics) */
  PSGD_ERR_CUDA = 4,          /* a CUDA runtime call failed                                */
  PSGD_ERR_COMM = 5,          /* the registered all-reduce hook failed                     */
  PSGD_ERR_NO_DEVICE = 6      /* no CUDA device: there is no CPU path                      */
} psgd_status;

/* Kronecker factor formats (README.md:37-39 of the reference; psgd.py:80-110). */
typedef enum psgd_factor_kind {
  PSGD_FACTOR_DENSE = 0, /* [N,N] upper-triangular Cholesky-like factor */
  PSGD_FACTOR_SCALE = 1, /* [1,N] diagonal                               */
  PSGD_FACTOR_NORM = 2   /* [2,N] diagonal + last column                 */
} psgd_factor_kind;

typedef struct psgd_ctx psgd_ctx;

/* ---- context ------------------------------------------------------------------------------ */
int psgd_abi_version(void);
const char* psgd_last_error(void);
/* stream: a cudaStream_t (0 = legacy default stream). */
int psgd_create(int device, void* stream, psgd_ctx** out);
int psgd_destroy(psgd_ctx* ctx);
int psgd_set_stream(psgd_ctx* ctx, void* stream);
/* Number of kernels this library has launched through ctx since creation (bench "gpu_launches"). */
int64_t psgd_launch_count(const psgd_ctx* ctx);
/* Bytes of device workspace currently owned by ctx. */
int64_t psgd_workspace_bytes(const psgd_ctx* ctx);
/* Options.  "direct": streaming kernels 0 = TMA bulk-copy pipeline (default), 1 = direct global loads (debug
 * cross-check; same arithmetic).  "uvd_fused": psgd_uvd_update 1 = two sweeps + a pass over d, the rank-2 step's
 * coefficients taken from the Gram table (default), 0 = three sweeps with direct reductions over a, b (cross-check).
 * "profile": see below.  "dense_scan": psgd_dense_update 1 = the O(n^2) column-scan form of Q - mu triu(a a^T - b b^T) Q
 * (default), 0 = the reference's n^3 matrix product (cross-check).  Others ("gemm_path", "tc_bn", "trsm_base",
 * "assume_triangular", ...) tune the dense Kron engine. */
int psgd_set_option(psgd_ctx* ctx, const char* key, int64_t value);

/* Per-kernel device timing.  After psgd_set_option(ctx, "profile", 1) every large kernel launch is bracketed
 * by CUDA events on ctx's stream; psgd_profile_read synchronises on them, writes up to `cap` (kernel id,
 * milliseconds, work) records in launch order, clears the log and returns the number written.  `work` (may be
 * NULL) is the launch's algorithmic work: dense-count flops for GEMM/TRSM launches (with "profile" = 2: the flops the
 * tcgen05 GEMM launch actually executes after triangular K clipping and tile skipping), 0 where the caller knows the
 * byte count from the shapes (UVd sweeps).  Kernel ids: */
#define PSGD_K_UVD_GRAM_UPDATE 1 /* update sweep 1: Gram/vector reductions over U,V,d,h,v  */
#define PSGD_K_UVD_MAP_UPDATE2 2 /* update sweep 2: per-row a,b,nablaD + max/sums            */
#define PSGD_K_UVD_MAP_UPDATE3 3 /* update sweep 3: write d and U (or V)                      */
#define PSGD_K_UVD_GRAM_APPLY 4  /* apply sweep 1: U^T U, U^T(dg), V^T(dg)                    */
#define PSGD_K_UVD_MAP_APPLY 5   /* apply sweep 2: write the preconditioned gradient          */
#define PSGD_K_EXCHANGE 6        /* one peer-memory exchange (push partials to all ranks, wait, reduce)  */
#define PSGD_K_UVD_MAP_FUSED 7   /* fused update sweep 2: a,b,nablaD per row + rank-2 update of U (or V) */
#define PSGD_K_UVD_D_UPDATE 8    /* fused update pass 3: d -= mu_d d nablaD                              */
#define PSGD_K_UVD_MAP_UPDAPP 9  /* update+apply sweep 2: PSGD_K_UVD_MAP_FUSED + Gram sums of the updated factors */
#define PSGD_K_UVD_MID 14        /* between two sweeps: partial reduction + peer exchange + r x r algebra in one launch */
#define PSGD_K_UVD_MAP_APPLY_D 13 /* update+apply sweep 3: d update fused with the apply's map sweep     */
#define PSGD_K_GEMM 10           /* one tcgen05 3xTF32 GEMM launch                            */
#define PSGD_K_GEMM_SIMT 12      /* one SIMT fp32 GEMM launch                                 */
#define PSGD_K_TRSM 11           /* one triangular-solve step                                 */
#define PSGD_K_DENSE_SCAN 15     /* dense update: |grad| maximum + the two column-scan passes that write Q' (work = bytes) */
int psgd_profile_read(psgd_ctx* ctx, int* ids, float* ms, double* work, int cap);

/* Cross-rank reduction hook for the chunk-sharded (multi-GPU) streaming paths.  When set, the
 * library calls it on ctx's stream between kernels with a device buffer of `count` float64
 * (op 0 = sum) or float32 (op 1 = max) values that must be all-reduced in place over all ranks.
 * Return 0 on success.  With no hook the library runs single-GPU. */
typedef int (*psgd_allreduce_fn)(void* user, void* device_buf, int64_t count, int op, void* stream);
int psgd_set_allreduce(psgd_ctx* ctx, psgd_allreduce_fn fn, void* user);

/* Peer-memory exchange: the B200-native alternative to the hook on one NVSwitch box (one process per GPU).  Each
 * rank exports a small device slab through CUDA IPC, gathers every rank's 64-byte handle with whatever host plumbing
 * it has (torch.distributed.all_gather_object in psgd_tf_b200/partition.py) and attaches.  From then on the
 * cross-rank reductions of the sharded paths are ONE tiny kernel per phase that stores the partials straight into the
 * peers' slabs over NVLink, waits on epoch flags and reduces in fixed rank order (bit-identical on every rank, no host
 * round trip, CUDA-graph capturable).  Takes precedence over the hook while attached.
 *   psgd_comm_export : allocate/zero the local slab, write its cudaIpcMemHandle_t (PSGD_COMM_HANDLE_BYTES) to handle_out
 *   psgd_comm_attach : handles = world x PSGD_COMM_HANDLE_BYTES, in rank order (this rank's own entry is ignored)
 *   psgd_comm_status : synchronises ctx's stream; error if a peer failed to publish within the in-kernel timeout;
 *                      epoch_out (may be NULL) = exchanges completed so far                                         */
#define PSGD_COMM_HANDLE_BYTES 64
int psgd_comm_export(psgd_ctx* ctx, void* handle_out);
int psgd_comm_attach(psgd_ctx* ctx, int rank, int world, const void* handles);
int psgd_comm_detach(psgd_ctx* ctx);
int psgd_comm_status(psgd_ctx* ctx, int64_t* epoch_out);

/* ---- UVd: Q = (I + U V^T) diag(d) ---------------------------------------------------------- */
/* Replaces update_precond_UVd_math_ (psgd.py:554-617).  U,V:[n,r]  d,v,h:[n].  In place on U,V,d.
 * The reference's two coin flips are arguments: balance = (uniform < 0.01) (psgd.py:562),
 * update_U = (uniform < 0.5) (psgd.py:588). */
int psgd_uvd_update(psgd_ctx* ctx, float* U, float* V, float* d, const float* v, const float* h,
                    int64_t n, int r, float step, float tiny, int balance, int update_U);
/* Replaces precond_grad_UVd_math (psgd.py:619-627): out = d*(I+VU^T)(I+UV^T)(d*g). */
int psgd_uvd_apply(psgd_ctx* ctx, const float* U, const float* V, const float* d, const float* g,
                   float* out, int64_t n, int r);
/* psgd_uvd_update followed by psgd_uvd_apply on the updated state (the sequence UVd.step runs, psgd.py:732-748), fused
 * into three sweeps: the apply's Gram sums of the UPDATED factors are accumulated in the sweep that writes them and the
 * d update rides in the apply's map sweep, so U and V are each read three times per step instead of five.  Same
 * results as the two calls.  g, out: [n]; out must not alias d or g. */
int psgd_uvd_update_apply(psgd_ctx* ctx, float* U, float* V, float* d, const float* v, const float* h, const float* g,
                          float* out, int64_t n, int r, float step, float tiny, int balance, int update_U);
/* Tail of UVd.step (psgd.py:747-762) on the flattened parameter vector: pre = precond_grad_UVd_math(U,V,d,g);
 * lr = lr_params * min(grad_clip_max_norm / (||pre||_2 + tiny), 1) (psgd.py:750-754; pass INFINITY for "no clipping",
 * then lr = lr_params); param -= lr * pre (+ v when v != NULL: the finite-difference perturbation, psgd.py:760-762).
 * Without clipping the parameter update is fused into the second apply sweep and pre is only written if pre_out is
 * given; with clipping the norm is reduced on the device (all-reduced when sharded) -- never a host sync.
 * v and pre_out may be NULL. */
int psgd_uvd_step_tail(psgd_ctx* ctx, const float* U, const float* V, const float* d, const float* g, float* param,
                       const float* v, float* pre_out, int64_t n, int r, float lr_params, float grad_clip_max_norm,
                       float tiny);
/* Pipeline geometry the library compiled for the update Gram sweep at rank r (rows per shared-memory stage, stages,
 * threads per CTA, warp roles) -- diagnostics and documentation only. */
int psgd_uvd_plan_info(int r, int* tile_rows, int* stages, int* threads, int* roles);
/* Replaces IpUVtmatvec (psgd.py:540-544): out = x + U (V^T x), x:[n,k] row-major. */
int psgd_ipuvt_matvec(psgd_ctx* ctx, const float* U, const float* V, const float* x, float* out,
                      int64_t n, int r, int k);

/* ---- diagonal and X-shape (README.md:11-15, :35; no reference code -- SURVEY.md appendix B) - */
int psgd_diag_update(psgd_ctx* ctx, float* q, const float* v, const float* h, int64_t n, float step, float tiny);
int psgd_diag_apply(psgd_ctx* ctx, const float* q, const float* g, float* out, int64_t n);
int psgd_xmat_update(psgd_ctx* ctx, float* a, float* b, const float* v, const float* h, int64_t n,
                     float step, float tiny);
int psgd_xmat_apply(psgd_ctx* ctx, const float* a, const float* b, const float* g, float* out, int64_t n);

/* ---- dense full-matrix preconditioner ------------------------------------------------------ */
/* Replaces update_precond_dense (psgd.py:26-42) on the already concatenated vectors dx, dg:[n]. */
int psgd_dense_update(psgd_ctx* ctx, const float* Q, const float* dx, const float* dg, float* Q_out,
                      int64_t n, float step, float tiny);
/* Replaces precond_grad_dense (psgd.py:45-63): out = Q^T (Q g). */
int psgd_dense_apply(psgd_ctx* ctx, const float* Q, const float* g, float* out, int64_t n);

/* ---- sparse-LU preconditioner Q = L U, L = [L1 0; L2 diag(l3)], U = [U1 U2; 0 diag(u3)] --------------------------- */
/* Replaces update_precond_splu (psgd.py:396-477) on the already concatenated dx, dg: [n].  L12 = [L1; L2]: [n, r],
 * U12 = [U1 U2]: [r, n], l3, u3: [n - r]; 1 <= r <= 32.  Functional: results in the *_out buffers (same shapes). */
int psgd_splu_update(psgd_ctx* ctx, const float* L12, const float* l3, const float* U12, const float* u3,
                     const float* dx, const float* dg, float* L12_out, float* l3_out, float* U12_out, float* u3_out,
                     int64_t n, int r, float step, float tiny);
/* Replaces precond_grad_splu (psgd.py:483-524): out = U^T L^T L U g, g, out: [n]. */
int psgd_splu_apply(psgd_ctx* ctx, const float* L12, const float* l3, const float* U12, const float* u3,
                    const float* g, float* out, int64_t n, int r);

/* ---- building block ------------------------------------------------------------------------ */
/* C[M,N] = op(A) op(B), row-major fp32 (ta/tb: transpose flags as in tf.matmul(transpose_a, transpose_b)), through
 * engine 0 = auto, 1 = SIMT fp32, 2 = tcgen05 3xTF32.  triu: zero the strictly lower triangle of C.  a_tri/b_tri:
 * 0 none, 1 = op(A)/op(B) is upper triangular, 2 = lower (lets the engine skip structurally-zero K blocks).
 * This is what every tf.matmul call site of psgd.py lowers to; exported for engine cross-checks. */
int psgd_gemm(psgd_ctx* ctx, int engine, int M, int N, int K, const float* A, int lda, int ta, const float* B,
              int ldb, int tb, float* C, int ldc, int triu, int a_tri, int b_tri);

/* ---- Kronecker-product preconditioners ----------------------------------------------------- */
/* One layer.  dX,dG,G,out are [M,N].  A DENSE left factor is [M,M], SCALE [1,M], NORM [2,M];
 * right factors likewise with N.  All seven combinations the reference dispatches
 * (psgd.py:82-110, :124-152) are accepted; (NORM,NORM), (SCALE,SCALE) return PSGD_ERR_UNSUPPORTED.
 * Functional: inputs untouched, results written to Ql_out / Qr_out (same shapes as Ql / Qr). */
int psgd_kron_update(psgd_ctx* ctx, int kind_l, int kind_r, const float* Ql, const float* Qr,
                     const float* dX, const float* dG, float* Ql_out, float* Qr_out, int64_t M, int64_t N,
                     float step, float tiny);
int psgd_kron_apply(psgd_ctx* ctx, int kind_l, int kind_r, const float* Ql, const float* Qr,
                    const float* G, float* out, int64_t M, int64_t N);

/* A ragged list of layers in one call ("batched launch over all layers").  Arrays are HOST arrays
 * of length count. */
typedef struct psgd_kron_layer {
  int32_t kind_l, kind_r;
  int64_t M, N;
  const float* Ql;
  const float* Qr;
  const float* dX; /* update only */
  const float* dG; /* update only */
  const float* G;  /* apply only  */
  float* Ql_out;   /* update only */
  float* Qr_out;   /* update only */
  float* out;      /* apply only  */
} psgd_kron_layer;
int psgd_kron_update_batched(psgd_ctx* ctx, const psgd_kron_layer* layers, int count, float step, float tiny);
int psgd_kron_apply_batched(psgd_ctx* ctx, const psgd_kron_layer* layers, int count);

/* ---- caller-side tail of a Kron training step, over a ragged list of layers (HOST arrays of length count) ---------- */
/* grad_norm = sqrt(sum_l sum(pre_l^2)); lr_adjust = min(grad_norm_clip_thr / grad_norm, 1) (mnist_with_lenet5.py:54-55;
 * pass INFINITY for no clipping); W_l -= lr_adjust * lr * pre_l (+ v_l when v != NULL: the finite-difference perturbation,
 * neural_machine_translation_with_attention.py:206).  One launch per phase for the whole list, norm reduced on the
 * device (all-reduced across ranks when sharded). */
typedef struct psgd_param_update {
  float* W;          /* parameters, updated in place   */
  const float* pre;  /* preconditioned gradient        */
  const float* v;    /* perturbation to remove or NULL */
  int64_t count;     /* elements                       */
} psgd_param_update;
int psgd_apply_updates(psgd_ctx* ctx, const psgd_param_update* items, int count, float lr, float grad_norm_clip_thr);
/* out_l = a_l - b_l for a list: the finite-difference Hessian-vector products dG = perturbed_g - g
 * (neural_machine_translation_with_attention.py:200). */
typedef struct psgd_diff_item {
  const float* a;
  const float* b;
  float* out;
  int64_t count;
} psgd_diff_item;
int psgd_multi_sub(psgd_ctx* ctx, const psgd_diff_item* items, int count);

#ifdef __cplusplus
}
#endif
#endif /* PSGD_B200_H_ */
