// tcgen05 3xTF32 engine of the device linear-algebra vocabulary (see linalg.cuh) + the size-based
// engine choice used by the Kronecker paths.
#pragma once

#include "linalg.cuh"

namespace psgd {
namespace tc {

// Picks the engine per call: ctx->opt_gemm_path 0 = auto (tensor cores when every dimension is large
// enough to fill 128-wide tiles), 1 = SIMT fp32, 2 = tcgen05 whenever the shape is supported.
int gemm_auto(psgd_ctx* ctx, const la::Gemm& g);
// tcgen05 engine directly; operands must be 16-byte aligned with leading dimensions that are multiples of 4
bool gemm_tc_supported(const la::Gemm& g);
int gemm_tc(psgd_ctx* ctx, const la::Gemm& g);
// The same GEMM for many layers: tensor-core eligible problems that share shape/flags go out as grouped launches
// (one kernel for up to 24 problems); the rest run one by one through gemm_auto.  force_tc: use the tensor-core
// engine whenever the operands allow it, regardless of size (TRSM recursion).
int gemm_many(psgd_ctx* ctx, const la::Gemm* gs, int count, bool force_tc);

// Grouped triangular solves.  zinv: per-problem scratch of trsm_scratch_floats(n) floats (explicit inverses of the
// diagonal base blocks + doubling scratch); work: per-problem scratch shaped like X (working copy of the right-hand
// side); B may alias X.
struct Trsm {
  const float* Q;
  const float* B;
  float* X;
  float* zinv;
  float* work;
};
size_t trsm_scratch_floats(int n);
// X = B Q^-1 : Q [n,n] upper, B,X [m,n]
int trsm_right_many(psgd_ctx* ctx, const Trsm* ts, int count, int ldq, int ldb, int ldx, int m, int n);
// X = Q^-T B : Q [n,n] upper, B,X [n,m]   (tf.linalg.triangular_solve(Q, B, lower=False, adjoint=True))
int trsm_left_many(psgd_ctx* ctx, const Trsm* ts, int count, int ldq, int ldb, int ldx, int n, int m);

}  // namespace tc
}  // namespace psgd
