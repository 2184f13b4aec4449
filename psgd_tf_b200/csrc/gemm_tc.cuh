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
int trsm_right_auto(psgd_ctx* ctx, const float* Q, int ldq, const float* B, int ldb, float* X, int ldx, int m, int n);
int trsm_left_auto(psgd_ctx* ctx, const float* Q, int ldq, const float* B, int ldb, float* X, int ldx, int n, int m);
// extra workspace (bytes) the tensor-core engine may carve for hi/lo operand planes of an [M,N] layer
size_t extra_ws_bytes(int64_t M, int64_t N);

}  // namespace tc
}  // namespace psgd
