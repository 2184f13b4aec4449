// tcgen05 3xTF32 GEMM engine -- engine selection.  (Tensor-core kernels land here; until a shape is
// covered by them the SIMT fp32 engine runs it.)
#include "gemm_tc.cuh"

namespace psgd {
namespace tc {

int gemm_auto(psgd_ctx* ctx, const la::Gemm& g) { return la::gemm_simt(ctx, g); }

int trsm_right_auto(psgd_ctx* ctx, const float* Q, int ldq, const float* B, int ldb, float* X, int ldx, int m, int n) {
  return la::trsm_right_upper(ctx, Q, ldq, B, ldb, X, ldx, m, n);
}
int trsm_left_auto(psgd_ctx* ctx, const float* Q, int ldq, const float* B, int ldb, float* X, int ldx, int n, int m) {
  return la::trsm_left_upper_adjoint(ctx, Q, ldq, B, ldb, X, ldx, n, m);
}
size_t extra_ws_bytes(int64_t, int64_t) { return 0; }

}  // namespace tc
}  // namespace psgd
