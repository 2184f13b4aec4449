// Small device linear-algebra vocabulary shared by the Kronecker / dense preconditioner paths.
// Everything is float32, row-major with explicit leading dimensions, enqueued on ctx->stream.
//
// Two GEMM engines sit behind the same descriptor:
//   * SIMT fp32 (this header + linalg.cu): exact fp32 FMA accumulation, any shape/alignment; the
//     engine for small/ragged layers (LeNet, NMT factors) and the cross-check for the tensor path;
//   * tcgen05 3xTF32 (gemm_tc.cu): TMA-staged, TMEM-accumulated split-precision GEMM for large layers.
#pragma once

#include "common.cuh"

namespace psgd {
namespace la {

// C = epilogue( op(A) op(B)  -  op(A2) op(B2) )
struct Gemm {
  int M = 0, N = 0, K = 0;
  const float* A = nullptr; int lda = 0; bool ta = false;   // op(A): [M,K]
  const float* B = nullptr; int ldb = 0; bool tb = false;   // op(B): [K,N]
  int K2 = 0;                                               // optional second product, subtracted
  const float* A2 = nullptr; int lda2 = 0; bool ta2 = false;
  const float* B2 = nullptr; int ldb2 = 0; bool tb2 = false;
  float* C = nullptr; int ldc = 0;
  // epilogue
  bool triu = false;              // zero the strictly lower triangle (tf.linalg.band_part(., 0, -1))
  float* maxabs = nullptr;        // atomic max of |C| (after masking) -- must be zeroed by the caller
  const float* D = nullptr; int ldd = 0;     // if set: C = D - mu * acc
  bool d_tri = false;             // D is the same upper-triangular factor as op(B) (Q' = Q - mu grad Q): with b_full's
                                  // blessing, tiles below the diagonal are zeros and D is not read for them
  const float* mu_max = nullptr;  // mu = step / (*mu_max + tiny)
  float step = 0.f, tiny = 0.f;
  // K-range hints (product 0 only; the engines may skip structurally-zero K blocks, never required for correctness
  // of dense data): 0 none, 1 = op(.) upper triangular, 2 = lower triangular, viewed as op(A)[M,K] / op(B)[K,N]
  int a_tri = 0, b_tri = 0;
  // Run-time validity of the hints: device flags written by kron::factor_scan (non-zero = the operand has entries
  // outside its assumed triangle, e.g. a caller-supplied factor that is not upper triangular).  When set, the
  // tensor-core engine drops the corresponding hint for that problem and multiplies the full matrix, which is what
  // tf.matmul does (psgd.py:173).  nullptr = the hint is valid by construction.
  const int* a_full = nullptr;
  const int* b_full = nullptr;
  // block-pair pattern (tensor-core engine only; triangular block-inverse doubling): only output tiles in blocks
  // (k, k+1), k even, of size pair_b exist; K range = the block of the column (kind 1) or of the row (kind 2)
  int pair_b = 0, pair_kind = 0;
  bool negate = false;               // C = -acc
  const float* colscale = nullptr;   // acc *= colscale[n]   (or its reciprocal)
  bool colscale_recip = false;
  bool colscale_sq = false;          // use colscale[n]^2
  // final scaling of the stored value by the balancing factor rho (psgd.py:166-170 folded into the last product of the
  // update instead of materialising Ql/rho and rho*Qr): 1 = C / *rho, 2 = C * *rho
  const float* rho = nullptr;
  int rho_mode = 0;
};

int gemm_simt(psgd_ctx* ctx, const Gemm& g);

// X = Q^-T B  (tf.linalg.triangular_solve(Q, B, lower=False, adjoint=True)): Q [n,n] upper, B,X [n,m].
// Reads only the upper triangle of Q.  X may alias B.
int trsm_left_upper_adjoint(psgd_ctx* ctx, const float* Q, int ldq, const float* B, int ldb, float* X, int ldx,
                            int n, int m);
// X = B Q^-1 : Q [n,n] upper, B,X [m,n].  (== transpose of the above applied to B^T.)  X may alias B.
int trsm_right_upper(psgd_ctx* ctx, const float* Q, int ldq, const float* B, int ldb, float* X, int ldx, int m,
                     int n);

// x = Q^-T b for ONE right-hand side (contiguous vectors), n of any size; ws: trsv_ws_floats(n) floats of workspace.
size_t trsv_ws_floats(int n);
int trsv_left_upper_adjoint(psgd_ctx* ctx, const float* Q, int ldq, const float* b, float* x, int n, float* ws);

// Panel forms: solve only rows [ib0, ib1) / columns [jb0, jb1) (at most 512 of them), assuming the contribution of
// earlier rows / columns has already been subtracted from B.  One launch each.
int trsm_left_block(psgd_ctx* ctx, const float* Q, int ldq, const float* B, int ldb, float* X, int ldx, int m, int ib0,
                    int ib1);
int trsm_right_block(psgd_ctx* ctx, const float* Q, int ldq, const float* B, int ldb, float* X, int ldx, int m, int jb0,
                     int jb1);

// out[c, r] = in[r, c]
int transpose(psgd_ctx* ctx, const float* in, int ld_in, float* out, int ld_out, int rows, int cols);

}  // namespace la
}  // namespace psgd
