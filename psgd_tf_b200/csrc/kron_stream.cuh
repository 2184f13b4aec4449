// Fused bandwidth-bound kernels for the (normalization, scaling) Kronecker pair and the full-matrix GEMVs
// (kron_stream.cu).  Everything is float32 row-major on ctx->stream.
#pragma once

#include "common.cuh"

namespace psgd {
namespace ks {

int col_tiles(int N);
int row_tiles(int M);

// out[j] = sum_i w(i) X[i, j];  mode 0: w = ql1[i] / (ql0[i] ql0[M-1]) (ql = [2, M]);  mode 1: w = wvec[i].
// partial: row_tiles(M) * N floats.
int col_wsum(psgd_ctx* ctx, int mode, const float* ql, const float* wvec, const float* X, int ldx, int M, int N,
             float* partial, float* out);
// the same without the finishing launch: partial[chunk][j], *chunks_out records; the consumer sums them
int col_wsum_partials(psgd_ctx* ctx, int mode, const float* ql, const float* wvec, const float* X, int ldx, int M, int N,
                      float* partial, int* chunks_out);
// out[i] = sum_j X[i, j] w[j]
int row_dot(psgd_ctx* ctx, const float* X, int ldx, const float* w, int M, int N, float* out);

// (normalization, scaling) update statistics from dX, dG in one pass (cpart / cchunks = col_wsum_partials mode 0 of dX):
// g1d, g1b [M], grad2 [N], *max1 = max(|g1d|, |g1b|), *max2 = max|grad2| (both must be zero on entry).
// When ns_finish_is_fused(M, N) (small factors) the new factors ql_out [2, M], qr_out [1, N] (psgd.py:362-369) are written
// by the same finishing launch and max1 / max2 are not touched; otherwise the caller applies the steps.
size_t ns_update_scratch_floats(int M, int N);
bool ns_finish_is_fused(int M, int N);
int ns_update_stats(psgd_ctx* ctx, const float* ql, const float* qr, const float* cpart, int cchunks, const float* dX,
                    const float* dG, int M, int N, float* scratch, float* g1d, float* g1b, float* grad2, float* max1,
                    float* max2, float* ql_out, float* qr_out, float step, float tiny);
// (normalization, scaling) apply: out = Ql^T Ql G Qr^T Qr in one pass over G
size_t ns_apply_scratch_floats(int M, int N);
int ns_apply(psgd_ctx* ctx, const float* ql, const float* qr, const float* G, float* out, int M, int N, float* scratch);

}  // namespace ks
}  // namespace psgd
