"""Formulations the CUDA kernels use in place of the reference's op sequences, checked on the CPU in float64.

csrc/dense.cu: the dense preconditioner's update

    Q' = Q - mu * triu(a a^T - b b^T) Q                                   (psgd.py:40-42)

needs no matrix product -- (triu(a a^T) Q)[i, k] = a_i * sum_{j >= i} a_j Q[j, k], a suffix scan down each column of Q --
for ANY square Q, and the chunked form the kernels use (column sums per 128-row chunk, sums of the chunks below, rescan
inside the chunk) gives the same numbers.  Also the small-layer solves' formulation: a block forward substitution with
explicitly inverted 32 x 32 diagonal blocks (csrc/linalg.cu) against scipy's triangular solve."""
import numpy as np
import pytest
from scipy.linalg import solve_triangular

from oracle import psgd_oracle as O


def scan_form(Q, a, b, mu, chunk=128):
    n = Q.shape[0]
    out = np.empty_like(Q)
    chunks = [(lo, min(n, lo + chunk)) for lo in range(0, n, chunk)]
    Pa = np.stack([a[lo:hi] @ Q[lo:hi] for lo, hi in chunks])          # pass 0: column sums of every chunk
    Pb = np.stack([b[lo:hi] @ Q[lo:hi] for lo, hi in chunks])
    below_a = np.zeros_like(Pa); below_b = np.zeros_like(Pb)            # suffix_kernel: everything below the chunk
    for c in range(len(chunks) - 2, -1, -1):
        below_a[c] = below_a[c + 1] + Pa[c + 1]
        below_b[c] = below_b[c + 1] + Pb[c + 1]
    for c, (lo, hi) in enumerate(chunks):                               # pass 1: rescan the chunk bottom to top
        sa, sb = below_a[c].copy(), below_b[c].copy()
        for i in range(hi - 1, lo - 1, -1):
            sa += a[i] * Q[i]; sb += b[i] * Q[i]
            out[i] = Q[i] - mu * (a[i] * sa - b[i] * sb)
    return out


@pytest.mark.parametrize("n", [1, 2, 127, 128, 129, 300])
@pytest.mark.parametrize("full", [False, True])
def test_scan_form_equals_matrix_product(n, full):
    rng = np.random.default_rng(n + 7 * full)
    Q = rng.standard_normal((n, n))
    if not full:
        Q = np.triu(Q)
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    grad = np.triu(np.outer(a, a) - np.outer(b, b))
    mu = 0.01 / (np.abs(grad).max() + 1e-300)
    want = Q - mu * grad @ Q
    got = scan_form(Q, a, b, mu)
    assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())


def test_scan_form_reproduces_the_oracle_update():
    rng = np.random.default_rng(3)
    n = 200
    Q = (np.triu(rng.standard_normal((n, n))) * 0.1 + np.eye(n)).astype(np.float64)
    dx, dg = rng.standard_normal((n, 1)), rng.standard_normal((n, 1))
    want = O.update_precond_dense(Q, [dx], [dg], 0.01)
    a = (Q @ dg)[:, 0]
    b = solve_triangular(Q, dx, lower=False, trans="T")[:, 0]
    mx = np.abs(np.triu(np.outer(a, a) - np.outer(b, b))).max()
    got = scan_form(Q, a, b, 0.01 / (mx + O._tiny(Q)))
    assert np.abs(got - want).max() <= 1e-12


def blocked_left_solve(Q, B, nb=32):
    """X = Q^-T B by block steps with explicitly inverted diagonal blocks, as trsm_panel_kernel<LEFT> does."""
    n = Q.shape[0]
    X = np.zeros_like(B)
    for i0 in range(0, n, nb):
        i1 = min(n, i0 + nb)
        W = np.linalg.inv(np.triu(Q[i0:i1, i0:i1]))
        X[i0:i1] = W.T @ (B[i0:i1] - Q[:i0, i0:i1].T @ X[:i0])
    return X


@pytest.mark.parametrize("n,m", [(32, 5), (257, 120), (600, 33)])
def test_block_inverse_solve_matches_substitution(n, m):
    rng = np.random.default_rng(n + m)
    Q = np.triu(rng.standard_normal((n, n))) * (0.3 / np.sqrt(n)) + np.diag(0.5 + rng.random(n))
    B = rng.standard_normal((n, m))
    want = solve_triangular(Q, B, lower=False, trans="T")
    got = blocked_left_solve(Q, B)
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()


def test_scale_dense_in_its_own_orientation_equals_the_reference_mirroring():
    """(scaling, dense): the reference transposes to (dense, scaling) (psgd.py:102-104, :144-146); csrc/kron.cu runs
    A = diag(ql)(dG Qr^T), Bt = diag(1/ql)(dX Qr^-1), grad(Qr) = triu(A^T A - Bt^T Bt), grad(ql)_i = |A_i|^2 - |Bt_i|^2 and
    out = diag(ql^2) G Qr^T Qr on X itself.  Same numbers (float64)."""
    rng = np.random.default_rng(11)
    M, N = 37, 12
    ql = (0.5 + rng.random((1, M)))
    Qr = np.triu(rng.standard_normal((N, N))) * 0.2 + np.diag(0.5 + rng.random(N))
    dX, dG, G = rng.standard_normal((M, N)), rng.standard_normal((M, N)), rng.standard_normal((M, N))
    step, tiny = 0.01, O._tiny(Qr)
    want_ql, want_qr = O.update_precond_kron(ql, Qr, dX, dG, step)
    want_pre = O.precond_grad_kron(ql, Qr, G)
    # own orientation, with the reference's balancing of the transposed problem (dense factor first)
    rho = np.sqrt(np.max(np.diag(Qr)) / np.max(ql))
    qrb, qlb = Qr / rho, ql * rho
    A = qlb.T * (dG @ qrb.T)
    Bt = solve_triangular(qrb, dX.T, lower=False, trans="T").T / qlb.T          # dX Qr^-1, rows scaled
    g2 = np.triu(A.T @ A - Bt.T @ Bt)
    got_qr = qrb - step / (np.abs(g2).max() + tiny) * g2 @ qrb
    g1 = (A * A).sum(1) - (Bt * Bt).sum(1)
    got_ql = qlb - step / (np.abs(g1).max() + tiny) * g1[None, :] * qlb
    assert np.abs(got_qr - want_qr).max() <= 1e-12 and np.abs(got_ql - want_ql).max() <= 1e-12
    got_pre = (ql.T ** 2) * (G @ Qr.T @ Qr)
    assert np.abs(got_pre - want_pre).max() <= 1e-12 * np.abs(want_pre).max()


def test_dense_norm_in_its_own_orientation_equals_the_reference_mirroring():
    """(dense, normalization): the reference transposes to (normalization, dense) (psgd.py:86, :128).  Written out for X
    itself (Qr = diag(q0) with q1 in the last ROW... of its transpose, i.e. X Qr^T = X diag(q0) + X[:, -1] q1):
      T = dG diag(q0) + dG[:, -1:] q1,  A = Ql T,
      S = dX / q0 with last column  S[:, -1] -= dX @ (q1 / (q0 q0[-1])),  Bt = Ql^-T S,
      grad(Ql) = triu(A A^T - Bt Bt^T),  d_j = |A[:, j]|^2 - |Bt[:, j]|^2,  bias_j = A[:, j].A[:, -1] - Bt[:, j].Bt[:, -1]."""
    rng = np.random.default_rng(21)
    M, N = 9, 14
    Ql = np.triu(rng.standard_normal((M, M))) * 0.2 + np.diag(0.5 + rng.random(M))
    qr = np.stack([0.5 + rng.random(N), 0.1 * rng.standard_normal(N)])
    qr[1, -1] = 0.0
    dX, dG, G = rng.standard_normal((M, N)), rng.standard_normal((M, N)), rng.standard_normal((M, N))
    step, tiny = 0.01, O._tiny(Ql)
    want_ql, want_qr = O.update_precond_kron(Ql, qr, dX, dG, step)
    want_pre = O.precond_grad_kron(Ql, qr, G)
    rho = np.sqrt(np.max(qr[0]) / np.max(np.diag(Ql)))      # the transposed problem balances the normalization factor first
    q = qr / rho
    Qlb = rho * Ql
    q0, q1 = q[0], q[1]
    T = dG * q0 + dG[:, -1:] * q1
    A = Qlb @ T
    S = dX / q0
    S[:, -1] = S[:, -1] - dX @ (q1 / (q0 * q0[-1]))
    Bt = solve_triangular(Qlb, S, lower=False, trans="T")
    g1 = np.triu(A @ A.T - Bt @ Bt.T)
    got_ql = Qlb - step / (np.abs(g1).max() + tiny) * g1 @ Qlb
    d = (A * A).sum(0) - (Bt * Bt).sum(0)
    bias = A[:, :-1].T @ A[:, -1] - Bt[:, :-1].T @ Bt[:, -1]
    bias = np.concatenate([bias, np.zeros(1)])
    s = step / (max(np.abs(d).max(), np.abs(bias).max()) + tiny)
    got_qr = np.stack([q0 - s * d * q0, q1 - s * (d * q1 + q0[-1] * bias)])
    assert np.abs(got_ql - want_ql).max() <= 1e-12 and np.abs(got_qr - want_qr).max() <= 1e-12
    # apply: out = Ql^T Ql (G diag(q0) + G[:, -1] q1) then the transposed out-op
    P = G * qr[0] + G[:, -1:] * qr[1]
    P = Ql.T @ (Ql @ P)
    out = P * qr[0]
    out[:, -1] = out[:, -1] + P @ qr[1]
    assert np.abs(out - want_pre).max() <= 1e-12 * np.abs(want_pre).max()
