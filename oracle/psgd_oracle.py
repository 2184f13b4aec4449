"""CPU oracle for the PSGD preconditioner hot path -- TEST INFRASTRUCTURE ONLY.

This is an op-for-op NumPy restatement of the reference module
``/root/reference/preconditioned_stochastic_gradient_descent.py`` (called
``psgd.py`` below).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it (those legs time its multi-threaded
torch-CPU twin, ``psgd_oracle_torch.py``, which ``tests/test_oracle_torch.py`` checks against this file); the
product package ``psgd_tf_b200`` never does.

Pin status.  PINNED AT SOURCE LEVEL, UNPINNED AT TENSORFLOW-KERNEL LEVEL: the reference ships no tests
or golden vectors and TensorFlow cannot be installed in this image (no network).  The restatement is pinned by
(0) ``tests/golden/reference_outputs.npz`` -- outputs of the reference's UNMODIFIED source file executed in this
container on a NumPy stand-in for the TensorFlow ops it calls (``tests/golden/make_reference_golden.py``,
``tests/golden/tf_numpy_shim``); this oracle reproduces all of them bit for bit (tests/test_reference_golden.py),
so dispatch, operation order and association are the reference's; (1) line-by-line citations on every function;
(2) the algebraic invariants of SURVEY.md section 4 (tests/test_oracle.py); (3) dense-densification cross checks of
every structured variant; (4) a float64 twin (``dtype=np.float64``) that bounds the float32 rounding of the oracle
itself.  What remains unpinned is the rounding of TensorFlow's own CPU kernels (Eigen contraction order), which
the 1e-5 tolerance absorbs.  The diagonal / X-shape functions have no reference code at all (spec-derived).

Every function takes ``dtype`` implicitly from its inputs: feed float32 arrays to
mirror the reference (psgd.py:20), float64 arrays for the twin.

Conventions that differ from the reference on purpose:
  * the two internal coin flips of ``update_precond_UVd_math_`` (psgd.py:562,
    psgd.py:588) are explicit arguments ``balance`` / ``update_U`` because TF's
    Philox stream cannot be reproduced without TF;
  * the UVd update returns the new (U, V, d) instead of assigning in place.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import solve_triangular as _solve_triangular

# psgd.py:20-22 -- smallest *normal* float32 (TF flushes denormals, so the halving
# loop of the reference stops at 2**-126; a NumPy loop would reach 2**-149).
TINY = np.float32(2.0 ** -126)

__all__ = [
    "TINY", "update_precond_dense", "precond_grad_dense", "update_precond_kron",
    "precond_grad_kron", "IpUVtmatvec", "update_precond_UVd_math", "precond_grad_UVd_math", "uvd_step_tail",
    "update_precond_Xmat", "precond_grad_Xmat", "update_precond_diag", "precond_grad_diag",
    "update_precond_splu", "precond_grad_splu",
]


def _tiny(x):
    return x.dtype.type(TINY)


def _triu_solve_adjoint(Q, B):
    """tf.linalg.triangular_solve(Q, B, lower=False, adjoint=True): Q^T X = B, reading
    only the upper triangle of Q (psgd.py:39, :174, :233, :298)."""
    if B.size == 0:
        return B.copy()
    return _solve_triangular(Q, B, lower=False, trans="T", check_finite=False).astype(Q.dtype, copy=False)


def _triu(X):
    """tf.linalg.band_part(X, 0, -1)."""
    return np.triu(X)


def _max_abs(X):
    return np.max(np.abs(X)) if X.size else X.dtype.type(0)


# ----------------------------------------------------------------------------------------
# dense (full matrix) preconditioner                                      psgd.py:26-63
# ----------------------------------------------------------------------------------------
def update_precond_dense(Q, dxs, dgs, step=0.01):
    """psgd.py:26-42."""
    t = Q.dtype.type
    dx = np.concatenate([np.reshape(x, [-1, 1]) for x in dxs], 0)          # :34
    dg = np.concatenate([np.reshape(g, [-1, 1]) for g in dgs], 0)          # :35
    a = Q @ dg                                                               # :38
    b = _triu_solve_adjoint(Q, dx)                                           # :39
    grad = _triu(a @ a.T - b @ b.T)                                          # :40
    step0 = t(step) / (_max_abs(grad) + _tiny(Q))                            # :41
    return Q - (step0 * grad) @ Q                                            # :42


def precond_grad_dense(Q, grads):
    """psgd.py:45-63."""
    grad = [np.reshape(g, [-1, 1]) for g in grads]                           # :51
    lens = [g.shape[0] for g in grad]                                        # :52
    grad = np.concatenate(grad, 0)                                           # :53
    pre_grad = Q.T @ (Q @ grad)                                              # :55
    pre_grads, idx = [], 0
    for i in range(len(grads)):                                              # :59-61
        pre_grads.append(np.reshape(pre_grad[idx: idx + lens[i]], np.shape(grads[i])))
        idx += lens[i]
    return pre_grads


# ----------------------------------------------------------------------------------------
# Kronecker-product preconditioners                                       psgd.py:67-391
# ----------------------------------------------------------------------------------------
class UnknownKron(Exception):
    """Raised internally; the public wrappers mirror the reference's print-and-passthrough."""


def update_precond_kron(Ql, Qr, dX, dG, step=0.01):
    """Shape dispatch of psgd.py:72-110 (square => dense is tested first)."""
    m, n = Ql.shape
    p, q = Qr.shape
    if m == n:
        if p == q:
            return _update_precond_dense_dense(Ql, Qr, dX, dG, step)                     # :84
        elif p == 2:
            return _update_precond_norm_dense(Qr, Ql, dX.T, dG.T, step)[::-1]            # :86
        elif p == 1:
            return _update_precond_dense_scale(Ql, Qr, dX, dG, step)                     # :88
    elif m == 2:
        if p == q:
            return _update_precond_norm_dense(Ql, Qr, dX, dG, step)                      # :94
        elif p == 1:
            return _update_precond_norm_scale(Ql, Qr, dX, dG, step)                      # :96
    elif m == 1:
        if p == q:
            return _update_precond_dense_scale(Qr, Ql, dX.T, dG.T, step)[::-1]           # :102
        elif p == 2:
            return _update_precond_norm_scale(Qr, Ql, dX.T, dG.T, step)[::-1]            # :104
    print("Unknown Kronecker product preconditioner, no update")                         # :90 ...
    return Ql, Qr


def precond_grad_kron(Ql, Qr, Grad):
    """Shape dispatch of psgd.py:116-152."""
    m, n = Ql.shape
    p, q = Qr.shape
    if m == n:
        if p == q:
            return _precond_grad_dense_dense(Ql, Qr, Grad)                               # :126
        elif p == 2:
            return _precond_grad_norm_dense(Qr, Ql, Grad.T).T                            # :128
        elif p == 1:
            return _precond_grad_dense_scale(Ql, Qr, Grad)                               # :130
    elif m == 2:
        if p == q:
            return _precond_grad_norm_dense(Ql, Qr, Grad)                                # :136
        elif p == 1:
            return _precond_grad_norm_scale(Ql, Qr, Grad)                                # :138
    elif m == 1:
        if p == q:
            return _precond_grad_dense_scale(Qr, Ql, Grad.T).T                           # :144
        elif p == 2:
            return _precond_grad_norm_scale(Qr, Ql, Grad.T).T                            # :146
    print("Unknown Kronecker product preconditioner, no preconditioning")                # :132 ...
    return Grad


def _update_precond_dense_dense(Ql, Qr, dX, dG, step=0.01):
    """psgd.py:156-179."""
    t = Ql.dtype.type
    max_l = np.max(np.diag(Ql))                                              # :166
    max_r = np.max(np.diag(Qr))                                              # :167
    rho = np.sqrt(max_l / max_r)                                             # :168
    Ql = Ql / rho                                                            # :169
    Qr = rho * Qr                                                            # :170
    A = Ql @ (dG @ Qr.T)                                                     # :173
    Bt = _triu_solve_adjoint(Ql, _triu_solve_adjoint(Qr, dX.T).T)            # :174
    grad1 = _triu(A @ A.T - Bt @ Bt.T)                                       # :175
    grad2 = _triu(A.T @ A - Bt.T @ Bt)                                       # :176
    step1 = t(step) / (_max_abs(grad1) + _tiny(Ql))                          # :177
    step2 = t(step) / (_max_abs(grad2) + _tiny(Ql))                          # :178
    return Ql - (step1 * grad1) @ Ql, Qr - (step2 * grad2) @ Qr              # :179


def _precond_grad_dense_dense(Ql, Qr, Grad):
    """psgd.py:182-192 (association switches on M < N)."""
    if Grad.shape[0] < Grad.shape[1]:
        return (((Ql.T @ Ql) @ Grad) @ Qr.T) @ Qr                            # :190
    else:
        return Ql.T @ (Ql @ (Grad @ (Qr.T @ Qr)))                            # :192


def _norm_left_products(ql, dX, dG):
    """Ql*dG and Ql^(-T)*dX for the [2, M] normalization format (psgd.py:218-219, :230-232)."""
    A = ql[0:1].T * dG                                                       # :218
    A = A + ql[1:].T @ dG[-1:]                                               # :219
    Bt = (1.0 / ql[0:1]).T.astype(ql.dtype) * dX                             # :230
    Bt = np.concatenate([Bt[:-1],
                         Bt[-1:] - (ql[1:] / (ql[0:1] * ql[0, -1])) @ dX], axis=0)   # :231-232
    return A, Bt


def _norm_left_update(ql, A, Bt, step):
    """grad1_diag / grad1_bias / step1 / new ql rows (psgd.py:235-241)."""
    t = ql.dtype.type
    grad1_diag = np.sum(A * A, axis=1) - np.sum(Bt * Bt, axis=1)             # :235
    grad1_bias = A[:-1] @ A[-1:].T - Bt[:-1] @ Bt[-1:].T                     # :236
    grad1_bias = np.concatenate([np.reshape(grad1_bias, [-1]), np.zeros(1, ql.dtype)], axis=0)  # :237
    step1 = t(step) / (np.maximum(_max_abs(grad1_diag), _max_abs(grad1_bias)) + _tiny(ql))      # :239
    new_ql0 = ql[0] - step1 * grad1_diag * ql[0]                             # :240
    new_ql1 = ql[1] - step1 * (grad1_diag * ql[1] + ql[0, -1] * grad1_bias)  # :241
    return np.stack((new_ql0, new_ql1))


def _update_precond_norm_dense(ql, Qr, dX, dG, step=0.01):
    """psgd.py:198-246."""
    t = ql.dtype.type
    max_l = np.max(ql[0])                                                    # :211
    max_r = np.max(np.diag(Qr))                                              # :212
    rho = np.sqrt(max_l / max_r)                                             # :213
    ql = ql / rho                                                            # :214
    Qr = rho * Qr                                                            # :215
    A, Bt = _norm_left_products(ql, dX, dG)                                  # :218-232
    A = A @ Qr.T                                                             # :220
    Bt = _triu_solve_adjoint(Qr, Bt.T).T                                     # :233
    new_ql = _norm_left_update(ql, A, Bt, step)                              # :235-241
    grad2 = _triu(A.T @ A - Bt.T @ Bt)                                       # :243
    step2 = t(step) / (_max_abs(grad2) + _tiny(ql))                          # :244
    return new_ql, Qr - (step2 * grad2) @ Qr                                 # :246


def _norm_left_apply_in(ql, Grad):
    preG = ql[0:1].T * Grad                                                  # :258 / :383
    return preG + ql[1:].T @ Grad[-1:]                                       # :259 / :384


def _norm_left_apply_out(ql, preG):
    add_last_row = ql[1:] @ preG                                             # :265 / :386
    preG = ql[0:1].T * preG                                                  # :266 / :387
    return np.concatenate([preG[:-1], preG[-1:] + add_last_row], axis=0)     # :267-268 / :388-389


def _precond_grad_norm_dense(ql, Qr, Grad):
    """psgd.py:249-270."""
    preG = _norm_left_apply_in(ql, Grad)
    if preG.shape[0] < preG.shape[1]:
        preG = (preG @ Qr.T) @ Qr                                            # :261
    else:
        preG = preG @ (Qr.T @ Qr)                                            # :263
    return _norm_left_apply_out(ql, preG)


def _update_precond_dense_scale(Ql, qr, dX, dG, step=0.01):
    """psgd.py:276-307."""
    t = Ql.dtype.type
    max_l = np.max(np.diag(Ql))                                              # :288
    max_r = np.max(qr)                                                       # :289
    rho = np.sqrt(max_l / max_r)                                             # :290
    Ql = Ql / rho                                                            # :291
    qr = rho * qr                                                            # :292
    A = Ql @ dG                                                              # :295
    A = A * qr                                                               # :296
    Bt = _triu_solve_adjoint(Ql, dX)                                         # :298
    Bt = Bt * (1.0 / qr).astype(Ql.dtype)                                    # :299
    grad1 = _triu(A @ A.T - Bt @ Bt.T)                                       # :301
    step1 = t(step) / (_max_abs(grad1) + _tiny(Ql))                          # :302
    grad2 = np.sum(A * A, axis=0, keepdims=True) - np.sum(Bt * Bt, axis=0, keepdims=True)   # :304
    step2 = t(step) / (_max_abs(grad2) + _tiny(Ql))                          # :305
    return Ql - (step1 * grad1) @ Ql, qr - step2 * grad2 * qr                # :307


def _precond_grad_dense_scale(Ql, qr, Grad):
    """psgd.py:310-322."""
    if Grad.shape[0] < Grad.shape[1]:
        preG = (Ql.T @ Ql) @ Grad                                            # :319
    else:
        preG = Ql.T @ (Ql @ Grad)                                            # :321
    return preG * (qr * qr)                                                  # :322


def _update_precond_norm_scale(ql, qr, dX, dG, step=0.01):
    """psgd.py:328-369."""
    t = ql.dtype.type
    max_l = np.max(ql[0])                                                    # :342
    max_r = np.max(qr)                                                       # :343
    rho = np.sqrt(max_l / max_r)                                             # :344
    ql = ql / rho                                                            # :345
    qr = rho * qr                                                            # :346
    A, Bt = _norm_left_products(ql, dX, dG)                                  # :349-350, :353-355
    A = A * qr                                                               # :351
    Bt = Bt * (1.0 / qr).astype(ql.dtype)                                    # :356
    new_ql = _norm_left_update(ql, A, Bt, step)                              # :358-364
    grad2 = np.sum(A * A, axis=0, keepdims=True) - np.sum(Bt * Bt, axis=0, keepdims=True)   # :366
    step2 = t(step) / (_max_abs(grad2) + _tiny(ql))                          # :367
    return new_ql, qr - step2 * grad2 * qr                                   # :369


def _precond_grad_norm_scale(ql, qr, Grad):
    """psgd.py:372-391."""
    preG = _norm_left_apply_in(ql, Grad)
    preG = preG * (qr * qr)                                                  # :385
    return _norm_left_apply_out(ql, preG)


# ----------------------------------------------------------------------------------------
# sparse LU preconditioner Q = L U                                        psgd.py:396-524
# ----------------------------------------------------------------------------------------
def _tri_solve(T, B, lower, adjoint):
    """tf.linalg.triangular_solve(T, B, lower=..., adjoint=...)."""
    if B.size == 0:
        return B.copy()
    return _solve_triangular(T, B, lower=lower, trans="T" if adjoint else "N", check_finite=False).astype(T.dtype, copy=False)


def update_precond_splu(L12, l3, U12, u3, dxs, dgs, step=0.01):
    """psgd.py:396-477.  L = [L1 0; L2 diag(l3)], U = [U1 U2; 0 diag(u3)]; returns (L12', l3', U12', u3')."""
    t = L12.dtype.type
    max_l = np.maximum(np.max(np.diag(L12)), np.max(l3))                     # :411
    max_u = np.maximum(np.max(np.diag(U12)), np.max(u3))                     # :412
    rho = np.sqrt(max_l / max_u)                                             # :413
    L12 = L12 / rho; l3 = l3 / rho; U12 = rho * U12; u3 = rho * u3           # :414-417
    r = U12.shape[0]                                                         # :420
    L1, L2, U1, U2 = L12[:r], L12[r:], U12[:, :r], U12[:, r:]                # :421-424
    dx = np.concatenate([np.reshape(x, [-1, 1]) for x in dxs], 0)            # :426
    dg = np.concatenate([np.reshape(g, [-1, 1]) for g in dgs], 0)            # :427
    Ug1 = U1 @ dg[:r] + U2 @ dg[r:]                                          # :430
    Ug2 = u3 * dg[r:]                                                        # :431
    Qg1 = L1 @ Ug1                                                           # :433
    Qg2 = L2 @ Ug1 + l3 * Ug2                                                # :434
    iUtx1 = _tri_solve(U1, dx[:r], lower=False, adjoint=True)                # :436
    iUtx2 = (dx[r:] - U2.T @ iUtx1) / u3                                     # :437
    iQtx2 = iUtx2 / l3                                                       # :439
    iQtx1 = _tri_solve(L1, iUtx1 - L2.T @ iQtx2, lower=True, adjoint=True)   # :440
    LtQg1 = L1.T @ Qg1 + L2.T @ Qg2                                          # :442
    LtQg2 = l3 * Qg2                                                         # :443
    Pg1 = U1.T @ LtQg1                                                       # :445
    Pg2 = U2.T @ LtQg1 + u3 * LtQg2                                          # :446
    iLiQtx1 = _tri_solve(L1, iQtx1, lower=True, adjoint=False)               # :448
    iLiQtx2 = (iQtx2 - L2 @ iLiQtx1) / l3                                    # :449
    iPx2 = iLiQtx2 / u3                                                      # :451
    iPx1 = _tri_solve(U1, iLiQtx1 - U2 @ iPx2, lower=False, adjoint=False)   # :452
    grad1 = np.tril(Qg1 @ Qg1.T - iQtx1 @ iQtx1.T)                           # :455-456
    grad2 = Qg2 @ Qg1.T - iQtx2 @ iQtx1.T                                    # :457
    grad3 = Qg2 * Qg2 - iQtx2 * iQtx2                                        # :458
    max_abs_grad = np.maximum(np.maximum(_max_abs(grad1), _max_abs(grad2)), _max_abs(grad3))   # :459-461
    step0 = t(step) / (max_abs_grad + _tiny(L12))                            # :462
    newL1 = L1 - (step0 * grad1) @ L1                                        # :463
    newL2 = L2 - (step0 * grad2) @ L1 - step0 * grad3 * L2                   # :464
    newl3 = l3 - step0 * grad3 * l3                                          # :465
    grad1 = np.triu(Pg1 @ dg[:r].T - dx[:r] @ iPx1.T)                        # :468-469
    grad2 = Pg1 @ dg[r:].T - dx[:r] @ iPx2.T                                 # :470
    grad3 = Pg2 * dg[r:] - dx[r:] * iPx2                                     # :471
    max_abs_grad = np.maximum(np.maximum(_max_abs(grad1), _max_abs(grad2)), _max_abs(grad3))   # :472-474
    step0 = t(step) / (max_abs_grad + _tiny(L12))                            # :475
    newU1 = U1 - U1 @ (step0 * grad1)                                        # :476
    newU2 = U2 - U1 @ (step0 * grad2) - step0 * grad3.T * U2                 # :477
    newu3 = u3 - step0 * grad3 * u3                                          # :478
    return np.concatenate([newL1, newL2], axis=0), newl3, np.concatenate([newU1, newU2], axis=1), newu3   # :480


def precond_grad_splu(L12, l3, U12, u3, grads):
    """psgd.py:483-524."""
    grad = [np.reshape(g, [-1, 1]) for g in grads]                           # :495
    lens = [g.shape[0] for g in grad]                                        # :496
    grad = np.concatenate(grad, 0)                                           # :497
    r = U12.shape[0]
    L1, L2, U1, U2 = L12[:r], L12[r:], U12[:, :r], U12[:, r:]                # :499-503
    Ug1 = U1 @ grad[:r] + U2 @ grad[r:]                                      # :506
    Ug2 = u3 * grad[r:]                                                      # :507
    Qg1 = L1 @ Ug1                                                           # :509
    Qg2 = L2 @ Ug1 + l3 * Ug2                                                # :510
    LtQg1 = L1.T @ Qg1 + L2.T @ Qg2                                          # :512
    LtQg2 = l3 * Qg2                                                         # :513
    pre_grad = np.concatenate([U1.T @ LtQg1, U2.T @ LtQg1 + u3 * LtQg2], axis=0)   # :515-516
    pre_grads, idx = [], 0
    for i in range(len(grads)):                                              # :518-522
        pre_grads.append(np.reshape(pre_grad[idx: idx + lens[i]], np.shape(grads[i])))
        idx += lens[i]
    return pre_grads


# ----------------------------------------------------------------------------------------
# UVd: Q = (I + U V^T) diag(d)                                            psgd.py:540-627
# ----------------------------------------------------------------------------------------
def IpUVtmatvec(U, V, x):
    """psgd.py:540-544."""
    return x + U @ (V.T @ x)


def update_precond_UVd_math(U, V, d, v, h, step, tiny=None, *, balance=False, update_U=True):
    """psgd.py:554-617 with the two coin flips explicit; returns new (U, V, d).

    ``balance`` stands for ``tf.random.uniform([]) < 0.01`` (:562) and ``update_U`` for
    ``tf.random.uniform([]) < 0.5`` (:588)."""
    t = U.dtype.type
    tiny = _tiny(U) if tiny is None else t(tiny)
    step = t(step)
    if balance:                                                              # :562-567
        maxU = _max_abs(U)
        maxV = _max_abs(V)
        rho = np.sqrt(maxU / maxV)
        U = U / rho
        V = rho * V
    Qh = IpUVtmatvec(U, V, d * h)                                            # :569
    Ph = d * IpUVtmatvec(V, U, Qh)                                           # :570
    VtU = V.T @ U                                                            # :574
    IpVtU = np.eye(VtU.shape[0], dtype=VtU.dtype) + VtU                      # :575
    invQtv = v / d                                                           # :576
    invQtv = invQtv - V @ np.linalg.solve(IpVtU.T, U.T @ invQtv)             # :577 (adjoint=True)
    invPv = invQtv - U @ np.linalg.solve(IpVtU, V.T @ invQtv)                # :578
    invPv = invPv / d                                                        # :579
    nablaD = Ph * h - v * invPv                                              # :581
    mu = step / (_max_abs(nablaD) + tiny)                                    # :582
    d_new = d - mu * d * nablaD                                              # :584
    a, b = Qh, invQtv                                                        # :587
    if update_U:                                                             # :588
        atV = a.T @ V                                                        # :589
        atVVt = atV @ V.T                                                    # :590
        btV = b.T @ V                                                        # :591
        btVVt = btV @ V.T                                                    # :592
        norm = np.sqrt(np.abs((a.T @ a) * (atVVt @ atVVt.T)                  # :594-596
                              + (b.T @ b) * (btVVt @ btVVt.T)
                              - 2 * (a.T @ b) * (atVVt @ btVVt.T)))
        mu = step / (norm + tiny)                                            # :597
        U_new = U - mu * (a @ (atV @ IpVtU) - b @ (btV @ IpVtU))             # :600-601
        V_new = V
    else:
        atU = a.T @ U                                                        # :603
        btU = b.T @ U                                                        # :604
        UUta = U @ atU.T                                                     # :605
        UUtb = U @ btU.T                                                     # :606
        norm = np.sqrt(np.abs((UUta.T @ UUta) * (a.T @ a)                    # :608-610
                              + (UUtb.T @ UUtb) * (b.T @ b)
                              - 2 * (UUta.T @ UUtb) * (a.T @ b)))
        mu = step / (norm + tiny)                                            # :611
        V_new = V - mu * ((a + V @ atU.T) @ atU - (b + V @ btU.T) @ btU)     # :614-615
        U_new = U
    return U_new.astype(U.dtype, copy=False), V_new.astype(U.dtype, copy=False), d_new.astype(U.dtype, copy=False)


def precond_grad_UVd_math(U, V, d, g):
    """psgd.py:619-627."""
    g = IpUVtmatvec(U, V, d * g)                                             # :625
    g = d * IpUVtmatvec(V, U, g)                                             # :626
    return g


def uvd_step_tail(U, V, d, grad, params, lr_params, grad_clip_max_norm=np.inf, v=None, tiny=None):
    """Tail of ``UVd.step`` (psgd.py:747-762) on the flattened vectors (``grad``, ``params``, ``v``: [N, 1]).
    Returns ``(new_params, pre_grad)``.  ``v``: the finite-difference perturbation still sitting on the parameters
    (psgd.py:760-762) or None.  Restated arithmetic; the class itself needs tf.GradientTape, so this tail is not
    covered by the reference-generated golden vectors (its ``precond_grad_UVd_math`` call is)."""
    dt = U.dtype
    tiny = _tiny(U) if tiny is None else dt.type(tiny)
    pre_grad = precond_grad_UVd_math(U, V, d, grad)                          # :748
    if np.isinf(grad_clip_max_norm):                                         # :750-751
        lr = dt.type(lr_params)
    else:
        grad_norm = np.sqrt(np.sum(pre_grad * pre_grad, dtype=dt)) + tiny    # :753
        lr = dt.type(lr_params) * min(dt.type(grad_clip_max_norm) / grad_norm, dt.type(1.0))   # :754
    if v is None:
        new_params = params - lr * pre_grad                                  # :757-759
    else:
        new_params = params - (lr * pre_grad + v)                            # :760-762
    return new_params.astype(dt, copy=False), pre_grad


# ----------------------------------------------------------------------------------------
# X-shape and diagonal preconditioners -- NO reference code (README.md:11-15, :35);
# spec-derived, parity unpinned, validated by densification tests.
# ----------------------------------------------------------------------------------------
def update_precond_Xmat(a, b, v, h, step=0.01, tiny=None):
    """Q = diag(a) + adiag(b); returns new (a, b).  SURVEY.md appendix B."""
    t = a.dtype.type
    tiny = _tiny(a) if tiny is None else t(tiny)
    flip = lambda x: x[::-1]
    Qh = a * h + b * flip(h)
    aflip, bflip = flip(a), flip(b)
    invQtv = (aflip * v - bflip * flip(v)) / (a * aflip - b * bflip)
    nablaA = Qh * Qh - invQtv * invQtv
    nablaB = Qh * flip(Qh) - invQtv * flip(invQtv)
    n = a.shape[0]
    if n % 2 == 1:
        nablaB = nablaB.copy()
        nablaB[n // 2] = 0
    mu = t(step) / (np.maximum(_max_abs(nablaA), _max_abs(nablaB)) + tiny)
    a_new = a - mu * (nablaA * a + nablaB * bflip)
    b_new = b - mu * (nablaA * b + nablaB * aflip)
    return a_new, b_new


def precond_grad_Xmat(a, b, g):
    """Q^T Q g for Q = diag(a) + adiag(b).  SURVEY.md appendix B."""
    flip = lambda x: x[::-1]
    ab = a * b
    return (a * a + flip(b * b)) * g + (ab + flip(ab)) * flip(g)


def update_precond_diag(q, v, h, step=0.01, tiny=None):
    """Diagonal/Jacobi preconditioner Q = diag(q) (X-shape with b == 0)."""
    t = q.dtype.type
    tiny = _tiny(q) if tiny is None else t(tiny)
    Qh = q * h
    invQtv = v / q
    nabla = Qh * Qh - invQtv * invQtv
    mu = t(step) / (_max_abs(nabla) + tiny)
    return q - mu * nabla * q


def precond_grad_diag(q, g):
    return q * q * g
