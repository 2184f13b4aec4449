"""Multi-threaded CPU timing twin of the oracle -- TEST INFRASTRUCTURE ONLY (same import rules as psgd_oracle.py).

``psgd_oracle.py`` is the numerical checker: NumPy, whose element-wise ops run on ONE thread.  The reference executes on
TensorFlow's CPU runtime, which spreads element-wise ops, reductions and GEMMs over every host core (Eigen thread
pool).  For the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- "the reference's CPU path with all the
host threads it can use" -- this file restates the two headline paths op for op, in the reference's order and
association, on torch CPU tensors (ATen: OpenMP element-wise kernels, MKL/oneDNN sgemm, LAPACK trsm / gesv -- the same
kernel classes TF-CPU dispatches to).  ``tests/test_oracle_torch.py`` checks it against the NumPy oracle.

    update_precond_UVd_math / precond_grad_UVd_math     psgd.py:554-627
    update_precond_dense_dense / precond_grad_dense_dense   psgd.py:156-192 (the 24 x 4096^2 Kron stack)
"""
from __future__ import annotations

import torch

TINY = 2.0 ** -126          # psgd.py:20-22 (smallest normal float32; TF flushes denormals)


def IpUVtmatvec(U, V, x):
    """psgd.py:540-544."""
    return x + U @ (V.t() @ x)


def update_precond_UVd_math(U, V, d, v, h, step, tiny=TINY, *, balance=False, update_U=True):
    """psgd.py:554-617 with the two coin flips explicit; returns new (U, V, d)."""
    if balance:                                                              # :562-567
        rho = torch.sqrt(U.abs().max() / V.abs().max())
        U = U / rho
        V = rho * V
    Qh = IpUVtmatvec(U, V, d * h)                                            # :569
    Ph = d * IpUVtmatvec(V, U, Qh)                                           # :570
    VtU = V.t() @ U                                                          # :574
    IpVtU = torch.eye(VtU.shape[0], dtype=VtU.dtype) + VtU                   # :575
    invQtv = v / d                                                           # :576
    invQtv = invQtv - V @ torch.linalg.solve(IpVtU.t(), U.t() @ invQtv)      # :577 (adjoint=True)
    invPv = invQtv - U @ torch.linalg.solve(IpVtU, V.t() @ invQtv)           # :578
    invPv = invPv / d                                                        # :579
    nablaD = Ph * h - v * invPv                                              # :581
    mu = step / (nablaD.abs().max() + tiny)                                  # :582
    d_new = d - mu * d * nablaD                                              # :584
    a, b = Qh, invQtv                                                        # :587
    if update_U:                                                             # :588
        atV = a.t() @ V                                                      # :589
        atVVt = atV @ V.t()                                                  # :590
        btV = b.t() @ V                                                      # :591
        btVVt = btV @ V.t()                                                  # :592
        norm = torch.sqrt(torch.abs((a.t() @ a) * (atVVt @ atVVt.t())        # :594-596
                                    + (b.t() @ b) * (btVVt @ btVVt.t())
                                    - 2 * (a.t() @ b) * (atVVt @ btVVt.t())))
        mu = step / (norm + tiny)                                            # :597
        return U - mu * (a @ (atV @ IpVtU) - b @ (btV @ IpVtU)), V, d_new    # :600-601
    atU = a.t() @ U                                                          # :603
    btU = b.t() @ U                                                          # :604
    UUta = U @ atU.t()                                                       # :605
    UUtb = U @ btU.t()                                                       # :606
    norm = torch.sqrt(torch.abs((UUta.t() @ UUta) * (a.t() @ a)              # :608-610
                                + (UUtb.t() @ UUtb) * (b.t() @ b)
                                - 2 * (UUta.t() @ UUtb) * (a.t() @ b)))
    mu = step / (norm + tiny)                                                # :611
    return U, V - mu * ((a + V @ atU.t()) @ atU - (b + V @ btU.t()) @ btU), d_new     # :614-615


def precond_grad_UVd_math(U, V, d, g):
    """psgd.py:619-627."""
    g = IpUVtmatvec(U, V, d * g)                                             # :625
    return d * IpUVtmatvec(V, U, g)                                          # :626


def _triu_solve_adjoint(Q, B):
    """tf.linalg.triangular_solve(Q, B, lower=False, adjoint=True): Q^T X = B from the upper triangle of Q."""
    return torch.linalg.solve_triangular(Q.triu().t(), B, upper=False)


def update_precond_dense_dense(Ql, Qr, dX, dG, step=0.01, tiny=TINY):
    """psgd.py:156-179."""
    rho = torch.sqrt(Ql.diagonal().max() / Qr.diagonal().max())              # :166-168
    Ql = Ql / rho                                                            # :169
    Qr = rho * Qr                                                            # :170
    A = Ql @ (dG @ Qr.t())                                                   # :173
    Bt = _triu_solve_adjoint(Ql, _triu_solve_adjoint(Qr, dX.t()).t())        # :174
    grad1 = torch.triu(A @ A.t() - Bt @ Bt.t())                              # :175
    grad2 = torch.triu(A.t() @ A - Bt.t() @ Bt)                              # :176
    step1 = step / (grad1.abs().max() + tiny)                                # :177
    step2 = step / (grad2.abs().max() + tiny)                                # :178
    return Ql - (step1 * grad1) @ Ql, Qr - (step2 * grad2) @ Qr              # :179


def precond_grad_dense_dense(Ql, Qr, Grad):
    """psgd.py:182-192 (association switches on M < N)."""
    if Grad.shape[0] < Grad.shape[1]:
        return (((Ql.t() @ Ql) @ Grad) @ Qr.t()) @ Qr                        # :190
    return Ql.t() @ (Ql @ (Grad @ (Qr.t() @ Qr)))                            # :192
