"""The oracle reproduces the committed golden outputs bit for bit (guards the restatement against drift)."""
import os
import sys

import numpy as np

from oracle import psgd_oracle as O
from tests import cases
from tests.golden import make_golden as MG

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_outputs.npz"))

# BLAS builds may differ in summation order between the container that wrote the fixtures and the box that checks
# them; allow a few ulps but nothing that would matter at the 1e-5 parity bar.
TOL = 2e-6


def _close(a, b):
    assert cases.rel_err(a, b) <= TOL, cases.rel_err(a, b)


def test_kron_golden():
    for seed, kl, kr, M, N in MG.KRON_GOLDEN:
        c = cases.kron_case(seed, kl, kr, M, N)
        ql, qr = O.update_precond_kron(c["Ql"], c["Qr"], c["dX"], c["dG"], 0.01)
        _close(ql, GOLD[f"kron{seed}_Ql"]); _close(qr, GOLD[f"kron{seed}_Qr"])
        _close(O.precond_grad_kron(c["Ql"], c["Qr"], c["G"]), GOLD[f"kron{seed}_pre"])


def test_uvd_golden():
    for seed, n, r in MG.UVD_GOLDEN:
        c = cases.uvd_case(seed, n, r)
        U, V, d = O.update_precond_UVd_math(c["U"], c["V"], c["d"], c["v"], c["h"], 0.01, update_U=True)
        _close(U, GOLD[f"uvd{seed}U_U"]); _close(d, GOLD[f"uvd{seed}U_d"])
        U, V, d = O.update_precond_UVd_math(c["U"], c["V"], c["d"], c["v"], c["h"], 0.01, update_U=False)
        _close(V, GOLD[f"uvd{seed}V_V"])
        _close(O.precond_grad_UVd_math(c["U"], c["V"], c["d"], c["g"]), GOLD[f"uvd{seed}_pre"])


def test_vec_and_dense_golden():
    for seed, n in MG.VEC_GOLDEN:
        c = cases.vec_case(seed, n)
        a, b = O.update_precond_Xmat(c["a"], c["b"], c["v"], c["h"], 0.01)
        _close(a, GOLD[f"vec{seed}_a"]); _close(b, GOLD[f"vec{seed}_b"])
        _close(O.update_precond_diag(c["a"], c["v"], c["h"], 0.01), GOLD[f"vec{seed}_q"])
    for seed, shapes in MG.DENSE_GOLDEN:
        c = cases.dense_case(seed, shapes)
        _close(O.update_precond_dense(c["Q"], c["dxs"], c["dgs"], 0.01), GOLD[f"dense{seed}_Q"])
