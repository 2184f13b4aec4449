"""The multi-threaded torch-CPU timing twin (oracle/psgd_oracle_torch.py, used by bench.py's CPU legs) computes what the
NumPy oracle computes."""
import numpy as np
import pytest
import torch

from oracle import psgd_oracle as O
from oracle import psgd_oracle_torch as T
from tests import cases

t = lambda x: torch.from_numpy(np.ascontiguousarray(x))


@pytest.mark.parametrize("n,r", [(1021, 10), (20_000, 10), (300, 4)])
@pytest.mark.parametrize("kw", [dict(update_U=True, balance=False), dict(update_U=False, balance=False),
                                dict(update_U=True, balance=True)])
def test_uvd_twin_matches_numpy_oracle(n, r, kw):
    c = cases.uvd_case(900 + n, n, r, scale=3.0)
    Ur, Vr, dr = O.update_precond_UVd_math(c["U"], c["V"], c["d"], c["v"], c["h"], 0.01, **kw)
    U, V, d = T.update_precond_UVd_math(t(c["U"]), t(c["V"]), t(c["d"]), t(c["v"]), t(c["h"]), 0.01, **kw)
    for got, want in ((U, Ur), (V, Vr), (d, dr)):
        assert got.dtype == torch.float32 and cases.rel_err(got.numpy(), want) < 2e-6
    pre = T.precond_grad_UVd_math(U, V, d, t(c["g"]))
    assert cases.rel_err(pre.numpy(), O.precond_grad_UVd_math(Ur, Vr, dr, c["g"])) < 2e-6


@pytest.mark.parametrize("M,N", [(64, 64), (151, 16), (40, 130)])
def test_kron_dense_dense_twin_matches_numpy_oracle(M, N):
    k = cases.kron_case(7, "dense", "dense", M, N)
    qlr, qrr = O.update_precond_kron(k["Ql"], k["Qr"], k["dX"], k["dG"], 0.01)
    ql, qr = T.update_precond_dense_dense(t(k["Ql"]), t(k["Qr"]), t(k["dX"]), t(k["dG"]), 0.01)
    assert cases.rel_err(ql.numpy(), qlr) < 5e-6 and cases.rel_err(qr.numpy(), qrr) < 5e-6
    pg = T.precond_grad_dense_dense(ql, qr, t(k["G"]))
    assert cases.rel_err(pg.numpy(), O.precond_grad_kron(qlr, qrr, k["G"])) < 5e-6
