"""CPU check of the algebra behind the fused UVd sweeps (tests/uvd_pipeline_model.py) against the oracle: the
Gram-table re-association must reproduce update_precond_UVd_math_ / precond_grad_UVd_math (psgd.py:554-627) to float32
round-off.  The CUDA kernels implement exactly this model; the GPU tests compare them with the oracle directly."""
import numpy as np
import pytest

from oracle import psgd_oracle as O
from tests import cases
from tests import uvd_pipeline_model as M


@pytest.mark.parametrize("n,r", [(1021, 10), (37, 3), (5000, 16), (777, 1), (100_003, 10)])
@pytest.mark.parametrize("update_U", [True, False])
@pytest.mark.parametrize("scale", [1.0, 30.0])
def test_two_sweep_update_matches_oracle(n, r, update_U, scale):
    c = cases.uvd_case(300 + n + r, n, r, scale=scale)
    Ur, Vr, dr = O.update_precond_UVd_math(c["U"], c["V"], c["d"], c["v"], c["h"], 0.01, balance=False, update_U=update_U)
    U, V, d = M.update(c["U"], c["V"], c["d"], c["v"], c["h"], 0.01, update_U)
    assert cases.rel_err(U, Ur) < 2e-6 and cases.rel_err(V, Vr) < 2e-6 and cases.rel_err(d, dr) < 2e-6
    if update_U:
        assert np.array_equal(V, c["V"])
    else:
        assert np.array_equal(U, c["U"])


@pytest.mark.parametrize("n,r", [(1021, 10), (37, 3), (5000, 16), (100_003, 10)])
@pytest.mark.parametrize("update_U", [True, False])
def test_fused_update_apply_matches_oracle(n, r, update_U):
    c = cases.uvd_case(400 + n + r, n, r, scale=10.0)
    Ur, Vr, dr = O.update_precond_UVd_math(c["U"], c["V"], c["d"], c["v"], c["h"], 0.01, balance=False, update_U=update_U)
    pr = O.precond_grad_UVd_math(Ur, Vr, dr, c["g"])
    U, V, d, pre = M.update_apply(c["U"], c["V"], c["d"], c["v"], c["h"], c["g"], 0.01, update_U)
    assert cases.rel_err(U, Ur) < 2e-6 and cases.rel_err(V, Vr) < 2e-6 and cases.rel_err(d, dr) < 2e-6
    assert cases.rel_err(pre, pr) < 2e-6


def test_two_sweep_update_tracks_oracle_over_a_trajectory():
    n, r = 1021, 10
    c = cases.uvd_case(5, n, r)
    rng = np.random.default_rng(6)
    U, V, d = c["U"].copy(), c["V"].copy(), c["d"].copy()
    Ur, Vr, dr = U.copy(), V.copy(), d.copy()
    hdiag = (0.5 + 1.5 * rng.random((n, 1))).astype(np.float32)
    for t in range(100):
        v = rng.standard_normal((n, 1)).astype(np.float32)
        h = (hdiag * v).astype(np.float32)
        uu = bool(rng.random() < 0.5)
        U, V, d = M.update(U, V, d, v, h, 0.01, uu)
        Ur, Vr, dr = O.update_precond_UVd_math(Ur, Vr, dr, v, h, 0.01, balance=False, update_U=uu)
    assert cases.rel_err(U, Ur) < 1e-4 and cases.rel_err(V, Vr) < 1e-4 and cases.rel_err(d, dr) < 1e-4


def test_expansion_holds_over_random_shapes_and_scales():
    """Property check (hypothesis): for random N, r, factor scales and d ranges the Gram-table expansion reproduces the
    oracle's update + apply to float32 round-off, on both branches."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(n=st.integers(2, 3000), r=st.integers(1, 16), scale=st.sampled_from([0.1, 1.0, 10.0, 50.0]),
           dlo=st.sampled_from([0.05, 0.5, 2.0]), update_U=st.booleans(), seed=st.integers(0, 10_000))
    def check(n, r, scale, dlo, update_U, seed):
        c = cases.uvd_case(seed, n, r, scale=scale)
        d = (dlo * (1.0 + np.random.default_rng(seed).random((n, 1)))).astype(np.float32)
        Ur, Vr, dr = O.update_precond_UVd_math(c["U"], c["V"], d, c["v"], c["h"], 0.01, balance=False, update_U=update_U)
        pr = O.precond_grad_UVd_math(Ur, Vr, dr, c["g"])
        U, V, dn, pre = M.update_apply(c["U"], c["V"], d, c["v"], c["h"], c["g"], 0.01, update_U)
        for got, want in ((U, Ur), (V, Vr), (dn, dr), (pre, pr)):
            assert cases.rel_err(got, want) < 5e-6

    check()
