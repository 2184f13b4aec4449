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
    """Property check (hypothesis, derandomised): for random N, r, factor scales and d ranges the Gram-table expansion
    reproduces the reference's update + apply on both branches.

    Yardstick: at large factor scales (|U||V| ~ scale^2 >> 1) the preconditioned gradient is ill-conditioned and the
    float32 ORACLE itself sits 4e-6..6e-6 away from its float64 twin (n=824, r=1, scale=10: oracle 3.9e-6, expansion
    3.5e-6 from float64, 7.4e-6 from each other).  So every output is compared with the float64 twin and has to be
    as close to it as the float32 oracle is (factor 2), or within 2e-6 outright."""
    from hypothesis import given, settings, strategies as st

    f64 = lambda x: np.asarray(x, np.float64)

    @settings(max_examples=40, deadline=None, derandomize=True, database=None)
    @given(n=st.integers(2, 3000), r=st.integers(1, 16), scale=st.sampled_from([0.1, 1.0, 10.0, 50.0]),
           dlo=st.sampled_from([0.05, 0.5, 2.0]), update_U=st.booleans(), seed=st.integers(0, 10_000))
    def check(n, r, scale, dlo, update_U, seed):
        c = cases.uvd_case(seed, n, r, scale=scale)
        d = (dlo * (1.0 + np.random.default_rng(seed).random((n, 1)))).astype(np.float32)
        Ur, Vr, dr = O.update_precond_UVd_math(c["U"], c["V"], d, c["v"], c["h"], 0.01, balance=False, update_U=update_U)
        pr = O.precond_grad_UVd_math(Ur, Vr, dr, c["g"])
        U6, V6, d6 = O.update_precond_UVd_math(f64(c["U"]), f64(c["V"]), f64(d), f64(c["v"]), f64(c["h"]), 0.01,
                                               balance=False, update_U=update_U)
        p6 = O.precond_grad_UVd_math(U6, V6, d6, f64(c["g"]))
        U, V, dn, pre = M.update_apply(c["U"], c["V"], d, c["v"], c["h"], c["g"], 0.01, update_U)
        for got, o32, o64 in ((U, Ur, U6), (V, Vr, V6), (dn, dr, d6), (pre, pr, p6)):
            assert cases.rel_err(got, o64) <= max(2e-6, 2.0 * cases.rel_err(o32, o64))

    check()


def test_expansion_known_hard_case():
    """The example the random search once found: n=824, r=1, scale=10 -- pinned explicitly."""
    n, r, seed = 824, 1, 1
    f64 = lambda x: np.asarray(x, np.float64)
    for scale in (10.0, 50.0):
        c = cases.uvd_case(seed, n, r, scale=scale)
        d = (0.5 * (1.0 + np.random.default_rng(seed).random((n, 1)))).astype(np.float32)
        Ur, Vr, dr = O.update_precond_UVd_math(c["U"], c["V"], d, c["v"], c["h"], 0.01, balance=False, update_U=True)
        pr = O.precond_grad_UVd_math(Ur, Vr, dr, c["g"])
        U6, V6, d6 = O.update_precond_UVd_math(f64(c["U"]), f64(c["V"]), f64(d), f64(c["v"]), f64(c["h"]), 0.01,
                                               balance=False, update_U=True)
        p6 = O.precond_grad_UVd_math(U6, V6, d6, f64(c["g"]))
        _U, _V, _d, pre = M.update_apply(c["U"], c["V"], d, c["v"], c["h"], c["g"], 0.01, True)
        assert cases.rel_err(pre, p6) <= 2.0 * cases.rel_err(pr, p6)
        assert cases.rel_err(pre, pr) <= 1e-5          # and still inside the parity tolerance against the float32 oracle
