"""The oracle against the REFERENCE'S OWN SOURCE.

``tests/golden/reference_outputs.npz`` was written by ``tests/golden/make_reference_golden.py``: the unmodified
``/root/reference/preconditioned_stochastic_gradient_descent.py`` executed on a NumPy stand-in for the TensorFlow ops
it calls.  The oracle must reproduce every array (bit for bit in the container that wrote them; a few ulps where the
BLAS build differs).  When /root/reference is present (dev container only -- never on the GPU box) the reference is
additionally run live, side by side with the oracle, on fresh random shapes and over multi-step trajectories.
"""
import os

import numpy as np
import pytest

from oracle import psgd_oracle as O
from tests import cases
from tests.golden import make_golden as MG
from tests.golden import make_reference_golden as MR

REF = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.npz"))
TOL = 2e-6   # BLAS summation order may differ between the writing container and the checking box


def _close(a, key):
    b = REF[key]
    assert a.shape == b.shape, (key, a.shape, b.shape)
    assert cases.rel_err(a, b) <= TOL, (key, cases.rel_err(a, b))


def test_kron_vs_reference_source():
    for seed, kl, kr, M, N in MG.KRON_GOLDEN:
        c = cases.kron_case(seed, kl, kr, M, N)
        ql, qr = O.update_precond_kron(c["Ql"], c["Qr"], c["dX"], c["dG"], 0.01)
        _close(ql, f"kron{seed}_Ql"); _close(qr, f"kron{seed}_Qr")
        _close(O.precond_grad_kron(c["Ql"], c["Qr"], c["G"]), f"kron{seed}_pre")


def test_uvd_vs_reference_source():
    for seed, n, r in MG.UVD_GOLDEN:
        c = cases.uvd_case(seed, n, r)
        for tag, kw in (("U", dict(update_U=True)), ("V", dict(update_U=False)), ("B", dict(update_U=True, balance=True))):
            U, V, d = O.update_precond_UVd_math(c["U"], c["V"], c["d"], c["v"], c["h"], 0.01, **kw)
            _close(U, f"uvd{seed}{tag}_U"); _close(V, f"uvd{seed}{tag}_V"); _close(d, f"uvd{seed}{tag}_d")
        _close(O.precond_grad_UVd_math(c["U"], c["V"], c["d"], c["g"]), f"uvd{seed}_pre")
        x = np.random.default_rng(seed).standard_normal((n, 3)).astype(np.float32)
        _close(O.IpUVtmatvec(c["U"], c["V"], x), f"uvd{seed}_matvec")


def test_dense_and_splu_vs_reference_source():
    for seed, shapes in MG.DENSE_GOLDEN:
        c = cases.dense_case(seed, shapes)
        _close(O.update_precond_dense(c["Q"], c["dxs"], c["dgs"], 0.01), f"dense{seed}_Q")
        for i, p in enumerate(O.precond_grad_dense(c["Q"], c["gs"])):
            _close(p, f"dense{seed}_pre{i}")
    for seed, shapes, r in MR.SPLU_GOLDEN:
        c = MR.splu_case(seed, shapes, r)
        out = O.update_precond_splu(c["L12"], c["l3"], c["U12"], c["u3"], c["dxs"], c["dgs"], 0.01)
        for nm, a in zip(("L12", "l3", "U12", "u3"), out):
            _close(a, f"splu{seed}_{nm}")
        for i, p in enumerate(O.precond_grad_splu(c["L12"], c["l3"], c["U12"], c["u3"], c["gs"])):
            _close(p, f"splu{seed}_pre{i}")


# ---- live: the reference source next to the oracle (dev container only) ------------------------------------------
needs_ref = pytest.mark.skipif(not os.path.exists(MR.REF_FILE), reason="/root/reference is absent (GPU box)")


@pytest.fixture(scope="module")
def ref():
    return MR.load_reference()


@needs_ref
def test_live_kron_random_shapes_and_trajectories(ref):
    mod, tf = ref
    rng = np.random.default_rng(11)
    for trial in range(40):
        kl, kr = cases.KRON_COMBOS[trial % len(cases.KRON_COMBOS)]
        M, N = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        # square factors of size 1 or 2 dispatch as dense (psgd.py:82-83), so structured sides need >= 3
        if kl != "dense":
            M = max(M, 3)
        if kr != "dense":
            N = max(N, 3)
        c = cases.kron_case(1000 + trial, kl, kr, M, N)
        Ql, Qr = c["Ql"], c["Qr"]
        Ql_r, Qr_r = Ql, Qr
        for step in range(5):
            x = cases.kron_case(5000 + 10 * trial + step, kl, kr, M, N)
            Ql, Qr = O.update_precond_kron(Ql, Qr, x["dX"], x["dG"], 0.02)
            Ql_r, Qr_r = (np.asarray(t) for t in mod.update_precond_kron(Ql_r, Qr_r, x["dX"], x["dG"], 0.02))
            assert np.array_equal(Ql, Ql_r) and np.array_equal(Qr, Qr_r), (kl, kr, M, N, step)
        assert np.array_equal(O.precond_grad_kron(Ql, Qr, c["G"]), np.asarray(mod.precond_grad_kron(Ql_r, Qr_r, c["G"])))


@needs_ref
def test_live_uvd_trajectory(ref):
    mod, tf = ref
    for seed, n, r in [(31, 257, 4), (32, 1021, 10)]:
        c = cases.uvd_case(seed, n, r)
        U, V, d = c["U"], c["V"], c["d"]
        Ur, Vr, dr = tf.Variable(U), tf.Variable(V), tf.Variable(d)
        flips = np.random.default_rng(seed).random((20, 2))
        flips[3, 0] = 0.001                                   # force one balancing step
        for k in range(20):
            x = cases.uvd_case(seed * 100 + k, n, r)
            tf.random.uniform_queue[:] = list(flips[k])
            mod.update_precond_UVd_math_(Ur, Vr, dr, x["v"], x["h"], np.float32(0.05), mod._tiny)
            U, V, d = O.update_precond_UVd_math(U, V, d, x["v"], x["h"], 0.05, balance=flips[k, 0] < 0.01,
                                                update_U=flips[k, 1] < 0.5)
            assert np.array_equal(U, np.asarray(Ur)) and np.array_equal(V, np.asarray(Vr)) and np.array_equal(d, np.asarray(dr)), k
        assert np.array_equal(O.precond_grad_UVd_math(U, V, d, c["g"]), np.asarray(mod.precond_grad_UVd_math(Ur, Vr, dr, c["g"])))


@needs_ref
def test_live_splu_and_dense_trajectory(ref):
    mod, tf = ref
    c = MR.splu_case(77, [(9, 4), (13,)], 5)
    s_o = (c["L12"], c["l3"], c["U12"], c["u3"]); s_r = s_o
    for k in range(8):
        x = MR.splu_case(700 + k, [(9, 4), (13,)], 5)
        s_o = O.update_precond_splu(*s_o, x["dxs"], x["dgs"], 0.05)
        s_r = tuple(np.asarray(t) for t in mod.update_precond_splu(*s_r, x["dxs"], x["dgs"], np.float32(0.05)))
        for a, b in zip(s_o, s_r):
            assert np.array_equal(a, b), k
    for a, b in zip(O.precond_grad_splu(*s_o, c["gs"]), mod.precond_grad_splu(*s_r, c["gs"])):
        assert np.array_equal(a, np.asarray(b))
    dcase = cases.dense_case(5, [(4, 3), (6,)])
    Q = Qr = dcase["Q"]
    for k in range(8):
        x = cases.dense_case(50 + k, [(4, 3), (6,)])
        Q = O.update_precond_dense(Q, x["dxs"], x["dgs"], 0.05)
        Qr = np.asarray(mod.update_precond_dense(Qr, x["dxs"], x["dgs"], np.float32(0.05)))
        assert np.array_equal(Q, Qr), k
