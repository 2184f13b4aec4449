"""GPU parity tests: every public function of psgd_tf_b200 (through the C ABI) against the CPU oracle on the same
seeded inputs and against the committed golden outputs.  Tolerance: <= 1e-5 relative Frobenius error per output
(BASELINE.json north_star), written as TOL below.
"""
import os

import numpy as np
import pytest
import torch

from oracle import psgd_oracle as O
from tests import cases
from tests.golden import make_golden as MG

pytestmark = pytest.mark.gpu

TOL = 1e-5          # relative Frobenius error per step (north star)
_GDIR = os.path.join(os.path.dirname(__file__), "golden")
# Expectations: outputs of the REFERENCE'S OWN SOURCE (run on the NumPy TensorFlow stand-in, make_reference_golden.py)
# wherever the reference has code; the oracle's frozen outputs for the spec-derived diagonal / X-shape variants.
GOLD = dict(np.load(os.path.join(_GDIR, "oracle_outputs.npz")))
GOLD.update(np.load(os.path.join(_GDIR, "reference_outputs.npz")))


@pytest.fixture(scope="module")
def psgd():
    import psgd_tf_b200 as p
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device (B200); run with -m gpu on the GPU box")
    p.get_context()          # fails loudly if the extension is missing or the device is not sm_100
    return p


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def check(got, want, tol=TOL, what=""):
    e = cases.rel_err(host(got) if isinstance(got, torch.Tensor) else got, want)
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e}"
    return e


# ---------------------------------------------------------------------------------------------
# UVd
# ---------------------------------------------------------------------------------------------
UVD_SIZES = [(1021, 10), (37, 3), (256, 10), (255, 10), (257, 10), (4096, 10), (100_003, 10), (5000, 16), (777, 1),
             (3000, 7), (65_536, 4), (2048, 12), (1, 2)]


@pytest.mark.parametrize("n,r", UVD_SIZES)
@pytest.mark.parametrize("direct", [0, 1])
@pytest.mark.parametrize("fused", [1, 0])
def test_uvd_update_and_apply(psgd, n, r, direct, fused):
    """fused=1: two sweeps + d pass, rank-2 coefficients from the Gram table (default); fused=0: three sweeps."""
    ctx = psgd.get_context()
    ctx.set_option("direct", direct)
    ctx.set_option("uvd_fused", fused)
    try:
        c = cases.uvd_case(1000 + n + r, n, r)
        for kw in (dict(update_U=True, balance=False), dict(update_U=False, balance=False),
                   dict(update_U=True, balance=True)):
            U, V, d = dev(c["U"]), dev(c["V"]), dev(c["d"])
            assert psgd.update_precond_UVd_math_(U, V, d, dev(c["v"]), dev(c["h"]), 0.01, psgd._tiny, **kw) is None
            Ur, Vr, dr = O.update_precond_UVd_math(c["U"], c["V"], c["d"], c["v"], c["h"], 0.01, **kw)
            check(U, Ur, what=f"U {kw}"); check(V, Vr, what=f"V {kw}"); check(d, dr, what=f"d {kw}")
            if kw["update_U"] and not kw["balance"]:
                assert torch.equal(V, dev(c["V"])), "V must be untouched on the U branch"
            if not kw["update_U"]:
                assert torch.equal(U, dev(c["U"])), "U must be untouched on the V branch"
        pre = psgd.precond_grad_UVd_math(dev(c["U"]), dev(c["V"]), dev(c["d"]), dev(c["g"]))
        assert pre.shape == (n, 1)
        check(pre, O.precond_grad_UVd_math(c["U"], c["V"], c["d"], c["g"]), what="pre_grad")
        mv = psgd.IpUVtmatvec(dev(c["U"]), dev(c["V"]), dev(c["g"]))
        check(mv, O.IpUVtmatvec(c["U"], c["V"], c["g"]), what="IpUVtmatvec")
    finally:
        ctx.set_option("direct", 0)
        ctx.set_option("uvd_fused", 1)


@pytest.mark.parametrize("n,r", UVD_SIZES)
@pytest.mark.parametrize("direct", [0, 1])
def test_uvd_fused_update_apply(psgd, n, r, direct):
    """update_precond_and_grad_UVd (psgd_uvd_update_apply, three sweeps) == update then apply (psgd.py:732-748)."""
    ctx = psgd.get_context()
    ctx.set_option("direct", direct)
    try:
        c = cases.uvd_case(2000 + n + r, n, r, scale=3.0)
        for kw in (dict(update_U=True, balance=False), dict(update_U=False, balance=False),
                   dict(update_U=False, balance=True)):
            U, V, d = dev(c["U"]), dev(c["V"]), dev(c["d"])
            g = dev(c["g"])
            pre = psgd.update_precond_and_grad_UVd(U, V, d, dev(c["v"]), dev(c["h"]), g, 0.01, psgd._tiny, **kw)
            Ur, Vr, dr = O.update_precond_UVd_math(c["U"], c["V"], c["d"], c["v"], c["h"], 0.01, **kw)
            check(U, Ur, what=f"U {kw}"); check(V, Vr, what=f"V {kw}"); check(d, dr, what=f"d {kw}")
            assert pre.shape == (n, 1) and torch.equal(g, dev(c["g"]))
            check(pre, O.precond_grad_UVd_math(Ur, Vr, dr, c["g"]), what=f"pre_grad {kw}")
    finally:
        ctx.set_option("direct", 0)


def test_uvd_fused_update_apply_rejects_aliased_output(psgd):
    c = cases.uvd_case(13, 512, 4)
    ctx = psgd.get_context()
    U, V, d, v, h = (dev(c[k]) for k in ("U", "V", "d", "v", "h"))
    p = lambda t: t.data_ptr()
    rc = ctx.lib.psgd_uvd_update_apply(ctx.handle, p(U), p(V), p(d), p(v), p(h), p(d), p(d), 512, 4, 0.01, 1e-38, 0, 1)
    assert rc != 0 and b"alias" in ctx.lib.psgd_last_error()


@pytest.mark.parametrize("fused", [True, False])
def test_uvd_step_graph_replay_matches_eager(psgd, fused):
    """graphs.UVdStepGraphs replays exactly the kernels of update + apply: bit-identical state and output."""
    from psgd_tf_b200.graphs import UVdStepGraphs
    n, r = 50_021, 10
    c = cases.uvd_case(4242, n, r)
    ins = [tuple(dev(np.roll(c[k], s, 0)) for k in ("v", "h", "g")) for s in (0, 3)]
    Ue, Ve, de = dev(c["U"]), dev(c["V"]), dev(c["d"])
    Ug, Vg, dg = Ue.clone(), Ve.clone(), de.clone()
    gs = UVdStepGraphs(Ug, Vg, dg, 0.01, fused=fused)
    for i in range(9):
        v, h, g = ins[i % 2]
        flips = dict(balance=(i == 6), update_U=(i % 3 != 0))
        if fused:
            want = psgd.update_precond_and_grad_UVd(Ue, Ve, de, v, h, g, 0.01, psgd._tiny, **flips)
        else:
            psgd.update_precond_UVd_math_(Ue, Ve, de, v, h, 0.01, psgd._tiny, **flips)
            want = psgd.precond_grad_UVd_math(Ue, Ve, de, g)
        got = gs.step(v, h, g, **flips)
        torch.cuda.synchronize()
        assert torch.equal(got, want), i
        assert torch.equal(Ug, Ue) and torch.equal(Vg, Ve) and torch.equal(dg, de), i
    assert gs.replays >= 6


def test_uvd_golden(psgd):
    for seed, n, r in MG.UVD_GOLDEN:
        c = cases.uvd_case(seed, n, r)
        for tag, kw in (("U", dict(update_U=True, balance=False)), ("V", dict(update_U=False, balance=False)),
                        ("B", dict(update_U=True, balance=True))):
            U, V, d = dev(c["U"]), dev(c["V"]), dev(c["d"])
            psgd.update_precond_UVd(U, V, d, dev(c["v"]), dev(c["h"]), 0.01, psgd._tiny, **kw)
            check(U, GOLD[f"uvd{seed}{tag}_U"]); check(V, GOLD[f"uvd{seed}{tag}_V"]); check(d, GOLD[f"uvd{seed}{tag}_d"])
        check(psgd.precond_grad_UVd(dev(c["U"]), dev(c["V"]), dev(c["d"]), dev(c["g"])), GOLD[f"uvd{seed}_pre"])


def test_uvd_pipeline_and_direct_paths_agree_bitwise_on_maps(psgd):
    """The TMA-pipelined and direct-load kernels run the same per-row arithmetic; reductions differ only in
    summation order, so outputs agree to float32 round-off."""
    ctx = psgd.get_context()
    c = cases.uvd_case(77, 50_000, 10)
    outs = []
    for direct in (0, 1):
        ctx.set_option("direct", direct)
        U, V, d = dev(c["U"]), dev(c["V"]), dev(c["d"])
        psgd.update_precond_UVd_math_(U, V, d, dev(c["v"]), dev(c["h"]), 0.01, psgd._tiny, balance=False, update_U=True)
        outs.append((host(U), host(d)))
    ctx.set_option("direct", 0)
    assert cases.rel_err(outs[0][0], outs[1][0]) < 1e-6 and cases.rel_err(outs[0][1], outs[1][1]) < 1e-6


def test_uvd_is_deterministic(psgd):
    c = cases.uvd_case(78, 123_457, 10)
    res = []
    for _ in range(2):
        U, V, d = dev(c["U"]), dev(c["V"]), dev(c["d"])
        psgd.update_precond_UVd_math_(U, V, d, dev(c["v"]), dev(c["h"]), 0.01, psgd._tiny, balance=False, update_U=False)
        res.append((U.clone(), V.clone(), d.clone()))
    for a, b in zip(*res):
        assert torch.equal(a, b)


@pytest.mark.parametrize("form", ["separate", "separate-3sweep", "fused"])
def test_uvd_trajectory_100_steps(psgd, form):
    """Trajectory agreement over 100 steps with explicit inputs and coin flips (north star)."""
    psgd.get_context().set_option("uvd_fused", 0 if form == "separate-3sweep" else 1)
    n, r = 1021, 10                                        # cfg2: rnn_xor_UVd_preconditioner.py sizes
    c = cases.uvd_case(5, n, r)
    rng = np.random.default_rng(6)
    U, V, d = dev(c["U"]), dev(c["V"]), dev(c["d"])
    Ur, Vr, dr = c["U"].copy(), c["V"].copy(), c["d"].copy()
    hdiag = (0.5 + 1.5 * rng.random((n, 1))).astype(np.float32)
    worst = 0.0
    for t in range(100):
        v = rng.standard_normal((n, 1)).astype(np.float32)
        h = (hdiag * v).astype(np.float32)
        g = rng.standard_normal((n, 1)).astype(np.float32)
        kw = dict(balance=(t % 25 == 7), update_U=bool(rng.random() < 0.5))
        if form == "fused":
            pre = psgd.update_precond_and_grad_UVd(U, V, d, dev(v), dev(h), dev(g), 0.01, psgd._tiny, **kw)
        else:
            psgd.update_precond_UVd_math_(U, V, d, dev(v), dev(h), 0.01, psgd._tiny, **kw)
            pre = psgd.precond_grad_UVd_math(U, V, d, dev(g))
        Ur, Vr, dr = O.update_precond_UVd_math(Ur, Vr, dr, v, h, 0.01, **kw)
        worst = max(worst, cases.rel_err(host(pre), O.precond_grad_UVd_math(Ur, Vr, dr, g)))
    psgd.get_context().set_option("uvd_fused", 1)
    # per-step tolerance 1e-5; over a 100-step trajectory both float32 implementations drift independently
    assert worst < 1e-4, worst
    check(d, dr, 1e-4, "d after 100 steps"); check(U, Ur, 1e-4, "U after 100 steps"); check(V, Vr, 1e-4, "V after 100 steps")


def test_uvd_matvec_multi_column(psgd):
    c = cases.uvd_case(9, 3001, 5)
    x = np.random.default_rng(1).standard_normal((3001, 3)).astype(np.float32)
    check(psgd.IpUVtmatvec(dev(c["U"]), dev(c["V"]), dev(x)), O.IpUVtmatvec(c["U"], c["V"], x))


def test_uvd_rejects_bad_inputs(psgd):
    c = cases.uvd_case(10, 64, 4)
    with pytest.raises(RuntimeError, match="no CPU"):
        psgd.precond_grad_UVd_math(torch.from_numpy(c["U"]), dev(c["V"]), dev(c["d"]), dev(c["g"]))
    with pytest.raises(TypeError):
        psgd.precond_grad_UVd_math(dev(c["U"]).double(), dev(c["V"]), dev(c["d"]), dev(c["g"]))
    big = cases.uvd_case(11, 64, 17)
    with pytest.raises(ValueError, match="rank"):          # the Python mirror checks the documented limit up front
        psgd.precond_grad_UVd_math(dev(big["U"]), dev(big["V"]), dev(big["d"]), dev(big["g"]))
    with pytest.raises(ValueError, match="rank"):          # ... before class UVd re-homes any parameter
        w = torch.zeros(8, device="cuda", requires_grad=True)
        psgd.UVd([w], rank_of_modification=psgd.MAX_UVD_RANK + 1)
    assert w.data_ptr() != 0 and w.shape == (8,)
    ctx = psgd.get_context()                                # and the C ABI itself refuses, with an error message
    U, V, d, g = (dev(big[k]) for k in ("U", "V", "d", "g"))
    out = torch.empty_like(g)
    rc = ctx.lib.psgd_uvd_apply(ctx.handle, U.data_ptr(), V.data_ptr(), d.data_ptr(), g.data_ptr(), out.data_ptr(), 64, 17)
    assert rc != 0 and b"rank" in ctx.lib.psgd_last_error()


# ---------------------------------------------------------------------------------------------
# diagonal / X-shape
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 8, 16, 1001, 4096, 1_000_000, 1_000_003])
def test_xmat_and_diag(psgd, n):
    c = cases.vec_case(2000 + n, n)
    a, b = dev(c["a"]), dev(c["b"])
    assert psgd.update_precond_Xmat(a, b, dev(c["v"]), dev(c["h"]), 0.01) is None
    ar, br = O.update_precond_Xmat(c["a"], c["b"], c["v"], c["h"], 0.01)
    check(a, ar, what="a"); check(b, br, what="b")
    check(psgd.precond_grad_Xmat(dev(c["a"]), dev(c["b"]), dev(c["g"])), O.precond_grad_Xmat(c["a"], c["b"], c["g"]))
    q = dev(c["a"])
    psgd.update_precond_diag(q, dev(c["v"]), dev(c["h"]), 0.01)
    check(q, O.update_precond_diag(c["a"], c["v"], c["h"], 0.01), what="q")
    check(psgd.precond_grad_diag(dev(c["a"]), dev(c["g"])), O.precond_grad_diag(c["a"], c["g"]))


def test_vec_golden(psgd):
    for seed, n in MG.VEC_GOLDEN:
        c = cases.vec_case(seed, n)
        a, b = dev(c["a"]), dev(c["b"])
        psgd.update_precond_Xmat(a, b, dev(c["v"]), dev(c["h"]), 0.01)
        check(a, GOLD[f"vec{seed}_a"]); check(b, GOLD[f"vec{seed}_b"])
        q = dev(c["a"])
        psgd.update_precond_diag(q, dev(c["v"]), dev(c["h"]), 0.01)
        check(q, GOLD[f"vec{seed}_q"])
        check(psgd.precond_grad_Xmat(dev(c["a"]), dev(c["b"]), dev(c["g"])), GOLD[f"vec{seed}_xpre"])
        check(psgd.precond_grad_diag(dev(c["a"]), dev(c["g"])), GOLD[f"vec{seed}_dpre"])


# ---------------------------------------------------------------------------------------------
# Kronecker product preconditioners
# ---------------------------------------------------------------------------------------------
KRON_SHAPES = [(12, 9), (9, 12), (33, 65), (64, 64), (100, 37), (1, 10), (3, 1)]


@pytest.mark.parametrize("kl,kr", cases.KRON_COMBOS)
@pytest.mark.parametrize("M,N", KRON_SHAPES)
def test_kron_all_format_combinations(psgd, kl, kr, M, N):
    if (kl == "norm" and M == 2) or (kr == "norm" and N == 2) or (kl == "scale" and M == 1) or (kr == "scale" and N == 1):
        pytest.skip("square factor shapes are dense by the reference's dispatch order")
    if (kl != "dense" and M == 1 and kl == "norm") or (kr == "norm" and N == 1):
        pytest.skip("degenerate")
    c = cases.kron_case(3000 + 7 * M + N, kl, kr, M, N)
    ql, qr = psgd.update_precond_kron(dev(c["Ql"]), dev(c["Qr"]), dev(c["dX"]), dev(c["dG"]), 0.01)
    qlr, qrr = O.update_precond_kron(c["Ql"], c["Qr"], c["dX"], c["dG"], 0.01)
    assert tuple(ql.shape) == qlr.shape and tuple(qr.shape) == qrr.shape
    check(ql, qlr, what="Ql"); check(qr, qrr, what="Qr")
    pre = psgd.precond_grad_kron(dev(c["Ql"]), dev(c["Qr"]), dev(c["G"]))
    check(pre, O.precond_grad_kron(c["Ql"], c["Qr"], c["G"]), what="pre_grad")


@pytest.mark.parametrize("M,N", [(257, 120), (120, 257), (513, 33), (33, 513), (1101, 70), (70, 1101), (545, 130), (1025, 1)])
def test_kron_dense_pairs_on_the_panel_solves(psgd, M, N):
    """Dense pairs below the tensor-core thresholds (odd sizes, thin gradients): triangular solves by the 1024-thread
    panel kernels of csrc/linalg.cu -- one panel (n <= 512), several panels joined by SIMT GEMMs (n > 512), partial last
    blocks, slabs with fewer than 32 right-hand sides (psgd.py:171-172)."""
    c = cases.kron_case(9100 + 3 * M + N, "dense", "dense", M, N)
    ql, qr = psgd.update_precond_kron(dev(c["Ql"]), dev(c["Qr"]), dev(c["dX"]), dev(c["dG"]), 0.01)
    qlr, qrr = O.update_precond_kron(c["Ql"], c["Qr"], c["dX"], c["dG"], 0.01)
    check(ql, qlr, what="Ql"); check(qr, qrr, what="Qr")
    pre = psgd.precond_grad_kron(dev(c["Ql"]), dev(c["Qr"]), dev(c["G"]))
    check(pre, O.precond_grad_kron(c["Ql"], c["Qr"], c["G"]), what="pre_grad")


@pytest.mark.parametrize("n", [1, 2, 33, 129, 600, 1500])
@pytest.mark.parametrize("scan", [1, 0])
@pytest.mark.parametrize("full", [False, True])
def test_dense_update_scan_form_and_vector_solve_over_panels(psgd, n, scan, full):
    """Dense update (psgd.py:26-42) beyond one 512-row panel of the vector solve; scan = 1: Q - mu triu(a a^T - b b^T) Q as
    column suffix scans (O(n^2)), scan = 0: the reference's n^3 product.  full: Q with a non-zero lower triangle -- the
    solve reads only the upper one (tf.linalg.triangular_solve), the products all of it (tf.matmul)."""
    c = cases.dense_case(4100 + n, [(n,)])
    Q = c["Q"].copy()
    if full:
        Q += np.tril(0.05 * np.random.default_rng(n).standard_normal((n, n)).astype(np.float32), -1)
    ctx = psgd.get_context()
    ctx.set_option("dense_scan", scan)
    try:
        Qn = psgd.update_precond_dense(dev(Q), [dev(x) for x in c["dxs"]], [dev(x) for x in c["dgs"]], 0.01)
    finally:
        ctx.set_option("dense_scan", 1)
    check(Qn, O.update_precond_dense(Q, c["dxs"], c["dgs"], 0.01), what="Q")


def test_dense_update_scan_trajectory(psgd):
    """50 dense updates in a row, scan form, against the oracle's trajectory."""
    n = 300
    c = cases.dense_case(77, [(n,)])
    rng = np.random.default_rng(78)
    Qd, Qo = dev(c["Q"]), c["Q"]
    for _ in range(50):
        dx = rng.standard_normal(n).astype(np.float32)
        dg = (dx * (0.5 + rng.random(n)) + 0.1 * rng.standard_normal(n)).astype(np.float32)
        Qd = psgd.update_precond_dense(Qd, [dev(dx)], [dev(dg)], 0.01)
        Qo = O.update_precond_dense(Qo, [dx], [dg], 0.01)
    check(Qd, Qo, tol=5e-5, what="Q after 50 steps")


@pytest.mark.parametrize("M,N", [(2305, 1024), (1025, 4935), (300, 7), (5, 3000), (64, 256), (65, 257), (3, 1)])
@pytest.mark.parametrize("mirror", [False, True])
def test_kron_norm_scale_fused_streaming_kernels(psgd, M, N, mirror):
    """(normalization, scaling) and its mirror (scaling, normalization) at NMT-like and ragged shapes: the fused
    one-pass kernels of csrc/kron_stream.cu against the oracle (psgd.py:328-391)."""
    kl, kr = ("scale", "norm") if mirror else ("norm", "scale")
    if mirror:
        M, N = N, M
    if (kl == "norm" and M == 2) or (kr == "norm" and N == 2) or (kl == "scale" and M == 1) or (kr == "scale" and N == 1):
        pytest.skip("square factors are dense by the reference's dispatch order")
    c = cases.kron_case(5000 + M + N, kl, kr, M, N)
    ql, qr = psgd.update_precond_kron(dev(c["Ql"]), dev(c["Qr"]), dev(c["dX"]), dev(c["dG"]), 0.01)
    qlr, qrr = O.update_precond_kron(c["Ql"], c["Qr"], c["dX"], c["dG"], 0.01)
    check(ql, qlr, what="Ql"); check(qr, qrr, what="Qr")
    pre = psgd.precond_grad_kron(dev(c["Ql"]), dev(c["Qr"]), dev(c["G"]))
    check(pre, O.precond_grad_kron(c["Ql"], c["Qr"], c["G"]), what="pre_grad")


def test_dense_apply_large_gemv_pair(psgd):
    """precond_grad_dense at n = 3000 (not a multiple of the tile sizes): two coalesced GEMVs over Q."""
    rng = np.random.default_rng(11)
    n = 3000
    Q = (np.triu(rng.standard_normal((n, n))) * 0.02 + np.eye(n)).astype(np.float32)
    gs = [rng.standard_normal((40, 50)).astype(np.float32), rng.standard_normal((1000,)).astype(np.float32)]
    got = psgd.precond_grad_dense(dev(Q), [dev(g) for g in gs])
    want = O.precond_grad_dense(Q, gs)
    for a, b in zip(got, want):
        assert tuple(a.shape) == b.shape
        check(a, b, what="dense apply")


def test_kron_golden(psgd):
    for seed, kl, kr, M, N in MG.KRON_GOLDEN:
        c = cases.kron_case(seed, kl, kr, M, N)
        ql, qr = psgd.update_precond_kron(dev(c["Ql"]), dev(c["Qr"]), dev(c["dX"]), dev(c["dG"]), 0.01)
        check(ql, GOLD[f"kron{seed}_Ql"], what=f"{kl},{kr} {M}x{N} Ql")
        check(qr, GOLD[f"kron{seed}_Qr"], what=f"{kl},{kr} {M}x{N} Qr")
        check(psgd.precond_grad_kron(dev(c["Ql"]), dev(c["Qr"]), dev(c["G"])), GOLD[f"kron{seed}_pre"],
              what=f"{kl},{kr} {M}x{N} pre")


def test_kron_inputs_untouched_and_identity_fixed_point(psgd):
    c = cases.kron_case(1, "dense", "dense", 20, 30)
    Ql, Qr, dX, dG = dev(c["Ql"]), dev(c["Qr"]), dev(c["dX"]), dev(c["dG"])
    keep = [t.clone() for t in (Ql, Qr, dX, dG)]
    psgd.update_precond_kron(Ql, Qr, dX, dG, torch.tensor(0.01))
    for a, b in zip((Ql, Qr, dX, dG), keep):
        assert torch.equal(a, b), "functional API must not modify its inputs"
    G = dev(c["G"])
    for L in (torch.eye(20), torch.stack([torch.ones(20), torch.zeros(20)]), torch.ones(1, 20)):
        for R in (torch.eye(30), torch.stack([torch.ones(30), torch.zeros(30)]), torch.ones(1, 30)):
            if L.shape[0] == R.shape[0] and L.shape[0] in (1, 2):
                continue
            check(psgd.precond_grad_kron(L.cuda(), R.cuda(), G), c["G"], 1e-6)      # README.md:48


def test_kron_unknown_combination_prints_and_passes_through(psgd, capsys):
    rng = np.random.default_rng(0)
    a, b = dev(cases.norm_factor(rng, 5)), dev(cases.norm_factor(rng, 4))
    G = dev(rng.standard_normal((5, 4)).astype(np.float32))
    x, y = psgd.update_precond_kron(a, b, G, G, 0.01)
    assert x is a and y is b
    assert psgd.precond_grad_kron(a, b, G) is G
    assert "Unknown Kronecker product preconditioner" in capsys.readouterr().out


def test_kron_triangular_solve_ignores_lower_triangle_of_Q(psgd):
    """tf.linalg.triangular_solve(lower=False) reads only the upper triangle; tf.matmul reads everything
     -- the oracle makes the same distinction."""
    c = cases.kron_case(2, "dense", "dense", 24, 17)
    rng = np.random.default_rng(3)
    Ql = c["Ql"] + np.tril(rng.standard_normal((24, 24)).astype(np.float32) * 0.01, -1)
    ql, qr = psgd.update_precond_kron(dev(Ql), dev(c["Qr"]), dev(c["dX"]), dev(c["dG"]), 0.01)
    qlr, qrr = O.update_precond_kron(Ql, c["Qr"], c["dX"], c["dG"], 0.01)
    check(ql, qlr); check(qr, qrr)


def test_kron_lenet_batched_matches_per_layer(psgd):
    cs = [cases.kron_case(50 + i, "dense", "dense", M, N) for i, (M, N) in enumerate(cases.LENET_SHAPES)]
    outs = psgd.update_precond_kron_batched([dev(c["Ql"]) for c in cs], [dev(c["Qr"]) for c in cs],
                                            [dev(c["dX"]) for c in cs], [dev(c["dG"]) for c in cs], 0.01)
    pres = psgd.precond_grad_kron_batched([dev(c["Ql"]) for c in cs], [dev(c["Qr"]) for c in cs], [dev(c["G"]) for c in cs])
    for c, (ql, qr), pre in zip(cs, outs, pres):
        qlr, qrr = O.update_precond_kron(c["Ql"], c["Qr"], c["dX"], c["dG"], 0.01)
        check(ql, qlr); check(qr, qrr)
        check(pre, O.precond_grad_kron(c["Ql"], c["Qr"], c["G"]))


def test_kron_step_graph_replay_matches_eager(psgd):
    """graphs.KronStepGraphs (factors ping-pong between two state sets, one CUDA graph per direction) replays exactly the
    kernels of update_precond_kron_batched + precond_grad_kron_batched: bit-identical factors and results over a
    mixed-format layer list (NMT-like: every canonical and mirrored combination)."""
    from psgd_tf_b200.graphs import KronStepGraphs
    layers = [("dense", "dense", 26, 6), ("scale", "dense", 300, 32), ("norm", "scale", 129, 260), ("dense", "norm", 24, 70),
              ("dense", "dense", 1, 10), ("scale", "norm", 33, 40)]
    cs = [cases.kron_case(900 + i, kl, kr, M, N) for i, (kl, kr, M, N) in enumerate(layers)]
    ins = [([dev(np.roll(c["dX"], s, 0)) for c in cs], [dev(np.roll(c["dG"], s, 0)) for c in cs],
            [dev(np.roll(c["G"], s, 0)) for c in cs]) for s in (0, 1)]
    Ql, Qr = [dev(c["Ql"]) for c in cs], [dev(c["Qr"]) for c in cs]
    gr = KronStepGraphs(Ql, Qr, 0.01)
    for t in range(7):
        dX, dG, G = ins[t & 1]
        new = psgd.update_precond_kron_batched(Ql, Qr, dX, dG, 0.01)
        Ql, Qr = [a for a, _ in new], [b for _, b in new]
        want = psgd.precond_grad_kron_batched(Ql, Qr, G)
        got = gr.step(dX, dG, G)
        for a, b in zip(got, want):
            assert torch.equal(a, b)
        for (a, b), x, y in zip(gr.factors, Ql, Qr):
            assert torch.equal(a, x) and torch.equal(b, y)
    assert gr.replays >= 4
    # one step of the graph-replayed state against the oracle as well
    c0 = cs[0]
    q = O.update_precond_kron(c0["Ql"], c0["Qr"], c0["dX"], c0["dG"], 0.01)
    gr2 = KronStepGraphs([dev(c0["Ql"])], [dev(c0["Qr"])], 0.01)
    pre = gr2.step([dev(c0["dX"])], [dev(c0["dG"])], [dev(c0["G"])])
    check(pre[0], O.precond_grad_kron(q[0], q[1], c0["G"]))


def test_kron_batched_rejects_inconsistent_layers(psgd):
    """The batched forms run the same shape / device checks per layer as the single-layer entry points (a mismatched
    layer would make the kernels index out of bounds)."""
    c = cases.kron_case(3, "dense", "dense", 26, 6)
    good = [dev(c[k]) for k in ("Ql", "Qr", "dX", "dG", "G")]
    with pytest.raises(ValueError, match="layer 1"):
        psgd.update_precond_kron_batched([good[0], good[0]], [good[1], good[1]], [good[2], dev(c["dX"][:20])],
                                         [good[3], good[3]], 0.01)
    with pytest.raises(ValueError, match="layer 0"):
        psgd.precond_grad_kron_batched([dev(np.eye(25, dtype=np.float32))], [good[1]], [good[4]])
    with pytest.raises(ValueError, match="length"):
        psgd.precond_grad_kron_batched([good[0]], [good[1], good[1]], [good[4]])
    with pytest.raises(ValueError):
        psgd.precond_grad_UVd_math(dev(np.zeros((8, 2), np.float32)), dev(np.zeros((8, 3), np.float32)),
                                   dev(np.ones((8, 1), np.float32)), dev(np.ones((8, 1), np.float32)))
    with pytest.raises(ValueError):
        psgd.precond_grad_dense(dev(np.zeros((6, 4), np.float32)), [dev(np.zeros(6, np.float32))])


def test_kron_trajectory_100_steps(psgd):
    """100-step trajectory on the first LeNet5 layer (mnist_with_lenet5.py:12) with explicit inputs.

    The max-abs normalised update is a chaotic map once the factors have converged (a float32 and a float64 run of the
    *oracle itself* drift apart, and 1-ulp input noise is amplified ~1e3x over 100 steps), so agreement is judged
    against the float64 twin: the CUDA path must track it as closely as the float32 oracle does."""
    M, N = 26, 6
    rng = np.random.default_rng(11)
    Ql, Qr = dev(np.eye(M, dtype=np.float32)), dev(np.eye(N, dtype=np.float32))
    Q32 = [np.eye(M, dtype=np.float32), np.eye(N, dtype=np.float32)]
    Q64 = [np.eye(M), np.eye(N)]
    S = (0.5 + rng.random((M, 1))).astype(np.float32); T = (0.5 + rng.random((1, N))).astype(np.float32)
    worst_gpu, worst_o32, first50 = 0.0, 0.0, 0.0
    for t in range(100):
        dX = rng.standard_normal((M, N)).astype(np.float32)
        dG = (S * dX * T + 0.1 * rng.standard_normal((M, N))).astype(np.float32)
        G = rng.standard_normal((M, N)).astype(np.float32)
        Ql, Qr = psgd.update_precond_kron(Ql, Qr, dev(dX), dev(dG), 0.01)
        Q32 = O.update_precond_kron(Q32[0], Q32[1], dX, dG, 0.01)
        Q64 = O.update_precond_kron(Q64[0], Q64[1], dX.astype(np.float64), dG.astype(np.float64), 0.01)
        ref = O.precond_grad_kron(Q64[0], Q64[1], G.astype(np.float64))
        e_gpu = cases.rel_err(host(psgd.precond_grad_kron(Ql, Qr, dev(G))), ref)
        e_o32 = cases.rel_err(O.precond_grad_kron(Q32[0], Q32[1], G), ref)
        worst_gpu, worst_o32 = max(worst_gpu, e_gpu), max(worst_o32, e_o32)
        if t < 50:
            first50 = max(first50, e_gpu)
    assert first50 < 1e-5, first50                              # before sensitivity sets in: the per-step bar
    assert worst_gpu <= max(1e-5, 5 * worst_o32), (worst_gpu, worst_o32)


# ---------------------------------------------------------------------------------------------
# dense full-matrix preconditioner
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shapes", [[(2,)], [(3, 4), (5,), (2, 2, 2)], [(40, 10), (30,)], [(1,)]])
def test_dense_update_and_apply(psgd, shapes):
    c = cases.dense_case(4000 + len(shapes) + shapes[0][0], shapes)
    Qn = psgd.update_precond_dense(dev(c["Q"]), [dev(x) for x in c["dxs"]], [dev(x) for x in c["dgs"]], 0.01)
    check(Qn, O.update_precond_dense(c["Q"], c["dxs"], c["dgs"], 0.01), what="Q")
    pres = psgd.precond_grad_dense(dev(c["Q"]), [dev(x) for x in c["gs"]])
    refs = O.precond_grad_dense(c["Q"], c["gs"])
    assert len(pres) == len(refs)
    for p, r, g in zip(pres, refs, c["gs"]):
        assert tuple(p.shape) == g.shape
        check(p, r, what="pre_grad")


def test_dense_rosenbrock_trajectory(psgd):
    """hello_psgd.py:10-27 on the GPU path with closed-form derivatives."""
    rng = np.random.default_rng(21)
    x = np.array([-1.0, 1.0], np.float32)
    Q = dev((0.1 * np.eye(2)).astype(np.float32))
    f = lambda x: 100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2
    grad = lambda x: np.array([-400 * x[0] * (x[1] - x[0] ** 2) - 2 * (1 - x[0]), 200 * (x[1] - x[0] ** 2)], np.float32)
    hess = lambda x: np.array([[1200 * x[0] ** 2 - 400 * x[1] + 2, -400 * x[0]], [-400 * x[0], 200]], np.float32)
    f0 = f(x)
    for _ in range(500):
        dx = rng.standard_normal(2).astype(np.float32)
        Q = psgd.update_precond_dense(Q, [dev(dx)], [dev(hess(x) @ dx)], 0.2)
        x = x - 0.5 * host(psgd.precond_grad_dense(Q, [dev(grad(x))])[0])
    assert f(x) < 1e-3 * f0


def test_launch_counter_counts_kernels(psgd):
    ctx = psgd.get_context()
    c = cases.uvd_case(12, 4096, 10)
    before = ctx.launch_count
    psgd.precond_grad_UVd_math(dev(c["U"]), dev(c["V"]), dev(c["d"]), dev(c["g"]))
    assert ctx.launch_count - before == 3       # Gram sweep, mid kernel (partial reduce + r x r solve), map sweep
