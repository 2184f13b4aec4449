"""GPU tests of the tcgen05 3xTF32 engine: GEMM against float64, and the dense-factor Kron paths at sizes that route
through the tensor cores, against the CPU oracle.  Tolerance 1e-5 relative Frobenius error (north star)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import psgd_oracle as O
from tests import cases

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def psgd():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device (B200); run with -m gpu on the GPU box")
    import psgd_tf_b200 as p
    p.get_context()
    return p


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def gemm(psgd, engine, A, B, ta, tb, triu=0, a_tri=0, b_tri=0):
    from psgd_tf_b200._lib import check
    ctx = psgd.get_context()
    M = A.shape[1] if ta else A.shape[0]
    K = A.shape[0] if ta else A.shape[1]
    N = B.shape[0] if tb else B.shape[1]
    out = torch.full((M, N), float("nan"), device="cuda")
    check(ctx.lib.psgd_gemm(ctx.handle, engine, M, N, K, C.c_void_p(A.data_ptr()), A.shape[1], ta,
                            C.c_void_p(B.data_ptr()), B.shape[1], tb, C.c_void_p(out.data_ptr()), N, triu, a_tri, b_tri))
    return out


@pytest.mark.parametrize("ta,tb", [(0, 1), (1, 1), (0, 0), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(256, 256, 256), (384, 640, 320), (1000, 520, 264), (128, 128, 32), (132, 36, 40),
                                   (2048, 1024, 4096), (256, 256, 9416), (384, 128, 5000)])     # the last two: split-K
def test_gemm_tc_matches_float64(psgd, M, N, K, ta, tb):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn((K, M) if ta else (M, K), device="cuda", generator=g)
    B = torch.randn((N, K) if tb else (K, N), device="cuda", generator=g)
    ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
    for engine in (1, 2):
        out = gemm(psgd, engine, A, B, ta, tb)
        err = ((out.double() - ref).norm() / ref.norm()).item()
        assert err < 2e-6, (engine, err)


def test_gemm_tc_triangular_hints_and_mask(psgd):
    g = torch.Generator(device="cuda").manual_seed(5)
    n = 640
    A = torch.triu(torch.randn(n, n, device="cuda", generator=g))
    B = torch.triu(torch.randn(n, n, device="cuda", generator=g))
    ref = torch.triu(A.double() @ B.double())
    out = gemm(psgd, 2, A, B, 0, 0, triu=1, a_tri=1, b_tri=1)
    assert ((out.double() - ref).norm() / ref.norm()).item() < 2e-6
    # lower-triangular views: A^T (lower) times B (upper), no mask
    ref2 = A.double().t() @ B.double()
    out2 = gemm(psgd, 2, A, B, 1, 0, a_tri=2, b_tri=1)
    assert ((out2.double() - ref2).norm() / ref2.norm()).item() < 2e-6
    # X * B^T with B^T lower
    X = torch.randn(300, n, device="cuda", generator=g)
    ref3 = X.double() @ B.double().t()
    out3 = gemm(psgd, 2, X, B, 0, 1, b_tri=2)
    assert ((out3.double() - ref3).norm() / ref3.norm()).item() < 2e-6


@pytest.mark.parametrize("pair", [0, 1])
def test_gemm_tc_pair_kernel_and_splitk_options(psgd, pair):
    """The optional cta_group::2 kernel (tc_pair) and the split-K path against float64, with the options toggled."""
    ctx = psgd.get_context()
    g = torch.Generator(device="cuda").manual_seed(99 + pair)
    ctx.set_option("tc_pair", pair)
    try:
        for (M, N, K, ta, tb) in ((512, 384, 1024, 0, 1), (300, 260, 520, 1, 0), (256, 256, 9416, 0, 1)):
            A = torch.randn((K, M) if ta else (M, K), device="cuda", generator=g)
            B = torch.randn((N, K) if tb else (K, N), device="cuda", generator=g)
            ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
            for splitk in (1, 0):
                ctx.set_option("tc_splitk", splitk)
                out = gemm(psgd, 2, A, B, ta, tb)
                assert ((out.double() - ref).norm() / ref.norm()).item() < 2e-6, (M, N, K, pair, splitk)
    finally:
        ctx.set_option("tc_pair", 0)
        ctx.set_option("tc_splitk", 1)


def test_gemm_tc_rejects_unaligned_leading_dimension(psgd):
    A = torch.randn(256, 258, device="cuda")
    B = torch.randn(258, 256, device="cuda")
    with pytest.raises(psgd.PsgdError):
        gemm(psgd, 2, A, B, 0, 0)
    out = gemm(psgd, 0, A, B, 0, 0)          # auto falls back to the SIMT engine
    ref = A.double() @ B.double()
    assert ((out.double() - ref).norm() / ref.norm()).item() < 2e-6


KRON_TC = [("dense", "dense", 512, 512), ("dense", "dense", 384, 640), ("dense", "dense", 640, 384),
           ("norm", "dense", 700, 512), ("dense", "scale", 512, 900), ("scale", "dense", 1000, 256),
           ("dense", "norm", 512, 600), ("dense", "dense", 1024, 1024),
           # NMT embedding shapes (mirrored, long side not a multiple of 4: padded transposed copies + split-K Gram products)
           ("scale", "dense", 1001, 256), ("scale", "dense", 4935, 256), ("scale", "dense", 9414, 256)]


@pytest.mark.parametrize("kl,kr,M,N", KRON_TC)
@pytest.mark.parametrize("path", [2, 1])
def test_kron_large_layers_through_tensor_cores(psgd, kl, kr, M, N, path):
    ctx = psgd.get_context()
    ctx.set_option("gemm_path", path)
    try:
        c = cases.kron_case(7000 + M + N, kl, kr, M, N)
        before = ctx.launch_count
        ql, qr = psgd.update_precond_kron(dev(c["Ql"]), dev(c["Qr"]), dev(c["dX"]), dev(c["dG"]), 0.01)
        pre = psgd.precond_grad_kron(dev(c["Ql"]), dev(c["Qr"]), dev(c["G"]))
        assert ctx.launch_count > before
        qlr, qrr = O.update_precond_kron(c["Ql"], c["Qr"], c["dX"], c["dG"], 0.01)
        e1, e2 = cases.rel_err(ql.cpu().numpy(), qlr), cases.rel_err(qr.cpu().numpy(), qrr)
        e3 = cases.rel_err(pre.cpu().numpy(), O.precond_grad_kron(c["Ql"], c["Qr"], c["G"]))
        assert max(e1, e2, e3) <= TOL, (e1, e2, e3)
    finally:
        ctx.set_option("gemm_path", 0)


def test_kron_tensor_core_trajectory_20_steps(psgd):
    """State carried across steps through the tcgen05 path stays on the float64 twin's trajectory."""
    n = 512
    rng = np.random.default_rng(3)
    Ql, Qr = dev(np.eye(n, dtype=np.float32)), dev(np.eye(n, dtype=np.float32))
    Q64 = [np.eye(n), np.eye(n)]
    S = (0.5 + 1.5 * rng.random((n, 1))); T = (0.5 + 1.5 * rng.random((1, n)))
    worst = 0.0
    for _ in range(20):
        dX = rng.standard_normal((n, n)).astype(np.float32)
        dG = (S * dX * T + 0.1 * rng.standard_normal((n, n))).astype(np.float32)
        G = rng.standard_normal((n, n)).astype(np.float32)
        Ql, Qr = psgd.update_precond_kron(Ql, Qr, dev(dX), dev(dG), 0.01)
        Q64 = O.update_precond_kron(Q64[0], Q64[1], dX.astype(np.float64), dG.astype(np.float64), 0.01)
        worst = max(worst, cases.rel_err(psgd.precond_grad_kron(Ql, Qr, dev(G)).cpu().numpy(),
                                         O.precond_grad_kron(Q64[0], Q64[1], G.astype(np.float64))))
    assert worst < TOL, worst


def test_dense_preconditioner_large(psgd):
    c = cases.dense_case(9, [(20, 30), (424,)])          # n = 1024 -> the n^3 product runs on the tensor cores
    Qn = psgd.update_precond_dense(dev(c["Q"]), [dev(x) for x in c["dxs"]], [dev(x) for x in c["dgs"]], 0.01)
    assert cases.rel_err(Qn.cpu().numpy(), O.update_precond_dense(c["Q"], c["dxs"], c["dgs"], 0.01)) <= TOL
