"""Small hot-path invocations for compute-sanitizer (tools/sanitize_gpu.sh): every kernel family once, at sizes a
racecheck run finishes in a minute or two.  Results are still checked against the oracle (a sanitizer-clean kernel
with wrong results would be no use)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import psgd_tf_b200 as psgd
from oracle import psgd_oracle as O
from tests import cases

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
ctx = psgd.get_context()
errs = {}


def uvd(direct):
    ctx.set_option("direct", direct)
    c = cases.uvd_case(1, 6001, 10)
    for uu in (True, False):
        U, V, d = dev(c["U"]), dev(c["V"]), dev(c["d"])
        pre = psgd.update_precond_and_grad_UVd(U, V, d, dev(c["v"]), dev(c["h"]), dev(c["g"]), 0.01, psgd._tiny, balance=False, update_U=uu)
        Ur, Vr, dr = O.update_precond_UVd_math(c["U"], c["V"], c["d"], c["v"], c["h"], 0.01, balance=False, update_U=uu)
        errs[f"uvd_fused_d{direct}_u{int(uu)}"] = max(cases.rel_err(U.cpu().numpy(), Ur), cases.rel_err(V.cpu().numpy(), Vr),
                                                      cases.rel_err(pre.cpu().numpy(), O.precond_grad_UVd_math(Ur, Vr, dr, c["g"])))
    U, V, d = dev(c["U"]), dev(c["V"]), dev(c["d"])
    psgd.update_precond_UVd_math_(U, V, d, dev(c["v"]), dev(c["h"]), 0.01, psgd._tiny, balance=True, update_U=True)
    pre = psgd.precond_grad_UVd_math(U, V, d, dev(c["g"]))
    Ur, Vr, dr = O.update_precond_UVd_math(c["U"], c["V"], c["d"], c["v"], c["h"], 0.01, balance=True, update_U=True)
    errs[f"uvd_two_call_d{direct}"] = cases.rel_err(pre.cpu().numpy(), O.precond_grad_UVd_math(Ur, Vr, dr, c["g"]))
    ctx.set_option("direct", 0)


def kron_tc(mode):
    ctx.set_option("tc_mode", mode)
    ctx.set_option("gemm_path", 2)
    ctx.set_option("trsm_base", 128)
    c = cases.kron_case(5, "dense", "dense", 384, 512)
    ql, qr = psgd.update_precond_kron(dev(c["Ql"]), dev(c["Qr"]), dev(c["dX"]), dev(c["dG"]), 0.01)
    pre = psgd.precond_grad_kron(ql, qr, dev(c["G"]))
    qlr, qrr = O.update_precond_kron(c["Ql"], c["Qr"], c["dX"], c["dG"], 0.01)
    errs[f"kron_tc_mode{mode}"] = max(cases.rel_err(ql.cpu().numpy(), qlr), cases.rel_err(qr.cpu().numpy(), qrr),
                                      cases.rel_err(pre.cpu().numpy(), O.precond_grad_kron(qlr, qrr, c["G"])))
    ctx.set_option("tc_mode", 1); ctx.set_option("gemm_path", 0); ctx.set_option("trsm_base", 1024)


def kron_pair():
    ctx.set_option("tc_pair", 1)
    try:
        kron_tc(1)
        errs["kron_pair"] = errs.pop("kron_tc_mode1")
    finally:
        ctx.set_option("tc_pair", 0)


def kron_stream():
    for kl, kr, M, N in (("norm", "scale", 300, 257), ("norm", "dense", 200, 64), ("scale", "dense", 130, 48),
                         ("dense", "norm", 64, 200), ("scale", "norm", 257, 40)):
        c = cases.kron_case(9, kl, kr, M, N)
        ql, qr = psgd.update_precond_kron(dev(c["Ql"]), dev(c["Qr"]), dev(c["dX"]), dev(c["dG"]), 0.01)
        pre = psgd.precond_grad_kron(ql, qr, dev(c["G"]))
        qlr, qrr = O.update_precond_kron(c["Ql"], c["Qr"], c["dX"], c["dG"], 0.01)
        errs[f"kron_{kl}_{kr}"] = max(cases.rel_err(ql.cpu().numpy(), qlr), cases.rel_err(qr.cpu().numpy(), qrr),
                                      cases.rel_err(pre.cpu().numpy(), O.precond_grad_kron(qlr, qrr, c["G"])))


def splu():
    for n in (3001, 3000):          # odd n: 4-byte copies in the ring; even n and r: 8-byte pairs
        _splu(n, 10)


def _splu(n, r):
    rng = np.random.default_rng(4)
    L12 = np.concatenate([np.tril(0.1 * rng.standard_normal((r, r))) + np.eye(r), 0.1 * rng.standard_normal((n - r, r))]).astype(np.float32)
    U12 = np.concatenate([np.triu(0.1 * rng.standard_normal((r, r))) + np.eye(r), 0.1 * rng.standard_normal((r, n - r))], axis=1).astype(np.float32)
    l3 = (0.5 + rng.random((n - r, 1))).astype(np.float32); u3 = (0.5 + rng.random((n - r, 1))).astype(np.float32)
    dx = rng.standard_normal((n, 1)).astype(np.float32); dg = (1.5 * dx + 0.1 * rng.standard_normal((n, 1))).astype(np.float32)
    out = psgd.update_precond_splu(dev(L12), dev(l3), dev(U12), dev(u3), [dev(dx)], [dev(dg)], 0.01)
    want = O.update_precond_splu(L12, l3, U12, u3, [dx], [dg], 0.01)
    errs[f"splu_update_{n}"] = max(cases.rel_err(a.cpu().numpy(), b) for a, b in zip(out, want))
    pre = psgd.precond_grad_splu(dev(L12), dev(l3), dev(U12), dev(u3), [dev(dx)])
    errs[f"splu_apply_{n}"] = cases.rel_err(pre[0].cpu().numpy(), O.precond_grad_splu(L12, l3, U12, u3, [dx])[0])


def vec():
    x = cases.vec_case(3, 10_007)
    a, b = dev(x["a"]), dev(x["b"])
    psgd.update_precond_Xmat(a, b, dev(x["v"]), dev(x["h"]), 0.01)
    ar, br = O.update_precond_Xmat(x["a"], x["b"], x["v"], x["h"], 0.01)
    errs["xmat"] = max(cases.rel_err(a.cpu().numpy(), ar), cases.rel_err(b.cpu().numpy(), br))
    q = dev(x["a"])
    psgd.update_precond_diag(q, dev(x["v"]), dev(x["h"]), 0.01)
    errs["diag"] = cases.rel_err(q.cpu().numpy(), O.update_precond_diag(x["a"], x["v"], x["h"], 0.01))


def small():
    """SIMT engine of csrc/linalg.cu (panel solves, small GEMM) and the dense preconditioner's scans + vector solve."""
    for M, N in ((257, 120), (70, 545), (33, 1)):
        c = cases.kron_case(11, "dense", "dense", M, N)
        ql, qr = psgd.update_precond_kron(dev(c["Ql"]), dev(c["Qr"]), dev(c["dX"]), dev(c["dG"]), 0.01)
        pre = psgd.precond_grad_kron(ql, qr, dev(c["G"]))
        qlr, qrr = O.update_precond_kron(c["Ql"], c["Qr"], c["dX"], c["dG"], 0.01)
        errs[f"kron_small_{M}x{N}"] = max(cases.rel_err(ql.cpu().numpy(), qlr), cases.rel_err(qr.cpu().numpy(), qrr),
                                          cases.rel_err(pre.cpu().numpy(), O.precond_grad_kron(qlr, qrr, c["G"])))
    for n in (33, 700):
        c = cases.dense_case(12 + n, [(n,)])
        Qn = psgd.update_precond_dense(dev(c["Q"]), [dev(x) for x in c["dxs"]], [dev(x) for x in c["dgs"]], 0.01)
        errs[f"dense_update_{n}"] = cases.rel_err(Qn.cpu().numpy(), O.update_precond_dense(c["Q"], c["dxs"], c["dgs"], 0.01))


jobs = {"small": small, "uvd_tma": lambda: uvd(0), "uvd_direct": lambda: uvd(1), "kron_ts": lambda: kron_tc(1), "kron_ss": lambda: kron_tc(0), "kron_pair": kron_pair,
        "kron_stream": kron_stream, "splu": splu, "vec": vec}
for name, fn in jobs.items():
    if which in ("all", name):
        fn()
torch.cuda.synchronize()
print("sanitize_cases", which, {k: f"{v:.1e}" for k, v in errs.items()})
bad = {k: v for k, v in errs.items() if not v <= 1e-5}
assert not bad, bad
