"""Generate tests/golden/*.npz from the oracle (oracle/psgd_oracle.py).

PARITY UNPINNED at the TensorFlow boundary (see the oracle's header): the reference has no golden vectors and
cannot run here, so these fixtures freeze the *restatement's* float32 outputs on seeded inputs.  They guard the
oracle against drift and give the GPU tests byte-stable expectations; the inputs are regenerated from the seeds
in tests/cases.py, only outputs are stored.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import psgd_oracle as O          # noqa: E402
from tests import cases                        # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

KRON_GOLDEN = [(100 + i, kl, kr, M, N) for i, (kl, kr, M, N) in enumerate(
    [(kl, kr, 12, 9) for kl, kr in cases.KRON_COMBOS] + [(kl, kr, 9, 12) for kl, kr in cases.KRON_COMBOS] +
    [("dense", "dense", M, N) for M, N in cases.LENET_SHAPES] +
    [("scale", "dense", 301, 256), ("norm", "scale", 65, 1024), ("scale", "dense", 2048, 10), ("dense", "dense", 1, 10)])]
UVD_GOLDEN = [(200, 1021, 10), (201, 37, 3), (202, 2048, 10), (203, 1500, 16), (204, 777, 1)]
VEC_GOLDEN = [(300, 1), (301, 2), (302, 1001), (303, 4096)]
DENSE_GOLDEN = [(400, [(2,)]), (401, [(3, 4), (5,), (2, 2, 2)])]


def main():
    out = {}
    for seed, kl, kr, M, N in KRON_GOLDEN:
        c = cases.kron_case(seed, kl, kr, M, N)
        ql, qr = O.update_precond_kron(c["Ql"], c["Qr"], c["dX"], c["dG"], 0.01)
        out[f"kron{seed}_Ql"], out[f"kron{seed}_Qr"] = ql, qr
        out[f"kron{seed}_pre"] = O.precond_grad_kron(c["Ql"], c["Qr"], c["G"])
    for seed, n, r in UVD_GOLDEN:
        c = cases.uvd_case(seed, n, r)
        for tag, kw in (("U", dict(update_U=True)), ("V", dict(update_U=False)), ("B", dict(update_U=True, balance=True))):
            U, V, d = O.update_precond_UVd_math(c["U"], c["V"], c["d"], c["v"], c["h"], 0.01, **kw)
            out[f"uvd{seed}{tag}_U"], out[f"uvd{seed}{tag}_V"], out[f"uvd{seed}{tag}_d"] = U, V, d
        out[f"uvd{seed}_pre"] = O.precond_grad_UVd_math(c["U"], c["V"], c["d"], c["g"])
    for seed, n in VEC_GOLDEN:
        c = cases.vec_case(seed, n)
        out[f"vec{seed}_a"], out[f"vec{seed}_b"] = O.update_precond_Xmat(c["a"], c["b"], c["v"], c["h"], 0.01)
        out[f"vec{seed}_xpre"] = O.precond_grad_Xmat(c["a"], c["b"], c["g"])
        out[f"vec{seed}_q"] = O.update_precond_diag(c["a"], c["v"], c["h"], 0.01)
        out[f"vec{seed}_dpre"] = O.precond_grad_diag(c["a"], c["g"])
    for seed, shapes in DENSE_GOLDEN:
        c = cases.dense_case(seed, shapes)
        out[f"dense{seed}_Q"] = O.update_precond_dense(c["Q"], c["dxs"], c["dgs"], 0.01)
        for i, p in enumerate(O.precond_grad_dense(c["Q"], c["gs"])):
            out[f"dense{seed}_pre{i}"] = p
    for k, v in out.items():
        assert np.asarray(v).dtype == np.float32, (k, np.asarray(v).dtype)
    np.savez_compressed(os.path.join(OUT, "oracle_outputs.npz"), **out)
    print(f"wrote {len(out)} arrays, {os.path.getsize(os.path.join(OUT, 'oracle_outputs.npz')) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
