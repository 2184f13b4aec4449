"""Generate tests/golden/reference_outputs.npz by RUNNING THE REFERENCE'S OWN SOURCE FILE.

    python tests/golden/make_reference_golden.py            # needs /root/reference (dev container only)

``/root/reference/preconditioned_stochastic_gradient_descent.py`` is imported unmodified, with
``tests/golden/tf_numpy_shim`` first on ``sys.path`` so that its ``import tensorflow as tf`` binds to the NumPy
stand-in (TensorFlow itself cannot be installed here).  Every golden array is therefore produced by the reference's
own control flow / dispatch / operation order; only the leaf array ops are NumPy + LAPACK.  The same seeded inputs as
``make_golden.py`` are used (tests/cases.py), so ``tests/test_reference_golden.py`` can hold the oracle, and the GPU
tests can hold the CUDA path, against the reference directly.

Two reference behaviours need a scripted environment:
  * the two coin flips inside update_precond_UVd_math_ (psgd.py:562, :588) come from ``tf.random.uniform([])``;
    the shim serves them from a queue that this script fills (0.005 / 0.5 => balance yes/no; 0.25 / 0.75 => U / V);
  * TF's CPU kernels flush denormals, which makes the reference's ``_tiny`` halving loop (psgd.py:22) stop at
    2**-126; NumPy would continue to 2**-149, so ``_tiny`` is overwritten with the TF value after import.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_FILE = "/root/reference/preconditioned_stochastic_gradient_descent.py"
sys.path.insert(0, ROOT)


def load_reference(ref_file=REF_FILE):
    """Import the unmodified reference module on top of the NumPy TensorFlow stand-in."""
    shim = os.path.join(HERE, "tf_numpy_shim")
    saved = sys.modules.pop("tensorflow", None)
    sys.path.insert(0, shim)
    try:
        import tensorflow as tf
        assert tf.__version__.endswith("numpy-shim"), "a real TensorFlow shadowed the shim"
        spec = importlib.util.spec_from_file_location("psgd_reference_on_shim", ref_file)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(shim)
        sys.modules.pop("tensorflow", None)
        if saved is not None:
            sys.modules["tensorflow"] = saved
    mod._tiny = np.float32(2.0 ** -126)        # TF flushes denormals (see module docstring)
    return mod, tf


SPLU_GOLDEN = [(500, [(7,), (3, 4)], 3), (501, [(40, 5), (11,)], 10), (502, [(6,)], 1)]


def splu_case(seed, shapes, r):
    """Inputs for update_precond_splu / precond_grad_splu (psgd.py:396-524; demo_usage_of_all_preconditioners.py:43-52
    initialises L12 = [I; 0], l3 = 1, U12 = [I, 0], u3 = 1 scaled; here perturbed so every block is exercised)."""
    rng = np.random.default_rng(seed)
    F = np.float32
    n = int(sum(int(np.prod(s)) for s in shapes))
    L12 = np.concatenate([np.tril(0.1 * rng.standard_normal((r, r))) + np.diag(0.5 + rng.random(r)),
                          0.1 * rng.standard_normal((n - r, r))], 0).astype(F)
    U12 = np.concatenate([np.triu(0.1 * rng.standard_normal((r, r))) + np.diag(0.5 + rng.random(r)),
                          0.1 * rng.standard_normal((r, n - r))], 1).astype(F)
    l3 = (0.5 + rng.random((n - r, 1))).astype(F)
    u3 = (0.5 + rng.random((n - r, 1))).astype(F)
    dxs = [rng.standard_normal(s).astype(F) for s in shapes]
    dgs = [(x * (0.5 + rng.random(s)) + 0.1 * rng.standard_normal(s)).astype(F) for x, s in zip(dxs, shapes)]
    gs = [rng.standard_normal(s).astype(F) for s in shapes]
    return dict(L12=L12, l3=l3, U12=U12, u3=u3, dxs=dxs, dgs=dgs, gs=gs)


def main():
    from tests import cases
    from tests.golden import make_golden as MG
    ref, tf = load_reference()
    A = np.asarray
    out = {}
    for seed, kl, kr, M, N in MG.KRON_GOLDEN:
        c = cases.kron_case(seed, kl, kr, M, N)
        ql, qr = ref.update_precond_kron(c["Ql"], c["Qr"], c["dX"], c["dG"], 0.01)
        out[f"kron{seed}_Ql"], out[f"kron{seed}_Qr"] = A(ql), A(qr)
        out[f"kron{seed}_pre"] = A(ref.precond_grad_kron(c["Ql"], c["Qr"], c["G"]))
    for seed, n, r in MG.UVD_GOLDEN:
        c = cases.uvd_case(seed, n, r)
        for tag, flips in (("U", [0.5, 0.25]), ("V", [0.5, 0.75]), ("B", [0.005, 0.25])):
            U, V, d = tf.Variable(c["U"]), tf.Variable(c["V"]), tf.Variable(c["d"])
            tf.random.uniform_queue[:] = flips
            ret = ref.update_precond_UVd_math_(U, V, d, c["v"], c["h"], np.float32(0.01), ref._tiny)
            assert ret is None and not tf.random.uniform_queue
            out[f"uvd{seed}{tag}_U"], out[f"uvd{seed}{tag}_V"], out[f"uvd{seed}{tag}_d"] = A(U), A(V), A(d)
        out[f"uvd{seed}_pre"] = A(ref.precond_grad_UVd_math(c["U"], c["V"], c["d"], c["g"]))
        x = np.random.default_rng(seed).standard_normal((n, 3)).astype(np.float32)
        out[f"uvd{seed}_matvec"] = A(ref.IpUVtmatvec(c["U"], c["V"], x))
    for seed, shapes in MG.DENSE_GOLDEN:
        c = cases.dense_case(seed, shapes)
        out[f"dense{seed}_Q"] = A(ref.update_precond_dense(c["Q"], c["dxs"], c["dgs"], np.float32(0.01)))
        for i, p in enumerate(ref.precond_grad_dense(c["Q"], c["gs"])):
            out[f"dense{seed}_pre{i}"] = A(p)
    for seed, shapes, r in SPLU_GOLDEN:
        c = splu_case(seed, shapes, r)
        L12, l3, U12, u3 = ref.update_precond_splu(c["L12"], c["l3"], c["U12"], c["u3"], c["dxs"], c["dgs"], np.float32(0.01))
        out[f"splu{seed}_L12"], out[f"splu{seed}_l3"], out[f"splu{seed}_U12"], out[f"splu{seed}_u3"] = A(L12), A(l3), A(U12), A(u3)
        for i, p in enumerate(ref.precond_grad_splu(c["L12"], c["l3"], c["U12"], c["u3"], c["gs"])):
            out[f"splu{seed}_pre{i}"] = A(p)
    # the unsupported-combination convention (psgd.py:89-91): inputs come back untouched
    c = cases.kron_case(7, "norm", "norm", 5, 6)
    ql, qr = ref.update_precond_kron(c["Ql"], c["Qr"], c["dX"], c["dG"], 0.01)
    assert np.array_equal(A(ql), c["Ql"]) and np.array_equal(A(qr), c["Qr"])
    for k, v in out.items():
        assert v.dtype == np.float32, (k, v.dtype)
        out[k] = np.ascontiguousarray(v)
    path = os.path.join(HERE, "reference_outputs.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {len(out)} arrays from the reference source, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
