"""Seeded input builders shared by the golden-vector generator, the CPU tests and the GPU parity tests.
All NumPy ``default_rng(seed)``, float32."""
import numpy as np

F = np.float32

# cfg1: LeNet5 layer shapes (mnist_with_lenet5.py:12-16); cfg5: NMT factor shapes scaled down for CPU-second goldens
LENET_SHAPES = [(26, 6), (151, 16), (257, 120), (121, 84), (85, 10)]
KRON_KINDS = ["dense", "norm", "scale"]
KRON_COMBOS = [("dense", "dense"), ("dense", "norm"), ("dense", "scale"), ("norm", "dense"), ("norm", "scale"),
               ("scale", "dense"), ("scale", "norm")]


def triu_factor(rng, n, jitter=0.1):
    Q = np.triu(rng.standard_normal((n, n)) * jitter / np.sqrt(max(n, 1))) + np.diag(0.5 + rng.random(n))
    return Q.astype(F)


def norm_factor(rng, n):
    q = np.stack([0.5 + rng.random(n), 0.3 * rng.standard_normal(n) / np.sqrt(max(n, 1))]).astype(F)
    q[1, -1] = 0
    return q


def scale_factor(rng, n):
    return (0.5 + rng.random((1, n))).astype(F)


def factor(rng, kind, n):
    return {"dense": triu_factor, "norm": norm_factor, "scale": scale_factor}[kind](rng, n)


def kron_case(seed, kind_l, kind_r, M, N):
    rng = np.random.default_rng(seed)
    Ql, Qr = factor(rng, kind_l, M), factor(rng, kind_r, N)
    dX = rng.standard_normal((M, N)).astype(F)
    # a symmetric-positive-ish "Hessian" action keeps the factors well conditioned over trajectories
    dG = (dX * (0.5 + rng.random((M, 1))) * (0.5 + rng.random((1, N))) + 0.1 * rng.standard_normal((M, N))).astype(F)
    G = rng.standard_normal((M, N)).astype(F)
    return dict(Ql=Ql, Qr=Qr, dX=dX, dG=dG, G=G)


def uvd_case(seed, n, r, scale=1.0):
    rng = np.random.default_rng(seed)
    uv = scale * (1.0 / (n * r)) ** 0.5                                   # psgd.py:687
    U = (rng.standard_normal((n, r)) * uv).astype(F)
    V = (rng.standard_normal((n, r)) * uv).astype(F)
    d = (0.5 + rng.random((n, 1))).astype(F)
    v = rng.standard_normal((n, 1)).astype(F)
    h = ((0.5 + 1.5 * rng.random((n, 1))) * v + 0.1 * rng.standard_normal((n, 1))).astype(F)
    g = rng.standard_normal((n, 1)).astype(F)
    return dict(U=U, V=V, d=d, v=v, h=h, g=g)


def vec_case(seed, n):
    rng = np.random.default_rng(seed)
    a = (0.5 + rng.random(n)).astype(F)
    b = (0.2 * rng.standard_normal(n)).astype(F)
    v = rng.standard_normal(n).astype(F)
    h = ((0.5 + 1.5 * rng.random(n)) * v + 0.1 * rng.standard_normal(n)).astype(F)
    g = rng.standard_normal(n).astype(F)
    return dict(a=a, b=b, v=v, h=h, g=g)


def dense_case(seed, shapes):
    rng = np.random.default_rng(seed)
    n = int(sum(int(np.prod(s)) for s in shapes))
    Q = triu_factor(rng, n)
    dxs = [rng.standard_normal(s).astype(F) for s in shapes]
    dgs = [(x * (0.5 + rng.random(s)) + 0.1 * rng.standard_normal(s)).astype(F) for x, s in zip(dxs, shapes)]
    gs = [rng.standard_normal(s).astype(F) for s in shapes]
    return dict(Q=Q, dxs=dxs, dgs=dgs, gs=gs)


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
