"""Caller-side tail of a Kron training step (mnist_with_lenet5.py:54-56, neural_machine_translation_with_attention.py:200,
:206) over ragged layer lists: global grad-norm clipping + in-place parameter update, finite-difference differences."""
import numpy as np
import pytest
import torch

from tests import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def psgd():
    import psgd_tf_b200 as p
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device (B200); run with -m gpu on the GPU box")
    p.get_context()
    return p


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


SHAPES = cases.LENET_SHAPES + [(1,), (70_001,), (3, 5, 7)]


@pytest.mark.parametrize("clip", [None, 0.05, 1e9])
@pytest.mark.parametrize("with_v", [False, True])
def test_apply_preconditioned_updates(psgd, clip, with_v):
    rng = np.random.default_rng(5)
    Ws = [rng.standard_normal(s).astype(np.float32) for s in SHAPES]
    gs = [rng.standard_normal(s).astype(np.float32) for s in SHAPES]
    vs = [(rng.standard_normal(s) * 2.0 ** -11.5).astype(np.float32) for s in SHAPES] if with_v else None
    lr = np.float32(0.1)
    # mnist_with_lenet5.py:54-56
    if clip is None:
        adj = np.float32(1.0)
    else:
        norm = np.sqrt(np.float32(sum(np.sum(g * g, dtype=np.float32) for g in gs)))
        adj = np.minimum(np.float32(clip) / norm, np.float32(1.0))
    want = [w - (adj * lr * g + (v if vs is not None else 0)) for w, g, v in zip(Ws, gs, vs or [0] * len(Ws))]
    Wd = [dev(w) for w in Ws]
    assert psgd.apply_preconditioned_updates(Wd, [dev(g) for g in gs], float(lr), clip, [dev(v) for v in vs] if vs else None) is None
    for got, w in zip(Wd, want):
        assert got.shape == w.shape
        assert cases.rel_err(got.cpu().numpy(), w.astype(np.float32)) <= 1e-6
    # 70 layers: more than one launch chunk (64 layers per launch)
    many = [rng.standard_normal((17,)).astype(np.float32) for _ in range(70)]
    Wm = [dev(np.zeros(17, np.float32)) for _ in many]
    psgd.apply_preconditioned_updates(Wm, [dev(g) for g in many], 1.0, 1.0)
    norm = np.sqrt(sum(float(np.sum(g.astype(np.float64) ** 2)) for g in many))
    for got, g in zip(Wm, many):
        assert np.allclose(got.cpu().numpy(), -min(1.0 / norm, 1.0) * g, rtol=1e-5, atol=1e-7)


def test_grad_differences(psgd):
    rng = np.random.default_rng(6)
    a = [rng.standard_normal(s).astype(np.float32) for s in SHAPES]
    b = [rng.standard_normal(s).astype(np.float32) for s in SHAPES]
    out = psgd.grad_differences([dev(x) for x in a], [dev(x) for x in b])
    for o, x, y in zip(out, a, b):
        assert o.shape == x.shape and np.array_equal(o.cpu().numpy(), x - y)
    assert psgd.grad_differences([], []) == []
