"""Convergence regressions: the reference's own "tests" are its demo scripts converging.  These are
the two self-contained ones -- hello_psgd.py (Rosenbrock, dense preconditioner) and
demo_usage_of_all_preconditioners.py (sparse tensor decomposition; dense + the three Kronecker format examples) -- run
through psgd_tf_b200 on the GPU with torch.autograd in the role of tf.GradientTape.  Thresholds were calibrated by running
the same loops through the CPU oracle (Rosenbrock reaches 0.0 within 500 iterations; the decomposition loss falls from
~4e4-7e4 to ~800-960 in 100 iterations for every variant)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def psgd():
    import psgd_tf_b200 as p
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device (B200); run with -m gpu on the GPU box")
    p.get_context()
    return p


def test_hello_psgd_rosenbrock(psgd):
    """hello_psgd.py:7-27: Q = 0.1 I, preconditioner step 0.2, learning rate 0.5, 500 iterations."""
    torch.manual_seed(0)
    xs = [torch.tensor([-1.0], device="cuda", requires_grad=True), torch.tensor([1.0], device="cuda", requires_grad=True)]
    Q = 0.1 * torch.eye(2, device="cuda")

    def f(xs):
        x1, x2 = xs
        return (100.0 * (x2 - x1 ** 2) ** 2 + (1.0 - x1) ** 2).sum()

    values = []
    for _ in range(500):
        y = f(xs)
        grads = torch.autograd.grad(y, xs, create_graph=True)
        vs = [torch.randn_like(x) for x in xs]
        hess_vs = torch.autograd.grad(sum((g * v).sum() for g, v in zip(grads, vs)), xs)
        values.append(y.item())
        Q = psgd.update_precond_dense(Q, vs, hess_vs, step=0.2)                    # hello_psgd.py:25
        pre = psgd.precond_grad_dense(Q, [g.detach() for g in grads])              # :26
        with torch.no_grad():
            for x, g in zip(xs, pre):
                x -= 0.5 * g                                                       # :27
    assert values[0] == pytest.approx(4.0)
    assert values[-1] < 1e-6 and min(values) < 1e-8, (values[100], values[-1])
    assert abs(xs[0].item() - 1.0) < 1e-3 and abs(xs[1].item() - 1.0) < 1e-3     # the minimiser of the Rosenbrock function


@pytest.mark.parametrize("case", ["dense", "kron1", "kron2", "kron3"])
def test_tensor_decomposition_demo(psgd, case):
    """demo_usage_of_all_preconditioners.py:7-20, :26-41 (dense) and :65-93 (Kronecker examples 1-3)."""
    torch.manual_seed(1)
    I, J, K, R = 10, 20, 50, 5
    dev = "cuda"
    T = torch.rand(I, J, K, device=dev)
    xyz = [torch.randn(R, n, device=dev, requires_grad=True) for n in (I, J, K)]

    def f():
        x, y, z = xyz
        err = T - torch.einsum("ri,rj,rk->ijk", x, y, z)
        return (err * err).sum() + 1e-3 * sum(w.abs().sum() for w in xyz)

    eye, ones, zeros = (lambda n: torch.eye(n, device=dev)), (lambda *s: torch.ones(*s, device=dev)), (lambda *s: torch.zeros(*s, device=dev))
    norm = lambda n: torch.stack([ones(n), zeros(n)])
    if case == "kron1":      # (dense, normalization), (scaling, dense), (scaling, normalization)        :68-70
        Qs = [[0.1 * eye(R), norm(I)], [0.1 * ones(1, R), eye(J)], [0.1 * ones(1, R), norm(K)]]
    elif case == "kron2":    # (normalization, dense), (dense, scaling), (normalization, scaling)        :73-75
        Qs = [[0.1 * norm(R), eye(I)], [0.1 * eye(R), ones(1, J)], [0.1 * norm(R), ones(1, K)]]
    elif case == "kron3":    # (dense, dense) everywhere                                                  :78
        Qs = [[0.1 * eye(w.shape[0]), eye(w.shape[1])] for w in xyz]
    else:
        Q = 0.1 * eye(sum(w.numel() for w in xyz))                                                        # :27-28
    values = []
    for _ in range(100):
        cost = f()
        grads = torch.autograd.grad(cost, xyz, create_graph=True)
        vs = [torch.randn_like(w) for w in xyz]
        hess_vs = torch.autograd.grad(grads, xyz, vs)
        values.append(cost.item())
        grads = [g.detach() for g in grads]
        if case == "dense":
            Q = psgd.update_precond_dense(Q, vs, hess_vs, step=0.1)                                       # :38
            pre = psgd.precond_grad_dense(Q, grads)                                                       # :39
        else:
            new = psgd.update_precond_kron_batched([q[0] for q in Qs], [q[1] for q in Qs], vs, list(hess_vs), 0.1)   # :88
            Qs = [list(q) for q in new]
            pre = psgd.precond_grad_kron_batched([q[0] for q in Qs], [q[1] for q in Qs], grads)           # :90
        psgd.apply_preconditioned_updates([w.data for w in xyz], pre, 0.1)                                # :40 / :91
    assert np.isfinite(values).all()
    assert values[-1] < 1100.0 and values[-1] < 0.05 * values[0], (values[0], values[50], values[-1])


def test_uvd_state_dict_round_trip(psgd):
    """Caller-owned state: save / restore the optimiser state and hyper-parameters, continue bit-identically."""
    torch.manual_seed(3)
    mk = lambda: [torch.randn(12, 7, device="cuda").requires_grad_(), torch.randn(7, device="cuda").requires_grad_()]
    p1 = mk()
    opt1 = psgd.UVd(p1, rank_of_modification=3, lr_params=0.05, grad_clip_max_norm=2.0)
    g = [torch.randn_like(p) for p in p1]; v = [torch.randn_like(p) for p in p1]; h = [1.2 * x for x in v]
    opt1.step_with(g, v, h, balance=False, update_U=True)
    sd = {k: (t.clone() if isinstance(t, torch.Tensor) else t) for k, t in opt1.state_dict().items()}
    p2 = [torch.zeros_like(p).requires_grad_() for p in p1]
    opt2 = psgd.UVd(p2, rank_of_modification=3)
    opt2.load_state_dict(sd)
    assert float(opt2.lr_params) == 0.05 and float(opt2.grad_clip_max_norm) == 2.0
    opt1.step_with(g, v, h, balance=False, update_U=False)
    opt2.step_with(g, v, h, balance=False, update_U=False)
    for a, b in zip(p1, p2):
        assert torch.equal(a, b)
    assert torch.equal(opt1._U, opt2._U) and torch.equal(opt1._V, opt2._V) and torch.equal(opt1._d, opt2._d)
