"""GPU tests of the ``UVd.step`` tail (psgd.py:747-762) and of the ``class UVd`` mirror (psgd.py:630-764).

The tail is checked against the oracle restatement; the class against (i) the functional API it is built from and
(ii) convergence on the reference's own demo problem family (a delayed-XOR-like regression small enough for seconds).
"""
import math

import numpy as np
import pytest
import torch

from oracle import psgd_oracle as O
from tests import cases

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def psgd():
    import psgd_tf_b200 as p
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device (B200); run with -m gpu on the GPU box")
    p.get_context()
    return p


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize("n,r", [(1021, 10), (70_001, 10), (513, 4), (5, 2)])
@pytest.mark.parametrize("clip", [math.inf, 0.5, 1e6])
@pytest.mark.parametrize("with_v", [False, True])
def test_step_tail_matches_oracle(psgd, n, r, clip, with_v):
    from psgd_tf_b200._lib import check
    from psgd_tf_b200.psgd import _p
    c = cases.uvd_case(31 + n, n, r)
    rng = np.random.default_rng(n)
    params = rng.standard_normal((n, 1)).astype(np.float32)
    v = (rng.standard_normal((n, 1)) * 2.0 ** -11.5).astype(np.float32) if with_v else None
    want_p, want_pre = O.uvd_step_tail(c["U"], c["V"], c["d"], c["g"], params, 0.02, clip, v)
    ctx = psgd.get_context()
    U, V, d, g = dev(c["U"]), dev(c["V"]), dev(c["d"]), dev(c["g"])      # keep the device buffers alive across the call
    vd = dev(v) if with_v else None
    for want_pre_out in (False, True):
        P = dev(params)
        pre = torch.empty(n, 1, device="cuda") if want_pre_out else None
        check(ctx.lib.psgd_uvd_step_tail(ctx.handle, _p(U), _p(V), _p(d), _p(g), _p(P), _p(vd) if with_v else None,
                                         _p(pre) if want_pre_out else None, n, r, 0.02, clip, psgd._tiny))
        assert cases.rel_err(P.cpu().numpy(), want_p) <= TOL
        # the update itself (params - new) must be right too, not just the (much larger) parameters
        assert cases.rel_err(params - P.cpu().numpy(), params - want_p) <= 20 * TOL
        if want_pre_out:
            assert cases.rel_err(pre.cpu().numpy(), want_pre) <= TOL


def test_class_step_with_equals_functional_api(psgd):
    """UVd.step_with == update_precond_UVd_math_ + precond_grad_UVd_math + clip + assign_sub done by hand."""
    torch.manual_seed(0)
    W1 = torch.randn(30, 17, device="cuda", requires_grad=True)
    b1 = torch.randn(17, device="cuda", requires_grad=True)
    frozen = torch.randn(3, device="cuda")                       # not trainable: must be ignored (psgd.py:670)
    opt = psgd.UVd([W1, [b1, frozen]], rank_of_modification=5, preconditioner_init_scale=0.7, lr_params=0.05,
                   lr_preconditioner=0.02, grad_clip_max_norm=0.3)
    n = 30 * 17 + 17
    assert opt._U.shape == (n, 5) and opt._d.shape == (n, 1) and torch.all(opt._d == 0.7)
    assert W1.data_ptr() == opt._flat_params.data_ptr()           # parameters are views of the flat buffer
    U, V, d = opt._U.clone(), opt._V.clone(), opt._d.clone()
    p0 = opt._flat_params.clone()
    g = [torch.randn(30, 17, device="cuda"), torch.randn(17, device="cuda")]
    vs = [torch.randn(30, 17, device="cuda"), torch.randn(17, device="cuda")]
    hs = [1.5 * v + 0.1 * torch.randn_like(v) for v in vs]
    pre = opt.step_with(g, vs, hs, balance=False, update_U=True, return_pre_grad=True)
    flat = lambda ts: torch.cat([t.reshape(-1) for t in ts])[:, None].contiguous()
    psgd.update_precond_UVd_math_(U, V, d, flat(vs), flat(hs), 0.02, psgd._tiny, balance=False, update_U=True)
    want = psgd.precond_grad_UVd_math(U, V, d, flat(g))
    assert torch.equal(opt._U, U) and torch.equal(opt._d, d)
    assert cases.rel_err(pre.cpu().numpy().reshape(-1), want.cpu().numpy().reshape(-1)) <= 1e-6
    lr = 0.05 * min(0.3 / (want.norm().item() + psgd._tiny), 1.0)
    # the update is ~1e-3 of the parameters, so recovering it as a float32 difference costs ~1e-7/1e-3 of relative accuracy
    assert cases.rel_err((p0 - opt._flat_params).cpu().numpy(), (lr * want.reshape(-1)).cpu().numpy()) <= 5e-4
    assert cases.rel_err(opt._flat_params.cpu().numpy(), (p0 - lr * want.reshape(-1)).cpu().numpy()) <= 1e-6
    assert torch.equal(W1.detach().reshape(-1), opt._flat_params[:510])


@pytest.mark.parametrize("perturbed", [False, True])
def test_class_step_with_without_clipping_uses_fused_call(psgd, perturbed):
    """No clipping: step_with = the fused update+apply call + one streaming parameter pass; same results as the
    reference sequence update_precond_UVd_math_ -> precond_grad_UVd_math -> assign_sub (psgd.py:732-762)."""
    torch.manual_seed(3)
    W = torch.randn(4099, device="cuda", requires_grad=True)
    opt = psgd.UVd(W, rank_of_modification=10, lr_params=0.05, lr_preconditioner=0.02)
    assert math.isinf(float(opt.grad_clip_max_norm))
    n = 4099
    U, V, d = opt._U.clone(), opt._V.clone(), opt._d.clone()
    g = torch.randn(n, device="cuda")
    scale = opt._delta_param_scale if perturbed else 1.0
    v = torch.randn(n, device="cuda") * scale
    h = 1.5 * v + 0.1 * scale * torch.randn(n, device="cuda")
    if perturbed:
        with torch.no_grad():
            opt._flat_params.add_(v)                               # psgd.py:717-718
    p0 = opt._flat_params.clone()
    ctx = psgd.get_context()
    before = ctx.launch_count
    pre = opt.step_with([g], [v], [h], params_perturbed=perturbed, balance=False, update_U=False, return_pre_grad=True)
    assert ctx.launch_count - before == 6                         # 5 of the fused call (3 sweeps, 2 mid kernels) + the parameter pass
    col = lambda t: t[:, None].contiguous()
    psgd.update_precond_UVd_math_(U, V, d, col(v / scale), col(h / scale), 0.02, psgd._tiny, balance=False, update_U=False)
    want = psgd.precond_grad_UVd_math(U, V, d, col(g)).reshape(-1)
    assert torch.equal(opt._V, V) and torch.equal(opt._d, d) and torch.equal(opt._U, U)
    assert cases.rel_err(pre.cpu().numpy(), want.cpu().numpy()) <= 1e-6
    want_p = p0 - (0.05 * want + (v if perturbed else 0.0))
    assert cases.rel_err(opt._flat_params.cpu().numpy(), want_p.cpu().numpy()) <= 1e-6


@pytest.mark.parametrize("exact", [True, False])
def test_class_step_converges_on_small_regression(psgd, exact):
    """step(closure) with exact and finite-difference Hessian-vector products (psgd.py:706-727) drives a tiny
    two-layer tanh regression to a small loss, and un-perturbs the parameters in the finite-difference mode."""
    torch.manual_seed(1)
    psgd.seed(1)
    X = torch.randn(256, 8, device="cuda")
    Wt = torch.randn(8, 1, device="cuda")
    y = torch.tanh(X @ Wt)
    W1 = (0.3 * torch.randn(8, 16, device="cuda")).requires_grad_()
    W2 = (0.3 * torch.randn(16, 1, device="cuda")).requires_grad_()
    opt = psgd.UVd([W1, W2], rank_of_modification=4, lr_params=0.1, lr_preconditioner=0.05, grad_clip_max_norm=10.0,
                   preconditioner_update_probability=0.8, exact_hessian_vector_product=exact)

    def closure():
        return ((torch.tanh(X @ W1) @ W2 - y) ** 2).mean()

    first = closure().item()
    for _ in range(300):
        loss = opt.step(closure)
    assert torch.isfinite(loss)
    assert closure().item() < 0.1 * first, (first, closure().item())
    opt.lr_params.assign(0.0)                                   # psgd.py note 4: members are changed with .assign
    before = opt._flat_params.clone()
    opt.step(closure)
    # lr = 0: the step must leave the parameters where they were (incl. removing the FD perturbation, psgd.py:760-762)
    assert torch.allclose(opt._flat_params, before, atol=1e-6)
