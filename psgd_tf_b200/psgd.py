"""Drop-in functional API of lixilinx/psgd_tf's ``preconditioned_stochastic_gradient_descent.py``
(``psgd.py`` below), executed by hand-written sm_100a CUDA behind ``include/psgd_b200.h``.

Same names, argument order, defaults and value semantics as the reference:

* Kron / dense functions are *functional*: inputs untouched, new tensors returned (psgd.py:42, :179);
* ``update_precond_UVd_math_`` updates U, V, d *in place* and returns None (psgd.py:554-617);
* an unknown Kronecker factor combination prints a warning and returns its inputs (psgd.py:89-91).

Tensors are exchanged zero-copy through DLPack: any object with ``__dlpack__`` living on a CUDA device
(torch, TF via ``tf.experimental.dlpack``, cupy, jax ...) is viewed as a ``torch.Tensor`` without a copy;
results come back as torch CUDA tensors (``tf_adapter.py`` converts them to TF).  torch is plumbing only
(device memory, streams); every arithmetic operation runs in this package's CUDA kernels.  There is no
CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import math
import random as _random
import threading

import torch

from . import _lib
from ._lib import FACTOR_DENSE, FACTOR_NORM, FACTOR_SCALE, KronLayer, PsgdError, check

dtype = torch.float32                      # psgd.py:20
_tiny = float(2.0 ** -126)                 # psgd.py:21-22: smallest normal float32

# Capability limits of the CUDA kernels (the reference accepts any rank): the UVd sweeps are instantiated per rank
# with register-resident Gram accumulators (csrc/uvd.cu, kMaxRank), the sparse-LU corner kernels run in one warp
# (csrc/splu.cu, kMaxR).  Checked up front so that a too-large rank fails before any state is touched.
MAX_UVD_RANK = 16
MAX_SPLU_RANK = 32

_ctx_local = threading.local()
_rng = _random.Random()


def seed(s: int) -> None:
    """Seed the host RNG that draws the two coin flips of ``update_precond_UVd_math_``."""
    _rng.seed(s)


# ---------------------------------------------------------------------------------------------
# plumbing
# ---------------------------------------------------------------------------------------------
def get_context(device: int | None = None) -> _lib.Context:
    """The calling thread's context for ``device`` (created on first use)."""
    if not torch.cuda.is_available():
        raise RuntimeError("psgd_tf_b200 needs a CUDA device (sm_100a); there is no CPU path")
    if device is None:
        device = torch.cuda.current_device()
    table = getattr(_ctx_local, "table", None)
    if table is None:
        table = _ctx_local.table = {}
    ctx = table.get(device)
    if ctx is None:
        ctx = table[device] = _lib.Context(device)
    ctx.set_stream(torch.cuda.current_stream(device).cuda_stream)
    return ctx


def _as_tensor(x, name: str) -> torch.Tensor:
    """Zero-copy view of a DLPack-capable CUDA array as a contiguous float32 torch tensor."""
    if not isinstance(x, torch.Tensor):
        if hasattr(x, "__dlpack__"):
            x = torch.from_dlpack(x)
        else:
            raise TypeError(f"{name}: expected a tensor supporting DLPack, got {type(x).__name__}")
    if not x.is_cuda:
        raise RuntimeError(f"{name}: tensor is on {x.device}; psgd_tf_b200 runs on CUDA only (no CPU fallback)")
    if x.dtype != torch.float32:
        raise TypeError(f"{name}: dtype {x.dtype} is not float32 (psgd.py:20, README.md:37)")
    return x


def _in(x, name: str) -> torch.Tensor:
    t = _as_tensor(x, name)
    return t if t.is_contiguous() else t.contiguous()


def _inplace(x, name: str) -> torch.Tensor:
    t = _as_tensor(x, name)
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: in-place state tensors must be contiguous")
    return t


def _p(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


def _scalar(x) -> float:
    return float(x.item()) if hasattr(x, "item") else float(x)


def _flat_cat(ts):
    """Flatten and concatenate a list of tensors in list order (psgd.py:34-35); a single tensor is viewed, not copied."""
    ts = [t.reshape(-1) for t in ts]
    return ts[0] if len(ts) == 1 else torch.cat(ts)


def _check_uvd_rank(r: int, what: str) -> None:
    if not 1 <= int(r) <= MAX_UVD_RANK:
        raise ValueError(f"{what}: rank_of_modification {r} is outside 1..{MAX_UVD_RANK}, the range the B200 UVd kernels "
                         "are built for (the reference accepts any rank; see psgd_tf_b200.MAX_UVD_RANK)")


def _kind(Q: torch.Tensor) -> int:
    """Factor format from its shape, in the reference's test order: square => dense first
    (psgd.py:82-83), then first dim 2 => normalization, 1 => scaling."""
    m, n = Q.shape
    if m == n:
        return FACTOR_DENSE
    if m == 2:
        return FACTOR_NORM
    if m == 1:
        return FACTOR_SCALE
    return -1


_SUPPORTED = {(FACTOR_DENSE, FACTOR_DENSE), (FACTOR_DENSE, FACTOR_NORM), (FACTOR_DENSE, FACTOR_SCALE),
              (FACTOR_NORM, FACTOR_DENSE), (FACTOR_NORM, FACTOR_SCALE), (FACTOR_SCALE, FACTOR_DENSE),
              (FACTOR_SCALE, FACTOR_NORM)}


# ---------------------------------------------------------------------------------------------
# dense (full matrix) preconditioner                                        psgd.py:26-63
# ---------------------------------------------------------------------------------------------
def update_precond_dense(Q, dxs, dgs, step=0.01):
    """psgd.py:26-42.  ``dxs``/``dgs``: lists of arbitrarily shaped tensors, flattened and concatenated
    in list order (psgd.py:34-35)."""
    Q = _in(Q, "Q")
    dx = _flat_cat([_in(x, "dxs") for x in dxs])
    dg = _flat_cat([_in(g, "dgs") for g in dgs])
    n = Q.shape[0]
    if Q.dim() != 2 or Q.shape != (n, n) or dx.numel() != n or dg.numel() != n:
        raise ValueError(f"update_precond_dense: Q {tuple(Q.shape)} vs {dx.numel()} parameters")
    out = torch.empty_like(Q)
    ctx = get_context(Q.device.index)
    check(ctx.lib.psgd_dense_update(ctx.handle, _p(Q), _p(dx), _p(dg), _p(out), n, _scalar(step), _tiny))
    return out


def precond_grad_dense(Q, grads):
    """psgd.py:45-63: ``Q^T Q g`` reshaped back to the shapes of ``grads``."""
    Q = _in(Q, "Q")
    gs = [_in(g, "grads") for g in grads]
    flat = _flat_cat(gs)
    n = Q.shape[0]
    if Q.dim() != 2 or Q.shape != (n, n) or flat.numel() != n:
        raise ValueError(f"precond_grad_dense: Q {tuple(Q.shape)} vs {flat.numel()} parameters")
    out = torch.empty_like(flat)
    ctx = get_context(Q.device.index)
    check(ctx.lib.psgd_dense_apply(ctx.handle, _p(Q), _p(flat), _p(out), n))
    pre, idx = [], 0
    for g in gs:
        pre.append(out[idx: idx + g.numel()].reshape(g.shape))
        idx += g.numel()
    return pre


# ---------------------------------------------------------------------------------------------
# sparse LU preconditioner                                                   psgd.py:396-524
# ---------------------------------------------------------------------------------------------
def _splu_check(L12, l3, U12, u3, n):
    r = U12.shape[0]
    if not 1 <= r <= MAX_SPLU_RANK:
        raise ValueError(f"SPLU: r = {r} is outside 1..{MAX_SPLU_RANK}, the range the B200 kernels are built for")
    if L12.shape != (n, r) or U12.shape != (r, n) or l3.numel() != n - r or u3.numel() != n - r:
        raise ValueError(f"SPLU: L12 {tuple(L12.shape)}, l3 {tuple(l3.shape)}, U12 {tuple(U12.shape)}, u3 {tuple(u3.shape)} "
                         f"do not describe a {n}-parameter preconditioner")
    return r


def update_precond_splu(L12, l3, U12, u3, dxs, dgs, step=0.01):
    """psgd.py:396-477: Q = L U with L = [L1 0; L2 diag(l3)], U = [U1 U2; 0 diag(u3)].  Functional: returns
    ``[L12', l3', U12', u3']`` (psgd.py:480)."""
    L12, l3, U12, u3 = _in(L12, "L12"), _in(l3, "l3"), _in(U12, "U12"), _in(u3, "u3")
    dx = _flat_cat([_in(x, "dxs") for x in dxs])                               # psgd.py:426
    dg = _flat_cat([_in(g, "dgs") for g in dgs])                               # psgd.py:427
    n = dx.numel()
    if dg.numel() != n:
        raise ValueError("update_precond_splu: dxs and dgs differ in size")
    r = _splu_check(L12, l3, U12, u3, n)
    outs = [torch.empty_like(t) for t in (L12, l3, U12, u3)]
    ctx = get_context(L12.device.index)
    check(ctx.lib.psgd_splu_update(ctx.handle, _p(L12), _p(l3), _p(U12), _p(u3), _p(dx), _p(dg), *[_p(o) for o in outs],
                                   n, r, _scalar(step), _tiny))
    return outs


def precond_grad_splu(L12, l3, U12, u3, grads):
    """psgd.py:483-524: ``U^T L^T L U g`` reshaped back to the shapes of ``grads``."""
    L12, l3, U12, u3 = _in(L12, "L12"), _in(l3, "l3"), _in(U12, "U12"), _in(u3, "u3")
    gs = [_in(g, "grads") for g in grads]
    flat = _flat_cat(gs)                                                       # psgd.py:495-497
    n = flat.numel()
    r = _splu_check(L12, l3, U12, u3, n)
    out = torch.empty_like(flat)
    ctx = get_context(L12.device.index)
    check(ctx.lib.psgd_splu_apply(ctx.handle, _p(L12), _p(l3), _p(U12), _p(u3), _p(flat), _p(out), n, r))
    pre, idx = [], 0
    for g in gs:                                                               # psgd.py:518-522
        pre.append(out[idx: idx + g.numel()].reshape(g.shape))
        idx += g.numel()
    return pre


# ---------------------------------------------------------------------------------------------
# Kronecker product preconditioners                                          psgd.py:67-391
# ---------------------------------------------------------------------------------------------
_FACTOR_ROWS = {FACTOR_NORM: 2, FACTOR_SCALE: 1}


def _check_kron_layer(what, kl, kr, Ql, Qr, X, dG=None, layer=None):
    """Shape / device consistency of one layer (the kernels index Ql as [*, M], Qr as [*, N] and X as [M, N]; a
    mismatch would read or write out of bounds).  Called by the single-layer AND the batched entry points."""
    at = "" if layer is None else f" (layer {layer})"
    names = (("Ql", Ql), ("Qr", Qr), ("dX" if dG is not None else "Grad", X)) + ((("dG", dG),) if dG is not None else ())
    for nm, t in names:
        if t.dim() != 2:
            raise ValueError(f"{what}{at}: {nm} must be rank-2 (psgd.py:67-70, :113-115), got shape {tuple(t.shape)}")
        if t.device != X.device:
            raise ValueError(f"{what}{at}: {nm} is on {t.device}, expected {X.device}")
    M, N = X.shape
    ok = Ql.shape[1] == M and Qr.shape[1] == N and (dG is None or dG.shape == X.shape)
    ok = ok and Ql.shape[0] == _FACTOR_ROWS.get(kl, M) and Qr.shape[0] == _FACTOR_ROWS.get(kr, N)
    if not ok:
        raise ValueError(f"{what}{at}: Ql {tuple(Ql.shape)}, Qr {tuple(Qr.shape)}, {names[2][0]} {tuple(X.shape)}"
                         + (f", dG {tuple(dG.shape)}" if dG is not None else "") + " are inconsistent")


def _kron_update(Ql, Qr, dX, dG, step, kinds=None):
    Ql, Qr, dX, dG = _in(Ql, "Ql"), _in(Qr, "Qr"), _in(dX, "dX"), _in(dG, "dG")
    for t, nm in ((Ql, "Ql"), (Qr, "Qr"), (dX, "dX"), (dG, "dG")):
        if t.dim() != 2:
            raise ValueError(f"{nm} must be rank-2 (psgd.py:67-70), got shape {tuple(t.shape)}")
    kl, kr = kinds if kinds is not None else (_kind(Ql), _kind(Qr))
    M, N = dX.shape
    if (kl, kr) not in _SUPPORTED:
        print("Unknown Kronecker product preconditioner, no update")          # psgd.py:90
        return Ql, Qr
    _check_kron_layer("update_precond_kron", kl, kr, Ql, Qr, dX, dG)
    Ql_out, Qr_out = torch.empty_like(Ql), torch.empty_like(Qr)
    ctx = get_context(Ql.device.index)
    check(ctx.lib.psgd_kron_update(ctx.handle, kl, kr, _p(Ql), _p(Qr), _p(dX), _p(dG), _p(Ql_out), _p(Qr_out),
                                   M, N, _scalar(step), _tiny))
    return Ql_out, Qr_out


def _kron_apply(Ql, Qr, Grad, kinds=None):
    Ql, Qr, Grad = _in(Ql, "Ql"), _in(Qr, "Qr"), _in(Grad, "Grad")
    for t, nm in ((Ql, "Ql"), (Qr, "Qr"), (Grad, "Grad")):
        if t.dim() != 2:
            raise ValueError(f"{nm} must be rank-2 (psgd.py:113-115), got shape {tuple(t.shape)}")
    kl, kr = kinds if kinds is not None else (_kind(Ql), _kind(Qr))
    M, N = Grad.shape
    if (kl, kr) not in _SUPPORTED:
        print("Unknown Kronecker product preconditioner, no preconditioning")  # psgd.py:132
        return Grad
    _check_kron_layer("precond_grad_kron", kl, kr, Ql, Qr, Grad)
    out = torch.empty_like(Grad)
    ctx = get_context(Ql.device.index)
    check(ctx.lib.psgd_kron_apply(ctx.handle, kl, kr, _p(Ql), _p(Qr), _p(Grad), _p(out), M, N))
    return out


def update_precond_kron(Ql, Qr, dX, dG, step=0.01):
    """psgd.py:72-110: shape-dispatched update of ``P = kron(Qr^T Qr, Ql^T Ql)``; returns ``(Ql', Qr')``."""
    return _kron_update(Ql, Qr, dX, dG, step)


def precond_grad_kron(Ql, Qr, Grad):
    """psgd.py:116-152: ``Ql^T Ql Grad Qr^T Qr`` for any supported factor-format combination."""
    return _kron_apply(Ql, Qr, Grad)


def _update_precond_dense_dense(Ql, Qr, dX, dG, step=0.01):
    """psgd.py:156-179."""
    return _kron_update(Ql, Qr, dX, dG, step, (FACTOR_DENSE, FACTOR_DENSE))


def _precond_grad_dense_dense(Ql, Qr, Grad):
    """psgd.py:182-192."""
    return _kron_apply(Ql, Qr, Grad, (FACTOR_DENSE, FACTOR_DENSE))


def _update_precond_norm_dense(ql, Qr, dX, dG, step=0.01):
    """psgd.py:198-246."""
    return _kron_update(ql, Qr, dX, dG, step, (FACTOR_NORM, FACTOR_DENSE))


def _precond_grad_norm_dense(ql, Qr, Grad):
    """psgd.py:249-270."""
    return _kron_apply(ql, Qr, Grad, (FACTOR_NORM, FACTOR_DENSE))


def _update_precond_dense_scale(Ql, qr, dX, dG, step=0.01):
    """psgd.py:276-307."""
    return _kron_update(Ql, qr, dX, dG, step, (FACTOR_DENSE, FACTOR_SCALE))


def _precond_grad_dense_scale(Ql, qr, Grad):
    """psgd.py:310-322."""
    return _kron_apply(Ql, qr, Grad, (FACTOR_DENSE, FACTOR_SCALE))


def _update_precond_norm_scale(ql, qr, dX, dG, step=0.01):
    """psgd.py:328-369."""
    return _kron_update(ql, qr, dX, dG, step, (FACTOR_NORM, FACTOR_SCALE))


def _precond_grad_norm_scale(ql, qr, Grad):
    """psgd.py:372-391."""
    return _kron_apply(ql, qr, Grad, (FACTOR_NORM, FACTOR_SCALE))


# ---- batched ("all layers in one call") forms: new API surface ----------------------
def _layer_array(ctx_fields):
    arr = (KronLayer * len(ctx_fields))()
    for i, f in enumerate(ctx_fields):
        for k, v in f.items():
            setattr(arr[i], k, v)
    return arr


def update_precond_kron_batched(Qls, Qrs, dXs, dGs, step=0.01, outs=None):
    """Update every layer's factor pair in one library call.  Lists of tensors in, list of ``(Ql', Qr')`` out.
    Equivalent to ``[update_precond_kron(*a, step) for a in zip(Qls, Qrs, dXs, dGs)]``
    (mnist_with_lenet5.py:51).  ``outs``: optional list of preallocated ``(Ql', Qr')`` buffers (they must not alias the
    inputs), so that a caller can ping-pong two state sets without allocating -- what CUDA-graph replay needs."""
    n = len(Qls)
    if not (len(Qrs) == len(dXs) == len(dGs) == n):
        raise ValueError("update_precond_kron_batched: the four lists differ in length")
    Qls = [_in(q, "Ql") for q in Qls]; Qrs = [_in(q, "Qr") for q in Qrs]
    dXs = [_in(x, "dX") for x in dXs]; dGs = [_in(g, "dG") for g in dGs]
    given = outs
    if given is not None and len(given) != n:
        raise ValueError("update_precond_kron_batched: outs differs in length")
    outs, fields, keep = [None] * n, [], []
    for i in range(n):
        if Qls[i].dim() != 2 or Qrs[i].dim() != 2:
            raise ValueError(f"update_precond_kron_batched (layer {i}): factors must be rank-2")
        kl, kr = _kind(Qls[i]), _kind(Qrs[i])
        if (kl, kr) not in _SUPPORTED:
            print("Unknown Kronecker product preconditioner, no update")
            outs[i] = (Qls[i], Qrs[i])
            continue
        _check_kron_layer("update_precond_kron_batched", kl, kr, Qls[i], Qrs[i], dXs[i], dGs[i], layer=i)
        if dXs[i].device != dXs[0].device:
            raise ValueError(f"update_precond_kron_batched (layer {i}): all layers of a call must live on one device")
        M, N = dXs[i].shape
        if given is not None:
            lo, ro = _inplace(given[i][0], "Ql_out"), _inplace(given[i][1], "Qr_out")
            if lo.shape != Qls[i].shape or ro.shape != Qrs[i].shape or lo.data_ptr() == Qls[i].data_ptr() or \
                    ro.data_ptr() == Qrs[i].data_ptr():
                raise ValueError(f"update_precond_kron_batched (layer {i}): outs must match the factors' shapes and not alias them")
        else:
            lo, ro = torch.empty_like(Qls[i]), torch.empty_like(Qrs[i])
        outs[i] = (lo, ro)
        fields.append(dict(kind_l=kl, kind_r=kr, M=M, N=N, Ql=Qls[i].data_ptr(), Qr=Qrs[i].data_ptr(),
                           dX=dXs[i].data_ptr(), dG=dGs[i].data_ptr(), Ql_out=lo.data_ptr(), Qr_out=ro.data_ptr()))
    if fields:
        ctx = get_context(Qls[0].device.index)
        arr = _layer_array(fields)
        check(ctx.lib.psgd_kron_update_batched(ctx.handle, arr, len(fields), _scalar(step), _tiny))
    return outs


def precond_grad_kron_batched(Qls, Qrs, Grads, outs=None):
    """``[precond_grad_kron(Ql, Qr, G) for ...]`` in one library call (mnist_with_lenet5.py:53).  ``outs``: optional
    preallocated result buffers (not aliasing ``Grads``)."""
    n = len(Qls)
    if not (len(Qrs) == len(Grads) == n):
        raise ValueError("precond_grad_kron_batched: the three lists differ in length")
    Qls = [_in(q, "Ql") for q in Qls]; Qrs = [_in(q, "Qr") for q in Qrs]
    Grads = [_in(g, "Grad") for g in Grads]
    given = outs
    if given is not None and len(given) != n:
        raise ValueError("precond_grad_kron_batched: outs differs in length")
    outs, fields = [None] * n, []
    for i in range(n):
        if Qls[i].dim() != 2 or Qrs[i].dim() != 2:
            raise ValueError(f"precond_grad_kron_batched (layer {i}): factors must be rank-2")
        kl, kr = _kind(Qls[i]), _kind(Qrs[i])
        if (kl, kr) not in _SUPPORTED:
            print("Unknown Kronecker product preconditioner, no preconditioning")
            outs[i] = Grads[i]
            continue
        _check_kron_layer("precond_grad_kron_batched", kl, kr, Qls[i], Qrs[i], Grads[i], layer=i)
        if Grads[i].device != Grads[0].device:
            raise ValueError(f"precond_grad_kron_batched (layer {i}): all layers of a call must live on one device")
        M, N = Grads[i].shape
        if given is not None:
            o = _inplace(given[i], "out")
            if o.shape != Grads[i].shape or o.data_ptr() == Grads[i].data_ptr():
                raise ValueError(f"precond_grad_kron_batched (layer {i}): outs must match Grads' shapes and not alias them")
        else:
            o = torch.empty_like(Grads[i])
        outs[i] = o
        fields.append(dict(kind_l=kl, kind_r=kr, M=M, N=N, Ql=Qls[i].data_ptr(), Qr=Qrs[i].data_ptr(),
                           G=Grads[i].data_ptr(), out=o.data_ptr()))
    if fields:
        ctx = get_context(Qls[0].device.index)
        arr = _layer_array(fields)
        check(ctx.lib.psgd_kron_apply_batched(ctx.handle, arr, len(fields)))
    return outs


# ---- caller-side tail of a Kron training step: one launch per phase for the whole layer list ----
def apply_preconditioned_updates(Ws, pre_grads, lr, grad_norm_clip_thr=None, vs=None):
    """``W -= lr_adjust*lr*g`` for every layer, in place, with the reference's optional global clipping
    ``lr_adjust = min(grad_norm_clip_thr / sqrt(sum_l sum(g_l^2)), 1)`` (mnist_with_lenet5.py:54-56); ``vs``: the
    finite-difference perturbations still sitting on the parameters, removed by the same pass
    (neural_machine_translation_with_attention.py:206: ``W.assign_sub(lr*g + v)``).  The norm never visits the host."""
    from ._lib import ParamUpdate
    Ws = [_inplace(w, "W") for w in Ws]
    gs = [_in(g, "pre_grad") for g in pre_grads]
    vv = [_in(v, "v") for v in vs] if vs is not None else [None] * len(Ws)
    if not (len(Ws) == len(gs) == len(vv)):
        raise ValueError("apply_preconditioned_updates: lists differ in length")
    arr = (ParamUpdate * max(len(Ws), 1))()
    for i, (w, g, v) in enumerate(zip(Ws, gs, vv)):
        if g.numel() != w.numel() or (v is not None and v.numel() != w.numel()):
            raise ValueError(f"apply_preconditioned_updates: layer {i} shapes differ")
        arr[i].W, arr[i].pre, arr[i].v, arr[i].count = w.data_ptr(), g.data_ptr(), (v.data_ptr() if v is not None else None), w.numel()
    if not Ws:
        return None
    ctx = get_context(Ws[0].device.index)
    clip = float("inf") if grad_norm_clip_thr is None else _scalar(grad_norm_clip_thr)
    check(ctx.lib.psgd_apply_updates(ctx.handle, arr, len(Ws), _scalar(lr), clip))
    return None


def grad_differences(perturbed_grads, grads):
    """``[pg - g for ...]`` in one launch: finite-difference Hessian-vector products
    (neural_machine_translation_with_attention.py:200)."""
    from ._lib import DiffItem
    a = [_in(x, "perturbed_grad") for x in perturbed_grads]
    b = [_in(x, "grad") for x in grads]
    outs = [torch.empty_like(x) for x in a]
    if not a:
        return outs
    arr = (DiffItem * len(a))()
    for i, (x, y, o) in enumerate(zip(a, b, outs)):
        if x.shape != y.shape:
            raise ValueError(f"grad_differences: layer {i} shapes differ")
        arr[i].a, arr[i].b, arr[i].out, arr[i].count = x.data_ptr(), y.data_ptr(), o.data_ptr(), x.numel()
    ctx = get_context(a[0].device.index)
    check(ctx.lib.psgd_multi_sub(ctx.handle, arr, len(a)))
    return outs


# ---------------------------------------------------------------------------------------------
# UVd: Q = (I + U V^T) diag(d)                                               psgd.py:540-627
# ---------------------------------------------------------------------------------------------
def _col(x: torch.Tensor, n: int, name: str) -> torch.Tensor:
    if x.numel() != n:
        raise ValueError(f"{name}: expected {n} elements, got shape {tuple(x.shape)}")
    return x


def IpUVtmatvec(U, V, x):
    """psgd.py:540-544: ``x + U (V^T x)`` for a column vector or an [N, k] matrix ``x``."""
    U, V, x = _in(U, "U"), _in(V, "V"), _in(x, "x")
    n, r = U.shape
    k = 1 if x.dim() == 1 else x.shape[1]
    if V.shape != U.shape or x.shape[0] != n:
        raise ValueError("IpUVtmatvec: shapes are inconsistent")
    _check_uvd_rank(r, "IpUVtmatvec")
    out = torch.empty_like(x)
    ctx = get_context(U.device.index)
    check(ctx.lib.psgd_ipuvt_matvec(ctx.handle, _p(U), _p(V), _p(x), _p(out), n, r, k))
    return out


def update_precond_UVd_math_(U, V, d, v, h, step, tiny=_tiny, *, balance=None, update_U=None):
    """psgd.py:554-617.  Updates U, V ([N, r]) and d ([N, 1]) in place; returns None.

    The reference draws two coin flips from TF's RNG inside this function (psgd.py:562, :588).  Here they
    come from this module's host RNG (see :func:`seed`) unless given explicitly: ``balance`` stands for
    ``uniform() < 0.01`` and ``update_U`` for ``uniform() < 0.5``."""
    U, V, d = _inplace(U, "U"), _inplace(V, "V"), _inplace(d, "d")
    v, h = _in(v, "v"), _in(h, "h")
    n, r = U.shape
    if V.shape != U.shape:
        raise ValueError("update_precond_UVd_math_: U and V must have the same shape")
    _check_uvd_rank(r, "update_precond_UVd_math_")
    _col(d, n, "d"); _col(v, n, "v"); _col(h, n, "h")
    if balance is None:
        balance = _rng.random() < 0.01
    if update_U is None:
        update_U = _rng.random() < 0.5
    ctx = get_context(U.device.index)
    check(ctx.lib.psgd_uvd_update(ctx.handle, _p(U), _p(V), _p(d), _p(v), _p(h), n, r, _scalar(step), _scalar(tiny),
                                  int(bool(balance)), int(bool(update_U))))
    return None


def precond_grad_UVd_math(U, V, d, g):
    """psgd.py:619-627: ``d * (I + V U^T)(I + U V^T)(d * g)``; same shape as ``g``."""
    U, V, d, g = _in(U, "U"), _in(V, "V"), _in(d, "d"), _in(g, "g")
    n, r = U.shape
    if V.shape != U.shape:
        raise ValueError("precond_grad_UVd_math: U and V must have the same shape")
    _check_uvd_rank(r, "precond_grad_UVd_math")
    _col(d, n, "d"); _col(g, n, "g")
    out = torch.empty_like(g)
    ctx = get_context(U.device.index)
    check(ctx.lib.psgd_uvd_apply(ctx.handle, _p(U), _p(V), _p(d), _p(g), _p(out), n, r))
    return out


def update_precond_and_grad_UVd(U, V, d, v, h, g, step, tiny=_tiny, *, balance=None, update_U=None):
    """``update_precond_UVd_math_(U, V, d, v, h, step, tiny)`` followed by ``precond_grad_UVd_math(U, V, d, g)`` -- the
    sequence ``UVd.step`` runs (psgd.py:732-748) -- as ONE fused call (``psgd_uvd_update_apply``): three sweeps over
    U, V instead of five.  Updates U, V, d in place and returns the preconditioned gradient of the UPDATED
    preconditioner; same results as the two calls."""
    U, V, d = _inplace(U, "U"), _inplace(V, "V"), _inplace(d, "d")
    v, h, g = _in(v, "v"), _in(h, "h"), _in(g, "g")
    n, r = U.shape
    if V.shape != U.shape:
        raise ValueError("update_precond_and_grad_UVd: U and V must have the same shape")
    _check_uvd_rank(r, "update_precond_and_grad_UVd")
    _col(d, n, "d"); _col(v, n, "v"); _col(h, n, "h"); _col(g, n, "g")
    if balance is None:
        balance = _rng.random() < 0.01
    if update_U is None:
        update_U = _rng.random() < 0.5
    out = torch.empty_like(g)
    ctx = get_context(U.device.index)
    check(ctx.lib.psgd_uvd_update_apply(ctx.handle, _p(U), _p(V), _p(d), _p(v), _p(h), _p(g), _p(out), n, r,
                                        _scalar(step), _scalar(tiny), int(bool(balance)), int(bool(update_U))))
    return out


# north-star spellings (BASELINE.json) of the same two functions
update_precond_UVd = update_precond_UVd_math_
precond_grad_UVd = precond_grad_UVd_math


# ---------------------------------------------------------------------------------------------
# diagonal and X-shape preconditioners (README.md:11-15, :35) -- in place, like UVd
# ---------------------------------------------------------------------------------------------
def update_precond_diag(q, v, h, step=0.01, tiny=_tiny):
    """Q = diag(q): ``q -= mu * ((q h)^2 - (v/q)^2) * q``, ``mu = step / (max|.| + tiny)``.  In place."""
    q = _inplace(q, "q"); v, h = _in(v, "v"), _in(h, "h")
    n = q.numel()
    _col(v, n, "v"); _col(h, n, "h")
    ctx = get_context(q.device.index)
    check(ctx.lib.psgd_diag_update(ctx.handle, _p(q), _p(v), _p(h), n, _scalar(step), _scalar(tiny)))
    return None


def precond_grad_diag(q, g):
    """``q^2 * g``."""
    q, g = _in(q, "q"), _in(g, "g")
    _col(g, q.numel(), "g")
    out = torch.empty_like(g)
    ctx = get_context(q.device.index)
    check(ctx.lib.psgd_diag_apply(ctx.handle, _p(q), _p(g), _p(out), q.numel()))
    return out


def update_precond_Xmat(a, b, v, h, step=0.01, tiny=_tiny):
    """Q = diag(a) + adiag(b) (subgroup {e, flipping}, README.md:13).  Updates a, b in place."""
    a, b = _inplace(a, "a"), _inplace(b, "b"); v, h = _in(v, "v"), _in(h, "h")
    n = a.numel()
    _col(b, n, "b"); _col(v, n, "v"); _col(h, n, "h")
    ctx = get_context(a.device.index)
    check(ctx.lib.psgd_xmat_update(ctx.handle, _p(a), _p(b), _p(v), _p(h), n, _scalar(step), _scalar(tiny)))
    return None


def precond_grad_Xmat(a, b, g):
    """``Q^T Q g`` for Q = diag(a) + adiag(b)."""
    a, b, g = _in(a, "a"), _in(b, "b"), _in(g, "g")
    n = a.numel()
    _col(b, n, "b"); _col(g, n, "g")
    out = torch.empty_like(g)
    ctx = get_context(a.device.index)
    check(ctx.lib.psgd_xmat_apply(ctx.handle, _p(a), _p(b), _p(g), _p(out), n))
    return out


# ---------------------------------------------------------------------------------------------
# class UVd: the reference's stateful wrapper                                psgd.py:630-764
# ---------------------------------------------------------------------------------------------
class _Hyper:
    """A mutable scalar hyper-parameter with the ``.assign`` spelling the reference's callers use on its
    ``tf.Variable`` members (psgd.py:660-661 note 4; rnn_xor_UVd_preconditioner.py:62-69)."""

    def __init__(self, value):
        self.value = value

    def assign(self, value):
        self.value = value.value if isinstance(value, _Hyper) else value
        return self

    def numpy(self):
        return self.value

    def __float__(self):
        return float(self.value)

    def __bool__(self):
        return bool(self.value)

    def __repr__(self):
        return f"_Hyper({self.value!r})"


class UVd:
    """Low-rank modification (UVd) preconditioner as a class -- mirror of psgd.py:630-764 for callers whose parameters
    are CUDA torch tensors (torch.autograd plays the part of tf.GradientTape; it is plumbing, every preconditioner
    operation runs in this package's CUDA kernels).

    Same constructor arguments, hyper-parameter members (changed with ``.assign``, psgd.py note 4), state layout
    (``_U, _V: [N, r]``, ``_d: [N, 1]``, ``uv_scale = (1/(N r))^0.5``, psgd.py:684-690) and ``step(closure)`` contract as
    the reference.  Two additions:

    * the parameters are re-homed into ONE flat device buffer (each ``param.data`` becomes a view of it), so the tail
      of ``step`` -- preconditioned gradient, optional clipping, ``param -= lr * pre_grad (+ v)`` (psgd.py:747-762) --
      is the fused ``psgd_uvd_step_tail`` call: without clipping the preconditioned gradient is never materialised;
    * ``step_with(grads, vs, Hvs)`` exposes that hot path to callers that bring their own gradients and
      Hessian-vector products (another autodiff system, or a TF caller through DLPack).
    """

    def __init__(self, params_with_grad, rank_of_modification: int = 10, preconditioner_init_scale=1.0,
                 lr_params=0.01, lr_preconditioner=0.01, grad_clip_max_norm=None,
                 preconditioner_update_probability=1.0, exact_hessian_vector_product: bool = True):
        _check_uvd_rank(rank_of_modification, "UVd")         # before the parameters are re-homed into the flat buffer
        params = [params_with_grad] if isinstance(params_with_grad, torch.Tensor) else list(params_with_grad)
        flat_list = []
        for p in params:                                   # tf.nest.flatten (psgd.py:669)
            flat_list.extend(p if isinstance(p, (list, tuple)) else [p])
        self._params_with_grad = [p for p in flat_list if p.requires_grad]            # psgd.py:670
        if not self._params_with_grad:
            raise ValueError("UVd: no parameter requires gradients")
        for p in self._params_with_grad:
            _as_tensor(p, "params_with_grad")
        self._dtype = self._params_with_grad[0].dtype
        dev = self._params_with_grad[0].device
        self.lr_params = _Hyper(lr_params)
        self.lr_preconditioner = _Hyper(lr_preconditioner)
        self.grad_clip_max_norm = _Hyper(float("inf") if grad_clip_max_norm is None else grad_clip_max_norm)   # psgd.py:675-678
        self.preconditioner_update_probability = _Hyper(preconditioner_update_probability)
        self.exact_hessian_vector_product = _Hyper(bool(exact_hessian_vector_product))
        self._tiny = _tiny                                                             # psgd.py:682
        self._delta_param_scale = float(2.0 ** -23) ** 0.5                             # psgd.py:683: sqrt(eps)
        self._param_sizes = [p.numel() for p in self._params_with_grad]
        self._param_cumsizes = []
        tot = 0
        for s in self._param_sizes:
            tot += s
            self._param_cumsizes.append(tot)
        num_params = tot
        # one flat buffer holding every parameter; each param becomes a view (same values, same shapes)
        self._flat_params = torch.empty(num_params, device=dev, dtype=torch.float32)
        off = 0
        with torch.no_grad():
            for p, s in zip(self._params_with_grad, self._param_sizes):
                self._flat_params[off: off + s].copy_(p.detach().reshape(-1))
                p.data = self._flat_params[off: off + s].view(p.shape)
                off += s
        uv_scale = (1.0 / (num_params * rank_of_modification)) ** 0.5                 # psgd.py:687
        self._U = torch.randn(num_params, rank_of_modification, device=dev, dtype=torch.float32) * uv_scale
        self._V = torch.randn(num_params, rank_of_modification, device=dev, dtype=torch.float32) * uv_scale
        self._d = torch.ones(num_params, 1, device=dev, dtype=torch.float32) * preconditioner_init_scale

    # -- state (de)serialisation: the reference leaves checkpointing to the caller (state = plain variables) -------
    _HYPERS = ("lr_params", "lr_preconditioner", "grad_clip_max_norm", "preconditioner_update_probability",
               "exact_hessian_vector_product")

    def state_dict(self):
        """Preconditioner state, flat parameters and hyper-parameters (tensors are the live buffers, not copies)."""
        d = {"U": self._U, "V": self._V, "d": self._d, "params": self._flat_params}
        d.update({k: getattr(self, k).value for k in self._HYPERS})
        return d

    def load_state_dict(self, state):
        with torch.no_grad():
            for name, dst in (("U", self._U), ("V", self._V), ("d", self._d), ("params", self._flat_params)):
                src = state[name]
                if tuple(src.shape) != tuple(dst.shape):
                    raise ValueError(f"load_state_dict: {name} has shape {tuple(src.shape)}, expected {tuple(dst.shape)}")
                dst.copy_(src)
        for k in self._HYPERS:
            if k in state:
                getattr(self, k).assign(state[k])

    # -- the hot path, autodiff-agnostic -----------------------------------------------------------
    @staticmethod
    def _flatten(ts):
        ts = [_in(t, "tensor").reshape(-1) for t in ts]
        return ts[0] if len(ts) == 1 else torch.cat(ts)                               # psgd.py:729-730, :747

    def step_with(self, grads, vs=None, Hvs=None, params_perturbed: bool = False, *, balance=None, update_U=None,
                  return_pre_grad: bool = False):
        """Preconditioner update with ``(vs, Hvs)`` when given (psgd.py:729-736), then preconditioned gradient,
        optional clipping and parameter update (psgd.py:747-762).  ``grads``/``vs``/``Hvs``: lists shaped like the
        parameters, or single flat tensors.  ``params_perturbed``: the parameters currently hold ``param + v`` (the
        finite-difference mode, psgd.py:717-718), so ``v`` is removed again by the update (psgd.py:760-762)."""
        flat = lambda x: self._flatten(x if isinstance(x, (list, tuple)) else [x])
        n, r = self._U.shape
        g = flat(grads)
        _col(g, n, "grads")
        v = None
        if vs is not None:
            v, h = flat(vs), flat(Hvs)
            # compensate the levels of v and h in the finite-difference mode (psgd.py:734-736)
            vu, hu = (v / self._delta_param_scale, h / self._delta_param_scale) if params_perturbed else (v, h)
            if math.isinf(float(self.grad_clip_max_norm)):
                # update + preconditioned gradient as ONE fused call (three sweeps over U, V), then the parameter update
                # as a streaming pass; with clipping the norm-reducing tail below is used instead
                pre = update_precond_and_grad_UVd(self._U, self._V, self._d, vu, hu, g, float(self.lr_preconditioner),
                                                  self._tiny, balance=balance, update_U=update_U)
                apply_preconditioned_updates([self._flat_params], [pre], float(self.lr_params), None,
                                             [v] if params_perturbed else None)                       # psgd.py:757-762
                return pre if return_pre_grad else None
            update_precond_UVd_math_(self._U, self._V, self._d, vu, hu, float(self.lr_preconditioner), self._tiny,
                                     balance=balance, update_U=update_U)
        pre = torch.empty_like(g) if return_pre_grad else None
        ctx = get_context(self._U.device.index)
        vp = _p(v) if (params_perturbed and v is not None) else None
        check(ctx.lib.psgd_uvd_step_tail(ctx.handle, _p(self._U), _p(self._V), _p(self._d), _p(g), _p(self._flat_params),
                                         vp, _p(pre) if pre is not None else None, n, r, float(self.lr_params),
                                         float(self.grad_clip_max_norm), self._tiny))
        return pre

    # -- the reference's step(closure), with torch.autograd in the role of tf.GradientTape ------------------------
    def step(self, closure):
        """psgd.py:692-764.  ``closure`` evaluates the loss of the parameters (a scalar tensor, or an iterable whose
        first element is the loss); whatever it returns is handed back unchanged."""
        params = self._params_with_grad
        first = lambda ret: ret if isinstance(ret, torch.Tensor) else ret[0]
        if _rng.random() < float(self.preconditioner_update_probability):               # psgd.py:703
            if bool(self.exact_hessian_vector_product):                                # psgd.py:706-714
                with torch.enable_grad():
                    closure_returns = closure()
                    grads = torch.autograd.grad(first(closure_returns), params, create_graph=True)
                    vs = [torch.randn_like(p) for p in params]
                    Hvs = torch.autograd.grad(grads, params, vs)
                grads = [g.detach() for g in grads]
                self.step_with(grads, vs, Hvs, params_perturbed=False)
            else:                                                                      # psgd.py:715-727
                with torch.enable_grad():
                    closure_returns = closure()
                    grads = torch.autograd.grad(first(closure_returns), params)
                vs = [torch.randn_like(p) * self._delta_param_scale for p in params]
                # param += v and Hv = perturbed_grad - grad through the library's multi-tensor kernels (csrc/multi.cu)
                apply_preconditioned_updates([self._flat_params], [self._flatten(vs)], -1.0)          # psgd.py:717-718
                with torch.enable_grad():
                    perturbed_grads = torch.autograd.grad(first(closure()), params)
                Hvs = grad_differences(perturbed_grads, grads)                                          # psgd.py:725
                self.step_with(grads, vs, Hvs, params_perturbed=True)
        else:                                                                          # psgd.py:737-744
            with torch.enable_grad():
                closure_returns = closure()
                grads = torch.autograd.grad(first(closure_returns), params)
            self.step_with(grads)
        return closure_returns
