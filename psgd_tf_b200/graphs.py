"""CUDA-graph replay of the UVd update+apply step.

One step of the UVd path is a fixed chain of ~8 kernels (three sweeps + small r x r kernels in the fused
update+apply form, ``psgd_uvd_update_apply``; ~11 as two separate calls), plus one peer-exchange kernel per phase when
the vector is sharded over GPUs (csrc/comm.cu).  At 100 M
parameters on one GPU the chain runs ~7 ms and launch latency is noise; sharded 8 ways it runs < 1 ms and the ~12
launches + two Python/ctypes calls become a visible fraction.  The chain has no host dependency (coin flips are
arguments, the cross-rank epoch counter lives in device memory), so it is captured once per (input buffers, coin
flips) and replayed with a single ``cudaGraphLaunch``.

This is plumbing around the reference API, not a different algorithm: the graph contains exactly the kernels that
``update_precond_and_grad_UVd`` launches -- or, with ``fused=False``, those of ``update_precond_UVd_math_``
(psgd.py:554-617) followed by ``precond_grad_UVd_math`` (psgd.py:619-627).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import psgd as _psgd


class UVdStepGraphs:
    """``step(v, h, g, balance, update_U)`` == ``update_precond_UVd_math_(U, V, d, v, h, step, tiny, ...)`` followed by
    ``precond_grad_UVd_math(U, V, d, g)`` on the state tensors given at construction, replayed from a CUDA graph.

    Graphs are keyed by the input buffers' addresses and the two coin flips, so callers should cycle through a fixed
    set of (v, h, g) buffers (e.g. double-buffered uploads).  The first call for a shape runs eagerly (it sizes the
    library workspace); later calls capture on first use of a key and replay afterwards.  The returned tensor belongs
    to the graph: consume it before the same key is replayed again.
    """

    def __init__(self, U: torch.Tensor, V: torch.Tensor, d: torch.Tensor, step: float = 0.01, tiny: float = _psgd._tiny,
                 fused: bool = True):
        if not (U.is_cuda and V.is_cuda and d.is_cuda):
            raise RuntimeError("UVdStepGraphs: state must live on a CUDA device (no CPU path)")
        self.U, self.V, self.d = U, V, d
        self.step_size, self.tiny = float(step), float(tiny)
        self.fused = bool(fused)
        self._stream = torch.cuda.Stream(device=U.device)
        self._graphs: Dict[Tuple, Tuple[torch.cuda.CUDAGraph, torch.Tensor]] = {}
        self._warm = False
        self._ws_bytes = -1
        self.replays = 0
        self.kernel_launches = 0          # kernels executed through graph replays (the library's own counter sees only captures)

    def _eager(self, v, h, g, balance, update_U):
        if self.fused:
            return _psgd.update_precond_and_grad_UVd(self.U, self.V, self.d, v, h, g, self.step_size, self.tiny,
                                                     balance=balance, update_U=update_U)
        _psgd.update_precond_UVd_math_(self.U, self.V, self.d, v, h, self.step_size, self.tiny, balance=balance,
                                       update_U=update_U)
        return _psgd.precond_grad_UVd_math(self.U, self.V, self.d, g)

    def step(self, v: torch.Tensor, h: torch.Tensor, g: torch.Tensor, balance: bool, update_U: bool) -> torch.Tensor:
        cur = torch.cuda.current_stream(self.U.device)
        s = self._stream
        s.wait_stream(cur)
        with torch.cuda.stream(s):          # the library context follows torch's current stream
            ctx = _psgd.get_context(self.U.device.index)      # switch (and drain) streams BEFORE any capture begins
            if ctx.workspace_bytes != self._ws_bytes:
                # the library workspace was re-allocated by a larger call since capture: its address is baked into
                # the graphs, so they are stale
                self._graphs.clear()
                self._warm = False
            if not self._warm:
                out = self._eager(v, h, g, balance, update_U)
                self._warm = True
                self._ws_bytes = ctx.workspace_bytes
            else:
                key = (v.data_ptr(), h.data_ptr(), g.data_ptr(), bool(balance), bool(update_U))
                entry = self._graphs.get(key)
                if entry is None:
                    graph = torch.cuda.CUDAGraph()
                    before = ctx.launch_count
                    with torch.cuda.graph(graph, stream=s):
                        out = self._eager(v, h, g, balance, update_U)
                    entry = self._graphs[key] = (graph, out, ctx.launch_count - before)
                graph, out, nk = entry
                graph.replay()
                self.replays += 1
                self.kernel_launches += nk
        cur.wait_stream(s)
        return out


class KronStepGraphs:
    """CUDA-graph replay of one Kron training-step hot path over a (ragged) list of layers:
    ``update_precond_kron_batched`` followed by ``precond_grad_kron_batched`` with the UPDATED factors
    (mnist_with_lenet5.py:51-53, neural_machine_translation_with_attention.py:203-205).

    Small layers (LeNet5, the NMT model's factors) are launch-latency bound: a 7-pair NMT step is ~60 launches of a few
    microseconds each plus two Python/ctypes calls.  The chain has no host dependency, so it is captured once per set of
    input buffers and replayed with one ``cudaGraphLaunch``.  The functional API returns NEW factors every step; to keep
    addresses stable the factors ping-pong between two preallocated state sets (one graph per direction), so nothing is
    copied.  ``factors`` returns the current ``[(Ql, Qr), ...]``; ``step`` returns the preconditioned gradients, which
    belong to the graph (consume them before the same graph is replayed again)."""

    def __init__(self, Qls, Qrs, step: float = 0.01):
        if not all(q.is_cuda for q in list(Qls) + list(Qrs)):
            raise RuntimeError("KronStepGraphs: factors must live on a CUDA device (no CPU path)")
        self._state = [[(ql.clone(), qr.clone()) for ql, qr in zip(Qls, Qrs)],
                       [(torch.empty_like(ql), torch.empty_like(qr)) for ql, qr in zip(Qls, Qrs)]]
        self._cur = 0
        self.step_size = float(step)
        self._stream = torch.cuda.Stream(device=Qls[0].device)
        self._graphs: Dict[Tuple, Tuple] = {}
        self._warm = False
        self._ws_bytes = -1
        self.replays = 0
        self.kernel_launches = 0

    @property
    def factors(self):
        return self._state[self._cur]

    def _eager(self, src, dst, dXs, dGs, Gs, pre_out=None):
        Qls, Qrs = [a for a, _ in self._state[src]], [b for _, b in self._state[src]]
        new = _psgd.update_precond_kron_batched(Qls, Qrs, dXs, dGs, self.step_size, outs=self._state[dst])
        return _psgd.precond_grad_kron_batched([a for a, _ in new], [b for _, b in new], Gs, outs=pre_out)

    def step(self, dXs, dGs, Gs):
        dev = self._state[0][0][0].device
        cur = torch.cuda.current_stream(dev)
        s = self._stream
        s.wait_stream(cur)
        src, dst = self._cur, 1 - self._cur
        with torch.cuda.stream(s):
            ctx = _psgd.get_context(dev.index)
            if ctx.workspace_bytes != self._ws_bytes:
                self._graphs.clear()
                self._warm = False
            if not self._warm:
                pre = self._eager(src, dst, dXs, dGs, Gs)
                self._warm = True
                self._ws_bytes = ctx.workspace_bytes
            else:
                key = (src,) + tuple(t.data_ptr() for t in list(dXs) + list(dGs) + list(Gs))
                entry = self._graphs.get(key)
                if entry is None:
                    pre_out = [torch.empty_like(g) for g in Gs]
                    graph = torch.cuda.CUDAGraph()
                    before = ctx.launch_count
                    with torch.cuda.graph(graph, stream=s):
                        pre = self._eager(src, dst, dXs, dGs, Gs, pre_out)
                    entry = self._graphs[key] = (graph, pre, ctx.launch_count - before)
                graph, pre, nk = entry
                graph.replay()
                self.replays += 1
                self.kernel_launches += nk
        self._cur = dst
        cur.wait_stream(s)
        return pre
