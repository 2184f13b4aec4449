"""Streaming rows of SURVEY.md section 8(a) that are not the two headline workloads: diagonal, X-shape, the (normalization,
scaling) Kronecker pair and the full-matrix apply.  Each is HBM bound; this reports update+apply steps/s and achieved
GB/s against the algorithmic byte counts of SURVEY.md section 8(d):

    diagonal   update 2*12N + 4N = 28N B, apply 12N B          (N = 1e8)
    X-shape    update 2*16N + 8N = 40N B, apply 16N B          (N = 1e8)
    norm,scale update 12MN B, apply 8MN B                      (NMT decoder shapes, SURVEY.md cfg5: [2305,1024], [1025,4935];
                                                                and a large [8192, 8192] layer so that the state exceeds L2)
    dense apply Q read twice = 8 n^2 B                         (n = 8192)
    SPLU       update 2*(8nr+16n) + 8nr + 8n = 24nr + 40n B,   (n = 5e7, r = 10; big inputs read twice -- reductions
               apply 2*(8nr+12n) + 4n = 16nr + 28n B            gate the maps -- outputs written once, like UVd)

Run by bench.py (nested under "aux" in the default JSON line) or directly:  python bench_aux.py
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _time(torch, fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run_aux(peak_gbs: float, steps: int = 10):
    import torch
    import psgd_tf_b200 as psgd

    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(7)
    rows = []

    def add(name, shape, ms_u, ms_a, bytes_u, bytes_a):
        tot = bytes_u + bytes_a
        rows.append(dict(path=name, shape=shape, update_ms=round(ms_u, 4), apply_ms=round(ms_a, 4),
                         steps_per_s=round(1e3 / (ms_u + ms_a), 2),
                         update_GBps=round(bytes_u / ms_u / 1e6, 1), apply_GBps=round(bytes_a / ms_a / 1e6, 1),
                         step_algorithmic_GB=round(tot / 1e9, 3), step_GBps=round(tot / (ms_u + ms_a) / 1e6, 1),
                         step_frac=round(tot / (ms_u + ms_a) / 1e6 / peak_gbs, 4)))

    # ---- diagonal and X-shape, N = 1e8 (12 x / 16 x 0.4 GB of state + inputs: far beyond L2) --------------------------
    N = 100_000_000
    v = torch.randn(N, device=dev, generator=g)
    h = (0.5 + 1.5 * torch.rand(N, device=dev, generator=g)) * v + 0.1 * torch.randn(N, device=dev, generator=g)
    gr = torch.randn(N, device=dev, generator=g)
    q = torch.ones(N, device=dev)
    ms_u = _time(torch, lambda: psgd.update_precond_diag(q, v, h, 0.01), steps)
    ms_a = _time(torch, lambda: psgd.precond_grad_diag(q, gr), steps)
    assert torch.isfinite(q).all()
    add("diagonal", [N], ms_u, ms_a, 28.0 * N, 12.0 * N)
    a = torch.ones(N, device=dev)
    b = torch.zeros(N, device=dev)
    ms_u = _time(torch, lambda: psgd.update_precond_Xmat(a, b, v, h, 0.01), steps)
    ms_a = _time(torch, lambda: psgd.precond_grad_Xmat(a, b, gr), steps)
    assert torch.isfinite(a).all() and torch.isfinite(b).all()
    add("X-shape", [N], ms_u, ms_a, 40.0 * N, 16.0 * N)
    del v, h, gr, q, a, b
    torch.cuda.empty_cache()

    # ---- (normalization, scaling) Kronecker pair ---------------------------------------------------------------------
    for (M, Nn) in ((2305, 1024), (1025, 4935), (8192, 8192)):
        ql = torch.stack([torch.ones(M, device=dev), torch.zeros(M, device=dev)])
        qr = torch.ones(1, Nn, device=dev)
        dX = torch.randn(M, Nn, device=dev, generator=g)
        dG = 1.3 * dX + 0.1 * torch.randn(M, Nn, device=dev, generator=g)
        G = torch.randn(M, Nn, device=dev, generator=g)
        state = [ql, qr]

        def upd():
            state[0], state[1] = psgd.update_precond_kron(state[0], state[1], dX, dG, 0.01)

        ms_u = _time(torch, upd, steps)
        ms_a = _time(torch, lambda: psgd.precond_grad_kron(state[0], state[1], G), steps)
        assert torch.isfinite(state[0]).all() and torch.isfinite(state[1]).all()
        add("kron (norm,scale)", [M, Nn], ms_u, ms_a, 12.0 * M * Nn, 8.0 * M * Nn)
        del dX, dG, G
    torch.cuda.empty_cache()

    # ---- dense full-matrix apply: two GEMVs, Q read twice ---------------------------------------------------------------
    n = 8192
    Q = torch.triu(torch.randn(n, n, device=dev, generator=g)) * 0.01 + torch.eye(n, device=dev)
    gv = [torch.randn(n, device=dev, generator=g)]
    ms_a = _time(torch, lambda: psgd.precond_grad_dense(Q, gv), steps)
    rows.append(dict(path="dense apply", shape=[n, n], apply_ms=round(ms_a, 4), apply_GBps=round(8.0 * n * n / ms_a / 1e6, 1),
                     apply_frac=round(8.0 * n * n / ms_a / 1e6 / peak_gbs, 4), step_algorithmic_GB=round(8.0 * n * n / 1e9, 3)))

    # ---- sparse-LU preconditioner (psgd.py:396-524), n = 5e7, r = 10 ---------------------------------------------------
    del Q, gv
    torch.cuda.empty_cache()
    n, r = 50_000_000, 10
    L12 = torch.cat([torch.eye(r, device=dev), torch.zeros(n - r, r, device=dev)])            # demo_usage_of_all_preconditioners.py:46-49
    U12 = torch.cat([torch.eye(r, device=dev), torch.zeros(r, n - r, device=dev)], 1)
    l3 = torch.ones(n - r, 1, device=dev)
    u3 = torch.ones(n - r, 1, device=dev)
    dx = torch.randn(n, device=dev, generator=g)
    dg = 1.3 * dx + 0.1 * torch.randn(n, device=dev, generator=g)
    gg = torch.randn(n, device=dev, generator=g)
    st = [L12, l3, U12, u3]

    def supd():
        st[0], st[1], st[2], st[3] = psgd.update_precond_splu(st[0], st[1], st[2], st[3], [dx], [dg], 0.01)

    ms_u = _time(torch, supd, steps)
    ms_a = _time(torch, lambda: psgd.precond_grad_splu(st[0], st[1], st[2], st[3], [gg]), steps)
    assert torch.isfinite(st[0]).all() and torch.isfinite(st[2]).all()
    add("SPLU", [n, r], ms_u, ms_a, (24.0 * r + 40.0) * n, (16.0 * r + 28.0) * n)
    return rows


if __name__ == "__main__":
    import torch
    from bench import load_peaks
    assert torch.cuda.is_available(), "bench_aux needs a CUDA device"
    for r in run_aux(load_peaks()["hbm"]):
        print(json.dumps(r))
