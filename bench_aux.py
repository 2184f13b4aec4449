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

and the latency-bound configurations of BASELINE.json ("report us/step for one batched launch"):

    cfg1  LeNet5, five (dense, dense) pairs  W in {[26,6], [151,16], [257,120], [121,84], [85,10]}   (mnist_with_lenet5.py:12-16)
    cfg2  UVd rank 10 on N = 1021 parameters                                                         (rnn_xor_UVd_preconditioner.py:37-41)
    cfg5  NMT, seven pairs in ONE batched call: enc-emb [9414,256] (scale,dense), enc-rnn [1281,1024] (norm,scale),
          att-in [2048,10] (scale,dense), att-v [1,10] (dense,dense), dec-emb [4935,256] (scale,dense),
          dec-rnn [2305,1024] (norm,scale), dec-fc [1025,4935] (norm,scale)     (neural_machine_translation_with_attention.py:98-148)
each as one update+apply step launched eagerly from Python and replayed from a CUDA graph (psgd_tf_b200/graphs.py),
plus the two mixed dense/structured Kron pairs at a size that fills the GPU: (norm, dense) [16384, 1024] and
(dense, scale) [1024, 16384].

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
    # dense update (psgd.py:26-44): a = Q dg, b = Q^-T dx (vector solve), Q - mu triu(a a^T - b b^T) Q (one n^3 product)
    dxv = [torch.randn(n, device=dev, generator=g)]
    dgv = [1.3 * dxv[0] + 0.1 * torch.randn(n, device=dev, generator=g)]
    qs = [Q]

    def dupd():
        qs[0] = psgd.update_precond_dense(qs[0], dxv, dgv, 0.01)

    ms_u = _time(torch, dupd, max(3, steps // 4))
    assert torch.isfinite(qs[0]).all()
    rows.append(dict(path="dense update", shape=[n, n], update_ms=round(ms_u, 3),
                     update_TFLOPs_dense_count=round(2.0 * n ** 3 / ms_u / 1e9, 1)))
    del dxv, dgv, qs

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
    del L12, U12, l3, u3, dx, dg, gg, st
    torch.cuda.empty_cache()
    rows.extend(run_small_configs(peak_gbs, steps=max(steps, 20)))
    rows.extend(run_mixed_pairs(peak_gbs, steps=steps))
    return rows


def _factor(torch, kind, n, dev):
    """Identity-type initial factors (README.md:48; neural_machine_translation_with_attention.py:94-95)."""
    if kind == "dense":
        return torch.eye(n, device=dev)
    if kind == "norm":
        return torch.stack([torch.ones(n, device=dev), torch.zeros(n, device=dev)])
    return torch.ones(1, n, device=dev)


def _kron_step_bytes(kl, kr, M, N):
    """Compulsory HBM bytes of one update+apply of a pair: dX, dG, G read and the result written (16 MN), dX once more
    when a global reduction gates the statistics ((norm, .) pairs), dense factors read once per use."""
    b = 16.0 * M * N + (4.0 * M * N if kl == "norm" else 0.0)
    for k, n in ((kl, M), (kr, N)):
        if k == "dense":
            b += 4.0 * n * n * 4          # update: product + solve; apply: two products
    return b


def run_small_configs(peak_gbs: float, steps: int = 20):
    """cfg1 / cfg2 / cfg5: microseconds per update+apply step, eager and from a CUDA graph."""
    import torch
    import psgd_tf_b200 as psgd
    from psgd_tf_b200.graphs import KronStepGraphs, UVdStepGraphs

    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(11)
    ctx = psgd.get_context()
    rows = []

    def kron_config(name, cite, layers):
        Ql = [_factor(torch, kl, M, dev) for kl, kr, M, N in layers]
        Qr = [_factor(torch, kr, N, dev) for kl, kr, M, N in layers]
        sets = []
        for _ in range(2):
            dX = [torch.randn(M, N, device=dev, generator=g) for _, _, M, N in layers]
            dG = [1.3 * x + 0.1 * torch.randn(x.shape, device=dev, generator=g) for x in dX]
            G = [torch.randn(M, N, device=dev, generator=g) for _, _, M, N in layers]
            sets.append((dX, dG, G))
        state = [Ql, Qr]
        k = [0]

        def eager():
            dX, dG, G = sets[k[0] & 1]; k[0] += 1
            new = psgd.update_precond_kron_batched(state[0], state[1], dX, dG, 0.01)
            state[0], state[1] = [a for a, _ in new], [b for _, b in new]
            return psgd.precond_grad_kron_batched(state[0], state[1], G)

        l0 = ctx.launch_count
        eager()
        per_step_launches = ctx.launch_count - l0
        ms_e = _time(torch, eager, steps)
        gr = KronStepGraphs(state[0], state[1], 0.01)

        def graphed():
            dX, dG, G = sets[k[0] & 1]; k[0] += 1
            return gr.step(dX, dG, G)

        for _ in range(6):            # eager warm-up + capture of the four (direction, input set) graphs
            graphed()
        ms_g = _time(torch, graphed, steps)
        pre = graphed()
        assert all(torch.isfinite(p).all() for p in pre) and all(torch.isfinite(a).all() for a, _ in gr.factors)
        byts = sum(_kron_step_bytes(kl, kr, M, N) for kl, kr, M, N in layers)
        rows.append(dict(path=name, cite=cite, layers=[[kl, kr, M, N] for kl, kr, M, N in layers],
                         launches_per_step=int(per_step_launches), eager_us_per_step=round(ms_e * 1e3, 1),
                         graph_us_per_step=round(ms_g * 1e3, 1), steps_per_s=round(1e3 / ms_g, 1),
                         step_compulsory_MB=round(byts / 1e6, 2), graph_GBps=round(byts / ms_g / 1e6, 1),
                         graph_frac=round(byts / ms_g / 1e6 / peak_gbs, 4)))

    lenet = [("dense", "dense", M, N) for M, N in ((26, 6), (151, 16), (257, 120), (121, 84), (85, 10))]
    kron_config("cfg1 LeNet5 five (dense,dense) pairs, one batched call", "mnist_with_lenet5.py:12-16,51-53", lenet)
    nmt = [("scale", "dense", 9414, 256), ("norm", "scale", 1281, 1024), ("scale", "dense", 2048, 10),
           ("dense", "dense", 1, 10), ("scale", "dense", 4935, 256), ("norm", "scale", 2305, 1024),
           ("norm", "scale", 1025, 4935)]
    kron_config("cfg5 NMT seven mixed pairs, one batched call", "neural_machine_translation_with_attention.py:98-148", nmt)
    kron_config("cfg5 NMT, the three (norm,scale) pairs only", "neural_machine_translation_with_attention.py:118,138,146",
                [l for l in nmt if l[:2] == ("norm", "scale")])

    # ---- cfg2: UVd rank 10 on the RNN-XOR model's 1021 parameters ----------------------------------------------------
    n, r = 1021, 10
    U = torch.randn(n, r, device=dev, generator=g) * (1.0 / (n * r)) ** 0.5
    V = torch.randn(n, r, device=dev, generator=g) * (1.0 / (n * r)) ** 0.5
    d = torch.ones(n, 1, device=dev)
    ins = []
    for _ in range(2):
        v = torch.randn(n, 1, device=dev, generator=g)
        ins.append((v, 1.3 * v + 0.1 * torch.randn(n, 1, device=dev, generator=g), torch.randn(n, 1, device=dev, generator=g)))
    k = [0]

    def eager_uvd():
        v, h, gg = ins[k[0] & 1]; k[0] += 1
        return psgd.update_precond_and_grad_UVd(U, V, d, v, h, gg, 0.01, psgd._tiny, balance=False, update_U=bool(k[0] & 2))

    l0 = ctx.launch_count
    eager_uvd()
    per_step_launches = ctx.launch_count - l0
    ms_e = _time(torch, eager_uvd, steps)
    gu = UVdStepGraphs(U, V, d, 0.01, psgd._tiny, fused=True)

    def graphed_uvd():
        v, h, gg = ins[k[0] & 1]; k[0] += 1
        return gu.step(v, h, gg, False, bool(k[0] & 2))

    for _ in range(10):
        graphed_uvd()
    ms_g = _time(torch, graphed_uvd, steps)
    assert torch.isfinite(graphed_uvd()).all() and torch.isfinite(U).all()
    rows.append(dict(path="cfg2 UVd rank 10, N = 1021 (fused update+apply call)", cite="rnn_xor_UVd_preconditioner.py:37-41",
                     launches_per_step=int(per_step_launches), eager_us_per_step=round(ms_e * 1e3, 1),
                     graph_us_per_step=round(ms_g * 1e3, 1), steps_per_s=round(1e3 / ms_g, 1)))
    return rows


def run_mixed_pairs(peak_gbs: float, steps: int = 10):
    """(norm, dense) and (dense, scale): one dense factor on the tensor cores + the streaming pieces of the structured
    one.  Dense flop count of the reference's op sequence: with the dense side n and the long side m,
    update 2mn^2 (A) + mn^2 (solve) + 4mn^2 (two Gram products) + 2n^3, apply 4mn^2 -> 11 m n^2 + 2 n^3 per step."""
    import torch
    import psgd_tf_b200 as psgd

    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(13)
    rows = []
    for kl, kr, M, N in (("norm", "dense", 16384, 1024), ("dense", "scale", 1024, 16384)):
        state = [_factor(torch, kl, M, dev), _factor(torch, kr, N, dev)]
        dX = torch.randn(M, N, device=dev, generator=g)
        dG = 1.3 * dX + 0.1 * torch.randn(M, N, device=dev, generator=g)
        G = torch.randn(M, N, device=dev, generator=g)

        def upd():
            state[0], state[1] = psgd.update_precond_kron(state[0], state[1], dX, dG, 0.01)

        ms_u = _time(torch, upd, steps)
        ms_a = _time(torch, lambda: psgd.precond_grad_kron(state[0], state[1], G), steps)
        assert torch.isfinite(state[0]).all() and torch.isfinite(state[1]).all()
        n, m = (N, M) if kr == "dense" else (M, N)
        fl_u, fl_a = 7.0 * m * n * n + 2.0 * n ** 3, 4.0 * m * n * n
        byts = _kron_step_bytes(kl, kr, M, N)
        rows.append(dict(path=f"kron ({kl},{kr})", shape=[M, N], update_ms=round(ms_u, 4), apply_ms=round(ms_a, 4),
                         steps_per_s=round(1e3 / (ms_u + ms_a), 2),
                         update_TFLOPs=round(fl_u / ms_u / 1e9, 1), apply_TFLOPs=round(fl_a / ms_a / 1e9, 1),
                         step_dense_GFLOP=round((fl_u + fl_a) / 1e9, 1), step_TFLOPs=round((fl_u + fl_a) / (ms_u + ms_a) / 1e9, 1),
                         step_compulsory_GB=round(byts / 1e9, 3), step_GBps=round(byts / (ms_u + ms_a) / 1e6, 1),
                         step_frac_hbm=round(byts / (ms_u + ms_a) / 1e6 / peak_gbs, 4)))
    return rows


def run_aux_sharded(peak_gbs: float, rank: int, world: int, steps: int = 10, n_total: int = 100_000_000):
    """Diagonal and X-shape at N > 1 GPUs: the vector is sharded by contiguous chunk (diagonal) or by
    mirrored chunk pairs (X-shape, partition.xmat_shard_slices); the only exchange is the max of |nabla| per update,
    through whatever exchange bench.py installed on the context (peer-memory kernel or the all-reduce hook).  Timed on
    the device between barriers, max over ranks; GB/s is the whole job's."""
    import torch
    import torch.distributed as dist
    import psgd_tf_b200 as psgd
    from psgd_tf_b200 import partition

    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(70 + rank)
    rows = []

    def timed(fn):
        for _ in range(3):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for name, bu, ba in (("diagonal", 28.0, 12.0), ("X-shape", 40.0, 16.0)):
        if name == "diagonal":
            lo, hi = partition.chunk_bounds(n_total, world, rank)
            n = hi - lo
        else:
            n = sum(b - a for a, b in partition.xmat_shard_slices(n_total, world, rank))
        v = torch.randn(n, device=dev, generator=g)
        h = (0.5 + 1.5 * torch.rand(n, device=dev, generator=g)) * v + 0.1 * torch.randn(n, device=dev, generator=g)
        gr = torch.randn(n, device=dev, generator=g)
        a = torch.ones(n, device=dev)
        b = torch.zeros(n, device=dev)
        if name == "diagonal":
            ms_u = timed(lambda: psgd.update_precond_diag(a, v, h, 0.01))
            ms_a = timed(lambda: psgd.precond_grad_diag(a, gr))
        else:
            ms_u = timed(lambda: psgd.update_precond_Xmat(a, b, v, h, 0.01))
            ms_a = timed(lambda: psgd.precond_grad_Xmat(a, b, gr))
        assert torch.isfinite(a).all()
        tot = (bu + ba) * n_total
        rows.append(dict(path=f"{name} sharded x{world}", shape=[n_total], rows_per_gpu=int(n), update_ms=round(ms_u, 4),
                         apply_ms=round(ms_a, 4), steps_per_s=round(1e3 / (ms_u + ms_a), 2),
                         step_algorithmic_GB=round(tot / 1e9, 3), step_GBps=round(tot / (ms_u + ms_a) / 1e6, 1),
                         step_frac=round(tot / (ms_u + ms_a) / 1e6 / (peak_gbs * world), 4)))
        del v, h, gr, a, b
        torch.cuda.empty_cache()
    return rows


if __name__ == "__main__":
    import torch
    from bench import load_peaks
    assert torch.cuda.is_available(), "bench_aux needs a CUDA device"
    for r in run_aux(load_peaks()["hbm"]):
        print(json.dumps(r))
