"""Kron-stack workload of bench.py: synthetic L-layer n x n MLP gradient stack, batched dense-dense Kron update+apply
(BASELINE.json configs[2]); layers sharded layer-wise over the ranks, preconditioned gradients all-gathered."""
from __future__ import annotations

import os
import time

import numpy as np


def tf32_peak_tflops(torch, n=8192, iters=10, sustain_s=2.0):
    """cuBLAS TF32 GEMM throughput on this GPU: calibration of the 3xTF32 roofline denominator only (MEASURED_PEAKS.json
    records bf16 but no TF32 figure); never on the product path.  Returns (burst, sustained): best of three 10-launch
    bursts, and the rate over the second half of `sustain_s` seconds of back-to-back launches -- dense tensor-core work
    pulls the board to its power cap within a fraction of a second and the SM clock settles ~15 % lower, which is the
    regime the GEMMs of a 150 ms Kron step run in (B200_PROFILING.md: burst peak for a kernel timed alone, sustained
    for a kernel timed inside a long step)."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
    for _ in range(3):
        a @ b
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(iters):
            a @ b
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    n_half = max(10, int(sustain_s * 0.5 / (best * 1e-3)))
    for _ in range(n_half):                  # first half: let the clocks settle under the power cap
        a @ b
    e0.record()
    for _ in range(n_half):
        a @ b
    e1.record(); torch.cuda.synchronize()
    sustained = e0.elapsed_time(e1) / n_half
    torch.backends.cuda.matmul.allow_tf32 = prev
    del a, b
    return 2.0 * n ** 3 / best / 1e9, 2.0 * n ** 3 / sustained / 1e9


def kron_flops(n):
    """Dense flop count of the reference's op sequence for one square dense-dense layer:
    update 18 n^3 + apply 8 n^3."""
    return 26.0 * n ** 3


def run_kron(args, rank, world, local, METRIC, UNIT, load_peaks, ClockSampler):
    import torch
    import torch.distributed as dist
    import psgd_tf_b200 as psgd
    from psgd_tf_b200 import partition
    from bench import kron_config

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L, n = args.layers, args.kron_n
    owned = partition.assign_layers([partition.kron_layer_cost(n, n)] * L, world)
    mine = owned[rank]
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    ctx = psgd.get_context(local)

    Ql = [torch.eye(n, device=dev) for _ in mine]                    # identity init (mnist_with_lenet5.py:61-62)
    Qr = [torch.eye(n, device=dev) for _ in mine]
    S = [0.5 + 1.5 * torch.rand(n, 1, device=dev, generator=gen) for _ in mine]
    T = [0.5 + 1.5 * torch.rand(1, n, device=dev, generator=gen) for _ in mine]
    POOL = 4          # cycle through >= 4 pre-generated input sets
    pool = []
    for _ in range(POOL):
        dX = [torch.randn(n, n, device=dev, generator=gen) for _ in mine]
        dG = [s * x * t + 0.1 * torch.randn(n, n, device=dev, generator=gen) for s, x, t in zip(S, dX, T)]
        G = [torch.randn(n, n, device=dev, generator=gen) for _ in mine]
        pool.append((dX, dG, G))
    shapes = [(n, n)] * L

    gmode = getattr(args, "kron_gather", "auto")
    if gmode == "auto":
        # measured: gathering slot by slot under the next layer's apply pays once the all-gather is a
        # large share of the step (8 GPUs, 3 layers each: +1.5 %); with more layers per GPU the grouped apply wins
        gmode = "slots" if 0 < len(mine) <= 3 else "once"
    gbuf = None
    if world > 1:
        gbuf = (partition.KronPeerGather if gmode.startswith("peer") else partition.KronGatherBuffer)(shapes, owned, rank, dev)

    def step(i, Ql, Qr, dX, dG, G):
        new = psgd.update_precond_kron_batched(Ql, Qr, dX, dG, 0.01)
        Ql2, Qr2 = [a for a, _ in new], [b for _, b in new]
        if gbuf is None:
            return Ql2, Qr2, psgd.precond_grad_kron_batched(Ql2, Qr2, G)
        # the apply writes straight into this rank's entries of the gather buffer (no torch.stack, no staging copy)
        outs = gbuf.local_outs()
        if gmode == "once":          # whole batched apply, then ONE in-place NCCL all-gather
            psgd.precond_grad_kron_batched(Ql2, Qr2, G, outs=outs)
            return Ql2, Qr2, gbuf.gather()
        if gmode == "peer-once":     # whole batched apply, then copy-engine pushes to every peer
            psgd.precond_grad_kron_batched(Ql2, Qr2, G, outs=outs)
            return Ql2, Qr2, gbuf.gather()
        # layer by layer, each followed by the transfer of its slot, which runs under the apply of the next layer:
        # "slots" = NCCL all-gather per slot, "peer" = copy-engine pushes per slot
        for j in range(len(outs)):
            psgd.precond_grad_kron_batched(Ql2[j:j + 1], Qr2[j:j + 1], G[j:j + 1], outs=outs[j:j + 1])
            gbuf.push_slot(j) if gmode == "peer" else gbuf.gather_slot(j)
        return Ql2, Qr2, gbuf.finish()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()          # before the warm-up: NVML start-up must not land inside the timed region
    for i in range(args.warmup):
        Ql, Qr, pre = step(i, Ql, Qr, *pool[i % POOL])
    ctx.set_option("profile", 2)          # per-launch work = EXECUTED flops (triangular clipping / tile skipping applied)
    ctx.profile_read(cap=1 << 20)
    barrier()
    clocks.mark()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        Ql, Qr, pre = step(args.warmup + i, Ql, Qr, *pool[i % POOL])
    e1.record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    prof = ctx.profile_read(cap=1 << 20)
    ctx.set_option("profile", 0)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    assert all(torch.isfinite(p).all() for p in pre[:1])
    value = args.steps / (ms / 1e3)
    gather_check = None
    if world > 1:
        # every rank holds every layer's result after the gather: per-layer float64 checksums of what each OWNER
        # computed against the checksums of what arrived here (bit-exact transfer => equal sums)
        mine_sums = torch.stack([pre[li].double().sum() for li in mine])
        all_sums = [torch.empty_like(mine_sums) for _ in range(world)]
        dist.all_gather(all_sums, mine_sums)
        want = torch.empty(L, dtype=torch.float64, device=dev)
        for k, layer_ids in enumerate(owned):
            for j, li in enumerate(layer_ids):
                want[li] = all_sums[k][j]
        got = torch.stack([p.double().sum() for p in pre])
        gather_check = bool(torch.equal(got, want))
        assert gather_check, "all-gathered preconditioned gradients differ from their owners' results"

    # ---- roofline: tensor pipe, 3xTF32 => ceiling = measured TF32 GEMM peak / 3 ---------------------
    tf32, tf32_sus = tf32_peak_tflops(torch)
    ceiling = tf32_sus / 3.0           # the GEMMs are timed inside a long step: sustained (power-capped) figure
    names = {10: "gemm_tc_kernel (tcgen05 3xTF32)", 11: "trsm_block_kernel (SIMT diagonal blocks)", 12: "gemm_simt_kernel"}
    agg = {}
    for kid, kms, work in prof:
        a = agg.setdefault(kid, [0.0, 0, 0.0])
        a[0] += kms; a[1] += 1; a[2] += work
    kernels = []
    for kid, (tot, cnt, work) in sorted(agg.items()):
        ach = work / (tot * 1e-3) / 1e12 if tot > 0 else 0.0
        kernels.append(dict(kernel=names.get(kid, str(kid)), launches=cnt, total_ms=round(tot, 3), avg_ms=round(tot / cnt, 4),
                            executed_TFLOP=round(work / 1e12, 3), achieved_TFLOPs=round(ach, 1),
                            frac=round(ach / ceiling, 4), share_of_step=round(tot / ms, 4)))
    # per-launch view of ONE step (the launch sequence repeats every step): which of the step's grouped launches run below
    # the plain-product rate.  Times are averaged over the timed steps, position by position.
    per_step = len(prof) // max(args.steps, 1)
    by_launch = []
    if per_step and per_step * args.steps == len(prof):
        for pos in range(per_step):
            rows = [prof[st * per_step + pos] for st in range(args.steps)]
            kid = rows[0][0]
            ms_avg = sum(r[1] for r in rows) / len(rows)
            work = rows[0][2]
            by_launch.append(dict(pos=pos, kernel={10: "gemm_tc", 11: "tri_inv", 12: "gemm_simt"}.get(kid, str(kid)),
                                  ms=round(ms_avg, 3), executed_TFLOP=round(work / 1e12, 3),
                                  TFLOPs=round(work / (ms_avg * 1e-3) / 1e12, 1) if ms_avg > 0 else 0.0))
    dom = max(kernels, key=lambda k: k["total_ms"]) if kernels else None
    step_flops = kron_flops(n) * len(mine)
    step_ach = step_flops / (ms / args.steps * 1e-3) / 1e12
    roofline = None
    if dom:
        roofline = dict(bound="tensor", kernel=dom["kernel"], achieved=dom["achieved_TFLOPs"], peak=round(ceiling, 1),
                        unit="TFLOP/s", frac=dom["frac"], traffic=None,
                        peak_source=f"cuBLAS TF32 8192^3 measured in this run, SUSTAINED over 2 s of back-to-back launches = "
                                    f"{tf32_sus:.0f} TFLOP/s (10-launch burst: {tf32:.0f}), divided by 3 (3xTF32); the kernel is "
                                    f"timed inside a ~150 ms power-capped step.  MEASURED_PEAKS.json has no TF32 figure (bf16 "
                                    f"burst {load_peaks()['bf16']:.0f}, sustained {load_peaks()['bf16_sustained']:.0f})",
                        peak_burst=round(tf32 / 3.0, 1), frac_of_burst=round(dom["achieved_TFLOPs"] / (tf32 / 3.0), 4),
                        step_achieved=round(step_ach, 1), step_frac=round(step_ach / ceiling, 4),
                        step_algorithmic_TFLOP=round(step_flops / 1e12, 2),
                        note="achieved/frac: the dominant kernel's EXECUTED fp32-equivalent flops (after triangular K "
                             "clipping and tile skipping) over its launch time; step_*: the dense count of the "
                             "reference's op sequence (26 n^3 per layer-step) over the step time -- skipping "
                             "structurally-zero tiles legitimately raises that fraction above the kernel's")

    # ---- parity at the benchmarked size, outside every timed region: layer 0's next step on the GPU and through the
    # CPU twin of the oracle (psgd.py:156-192) on the same inputs, from the factors the timed steps left behind
    parity = None
    if rank == 0 and world == 1 and not getattr(args, "no_parity", False) and mine:
        parity = kron_parity(psgd, Ql[0], Qr[0], pool[0][0][0], pool[0][1][0], pool[0][2][0])
        if roofline is not None:
            roofline["parity_rel_err"] = parity["max_rel_err"]

    # ---- end to end: host (pinned) dX, dG, G per step, preconditioned gradients read back ----------------
    # Same public API (the batched calls); the step's inputs are uploaded from pinned host memory and its results read
    # back to pinned host memory INSIDE the timed region, double-buffered on two copy streams so that the upload of
    # step i+1 and the read-back of step i-1 overlap the tensor-core work of step i.
    e2e = None
    if not args.no_e2e:
        del pool
        torch.cuda.empty_cache()
        NB = 2
        nl = len(mine)
        h_in = [[torch.randn(n, n).pin_memory() for _ in range(nl)] for _ in range(3)]       # dX, dG, G (one host copy)
        h_in[1] = [(1.3 * x + 0.1 * torch.randn(n, n)).pin_memory() for x in h_in[0]]
        h_out = [[torch.empty(n, n).pin_memory() for _ in range(nl)] for _ in range(NB)]
        d_in = [[[torch.empty(n, n, device=dev) for _ in range(nl)] for _ in range(3)] for _ in range(NB)]
        s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
        cur = torch.cuda.current_stream()

        def e2e_loop(steps, Ql, Qr):
            in_ready, buf_free, out_done, keep = [None] * NB, [None] * NB, [None] * NB, [None] * NB

            def upload(i):
                k = i % NB
                with torch.cuda.stream(s_h2d):
                    if buf_free[k] is not None:
                        s_h2d.wait_event(buf_free[k])
                    for dst, src in zip(d_in[k], h_in):
                        for a, b in zip(dst, src):
                            a.copy_(b, non_blocking=True)
                    ev = torch.cuda.Event(); ev.record(s_h2d); in_ready[k] = ev

            upload(0)
            for i in range(steps):
                k = i % NB
                if i + 1 < steps:
                    upload(i + 1)
                cur.wait_event(in_ready[k])
                new = psgd.update_precond_kron_batched(Ql, Qr, d_in[k][0], d_in[k][1], 0.01)
                Ql, Qr = [a for a, _ in new], [b for _, b in new]
                pre = psgd.precond_grad_kron_batched(Ql, Qr, d_in[k][2])
                if world > 1:
                    pre = [pre[j] for j in range(nl)]      # each rank reads back the layers it owns
                ev = torch.cuda.Event(); ev.record(cur); buf_free[k] = ev
                with torch.cuda.stream(s_d2h):
                    s_d2h.wait_event(ev)
                    if out_done[k] is not None:
                        s_d2h.wait_event(out_done[k])
                    for a, b in zip(h_out[k], pre):
                        a.copy_(b, non_blocking=True)
                        b.record_stream(s_d2h)
                    ev2 = torch.cuda.Event(); ev2.record(s_d2h); out_done[k] = ev2
                keep[k] = pre                                # alive until its read-back has been enqueued and replaced
            cur.wait_stream(s_d2h)
            return Ql, Qr

        Ql, Qr = e2e_loop(2, Ql, Qr)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ksteps = max(2, min(args.steps, 20))    # the first upload and the last read-back are not overlapped: amortise them
        f0.record()
        Ql, Qr = e2e_loop(ksteps, Ql, Qr)
        f1.record()
        barrier()
        ems = f0.elapsed_time(f1)
        t = torch.tensor([ems], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ems = float(t.item())
        assert np.isfinite(h_out[0][0].numpy()).all()
        e2e = dict(value=round(ksteps / (ems / 1e3), 4), unit=UNIT, h2d_bytes_per_step=int(3 * 4 * n * n * L),
                   d2h_bytes_per_step=int(4 * n * n * L), ms_per_step=round(ems / ksteps, 2), steps=ksteps,
                   note="dX, dG, G of every layer uploaded from pinned host memory and every preconditioned gradient read "
                        "back to pinned host memory each step, through the batched public API; factors stay "
                        "device-resident; copies double-buffered on two side streams")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_kron(L, n)
    if rank != 0:
        return None
    return dict(
        metric=METRIC, value=round(value, 4), unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=round(ms / args.steps, 3), higher_is_better=True, scaling="strong", vs_baseline=None,
        dtype="f32 (3xTF32 tensor-core products, fp32 accumulate)", data="synthetic",
        config=kron_config(L, n, world),
        run=dict(layers_per_gpu=len(mine), gather=gmode if world > 1 else None,
                 gather_checksums_match=gather_check), parity=parity,
        roofline=roofline, kernels=kernels, launches_of_one_step=by_launch, cpu_baseline=cpu, e2e=e2e,
        gpu_launches=int(launches), clocks=clk)


def kron_parity(psgd, Ql, Qr, dX, dG, G):
    import torch
    from oracle import psgd_oracle_torch as T
    from bench import rel_err_chunked
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    t0 = time.perf_counter()
    ql, qr = psgd.update_precond_kron(Ql, Qr, dX, dG, 0.01)
    pre = psgd.precond_grad_kron(ql, qr, G)
    got = dict(Ql=ql.cpu(), Qr=qr.cpu(), pre_grad=pre.cpu())
    qlr, qrr = T.update_precond_dense_dense(Ql.cpu(), Qr.cpu(), dX.cpu(), dG.cpu(), 0.01)
    want = dict(Ql=qlr, Qr=qrr, pre_grad=T.precond_grad_dense_dense(qlr, qrr, G.cpu()))
    errs = {k: rel_err_chunked(got[k], want[k], 4) for k in got}
    return dict(max_rel_err=float("%.3e" % max(errs.values())), rel_err={k: float("%.3e" % e) for k, e in errs.items()},
                tolerance=1e-5, passed=bool(max(errs.values()) <= 1e-5), layer_shape=list(dX.shape),
                against="oracle/psgd_oracle_torch.py (multi-threaded CPU twin of the NumPy oracle; psgd.py:156-192) on the "
                        "same inputs: layer 0's next update + apply from the factors the timed steps left",
                seconds=round(time.perf_counter() - t0, 1))


def cpu_baseline_kron(L, n, steps=2):
    """The CPU port (multi-threaded torch-CPU restatement of psgd.py:156-192) on ONE full-size layer per step, times
    the layer count (bounded to ~10 s; `bench.py --impl reference` runs the whole stack)."""
    import argparse
    from bench import run_reference_kron
    r = run_reference_kron(argparse.Namespace(kron_n=n, layers=L, steps=steps, warmup=1, scaling="strong"), full=False)
    return r["cpu_baseline"]
