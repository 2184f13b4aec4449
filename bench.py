#!/usr/bin/env python
"""bench.py -- precond update+apply steps/s for the PSGD hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload uvd|kron] [--impl ours|reference]

One "step" = one preconditioner update followed by one preconditioned-gradient apply on fresh synthetic
(v, h, g) / (dX, dG, G), state carried across steps.

Workloads (default "all": the UVd line is the headline JSON, the Kron stack result is nested under its "kron" key)
  uvd             UVd rank-10 on a flattened 100M-parameter vector (BASELINE.json configs[3]); at N GPUs the vector
                  is sharded by contiguous chunk (strong scaling) with all-reduces of the r x r / r-length partials.
  kron            synthetic 24-layer 4096x4096 stack, dense-dense Kron update+apply (configs[2]), layers sharded.

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` is the same metric through
the public API with HOST (pinned) inputs and a host read-back of the result inside the timed region; `roofline`
describes the dominant kernel (CUDA-event timed per launch inside the timed region); `cpu_baseline` is the CPU oracle
timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "precond update+apply steps/s"
UNIT = "steps/s"
KERNEL_NAMES = {1: "uvd_gram_update", 2: "uvd_map_update2", 3: "uvd_map_update3", 4: "uvd_gram_apply", 5: "uvd_map_apply",
                6: "peer_exchange", 7: "uvd_map_fused", 8: "uvd_d_update", 9: "uvd_map_updapp", 13: "uvd_map_apply_d",
                10: "gemm", 11: "trsm"}


# ---------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return dict(hbm=float(d["hbm_gbs"]), bf16=float(d.get("bf16_tflops", 1590.0)),
                        bf16_sustained=float(d.get("bf16_tflops_sustained", 1400.0)), source="measured")
        except Exception:
            pass
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe).

    ``start()`` launches ``nvidia-smi -lms 50`` and is called BEFORE the warm-up steps: NVML start-up takes 0.1-0.3 s on
    a fresh box and stalls kernel submission while it initialises, which must not land inside a 0.1 s timed region.
    ``mark()`` stamps the start of the timed region, ``stop()`` its end; only samples stamped in between are used (the
    nearest ones if the region is shorter than the sampling period)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.proc, self.idx, self.t0 = None, device_index, None
        self.path = f"/tmp/psgd_bench_clocks_{os.getpid()}_{id(self)}.csv"

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=self.fh, stderr=subprocess.DEVNULL)
            t_end = time.time() + 1.5
            while time.time() < t_end and os.path.getsize(self.path) == 0:      # wait until NVML is up and sampling
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def mark(self):
        import datetime
        self.t0 = datetime.datetime.now()

    def stop(self):
        import datetime
        t1 = datetime.datetime.now()
        if self.proc is None:
            return None
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        rows = []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                rows.append((ts, float(f[1]), float(f[2]), [v.lower().startswith("active") for v in f[5:9]]))
            except ValueError:
                continue
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not rows:
            return None
        t0 = self.t0 or rows[0][0]
        inside = [r for r in rows if t0 <= r[0] <= t1]
        if not inside:        # region shorter than the sampling period: the sample nearest to its middle
            mid = t0 + (t1 - t0) / 2
            inside = [min(rows, key=lambda r: abs((r[0] - mid).total_seconds()))]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({n for r in inside for n, on in zip(names, r[3]) if on})
        return dict(sm_mhz=float(np.median([r[1] for r in inside])), sm_max_mhz=float(max(r[2] for r in inside)),
                    reasons=reasons, samples=len(inside))


def load_traffic():
    """Per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernels, from the committed
    ncu --set full capture (profiles/*_traffic.json, written by tools/summarise_profiles.py).  Only valid for the
    default problem sizes the capture was taken at."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    if os.path.isdir(pdir):
        for f in sorted(os.listdir(pdir)):
            if f.endswith("_traffic.json"):
                best = os.path.join(pdir, f)
    if not best:
        return {}, None
    try:
        d = json.load(open(best))
        return d.get("bytes_per_launch", {}), os.path.relpath(best, ROOT)
    except Exception:
        return {}, None


def bind_to_gpu_numa(local):
    """Pin this rank's host threads (and, by first touch, the pinned staging buffers it allocates afterwards) to the NUMA
    node its GPU hangs off.  torchrun starts every rank unbound; round 1's end-to-end numbers at 8 GPUs were host-copy
    bound with all ranks' staging memory on node 0.  Best effort: returns a short description or None."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return "unbound: nvidia-smi gave no PCI bus id"
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:]}:{rest}/numa_node"
        if not os.path.exists(path):
            return f"unbound: {path} does not exist"
        node = int(open(path).read().strip())
        if node < 0:
            return "unbound: the platform reports numa_node = -1 for the GPU (no NUMA information in this container)"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return f"unbound: no allowed cpu on numa node {node}"
        os.sched_setaffinity(0, allowed)
        return f"numa node {node} ({len(allowed)} cpus)"
    except Exception as e:
        return f"unbound: {type(e).__name__}: {e}"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------------------------
# UVd workload
# ---------------------------------------------------------------------------------------------
def uvd_bytes(n, r, form="fused"):
    """Algorithmic bytes per launch = what the kernel has to read + the RESULTS it has to write (intermediates such as
    the stored nablaD are not counted; ncu's `traffic` shows them).
    SURVEY.md section 8d's step figure 4n(9r+12) models update and apply as separate calls (big inputs read twice
    each, outputs written once) and is what `step_frac_survey` is quoted against; the fused update+apply call needs
    less -- 4n(7r+11): sweep 1 reads U V d h v, sweep 2 reads U V d h v g and writes U (or V), sweep 3 reads U V d g
    and writes d and the preconditioned gradient."""
    per_kernel = {1: 4 * n * (2 * r + 3), 2: 4 * n * (2 * r + 3), 3: 4 * n * (r + 1), 4: 4 * n * (2 * r + 2),
                  5: 4 * n * (2 * r + 2) + 4 * n,
                  7: 4 * n * (2 * r + 3) + 4 * n * r, 8: 4 * n,
                  9: 4 * n * (2 * r + 4) + 4 * n * r, 13: 4 * n * (2 * r + 2) + 8 * n}
    step = 4 * n * (7 * r + 11) if form == "fused" else 4 * n * (9 * r + 12)
    return per_kernel, step, 4 * n * (9 * r + 12)


def uvd_config(n_total, r, world):
    """The `config` object of the UVd workload -- built by ONE function for both arms (`--impl ours` and `--impl
    reference` must name the same workload); everything specific to how an arm executes it goes under `run`."""
    rows = chunk_of(n_total, world, 0)[1]
    return dict(workload=f"UVd rank-{r} update+apply on a flattened {n_total:,}-parameter vector (BASELINE configs[3])",
                n_params=n_total, rank=r, parallelism=f"chunk-sharded x{world}",
                l2_policy="inputs larger than L2: >= %.1f GB of state+inputs streamed per GPU per step vs 126 MB L2" %
                          ((4 * rows * (2 * r + 4)) / 1e9),
                coin_flips="update_U alternates, balance every 100th step", step_size=0.01)


def kron_config(L, n, world):
    per = -(-L // world)
    return dict(workload=f"{L}-layer {n}x{n} dense-dense Kron update+apply, batched (BASELINE configs[2])", layers=L, n=n,
                parallelism=f"layer-sharded x{world} + all-gather of preconditioned gradients",
                l2_policy=f"inputs larger than L2: {per * 5 * 4 * n * n / 1e9:.1f} GB of factors+inputs per GPU per step vs 126 MB L2",
                step_size=0.01)


def rel_err_chunked(a, b, chunks=16):
    """||a - b||_F / ||b||_F of two (large) CPU tensors with float64 accumulation and bounded temporaries."""
    import torch
    a, b = a.reshape(-1), b.reshape(-1)
    num = den = 0.0
    step = -(-a.numel() // chunks)
    for i in range(0, a.numel(), max(step, 1)):
        x, y = a[i:i + step].double(), b[i:i + step].double()
        num += float(torch.sum((x - y) ** 2)); den += float(torch.sum(y ** 2))
    return (num / max(den, 1e-300)) ** 0.5


def chunk_of(n, world, rank, align=256):
    per = -(-n // world)
    per = -(-per // align) * align
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


def run_uvd(args, rank, world, local):
    import torch
    import torch.distributed as dist
    import psgd_tf_b200 as psgd
    from psgd_tf_b200 import partition

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    N, r = args.n, args.rank
    if args.scaling == "weak":
        lo, hi = 0, N
        N_total = N * world
    else:
        lo, hi = chunk_of(N, world, rank)
        N_total = N
    n = hi - lo
    gen = torch.Generator(device=dev).manual_seed(2024 + rank)
    uv = (1.0 / (N_total * r)) ** 0.5                                   # psgd.py:687
    U0 = torch.randn(n, r, device=dev, generator=gen) * uv
    V0 = torch.randn(n, r, device=dev, generator=gen) * uv
    d0 = torch.ones(n, 1, device=dev)
    POOL = 4
    pool = []
    for _ in range(POOL):
        v = torch.randn(n, 1, device=dev, generator=gen)
        c = 0.5 + 1.5 * torch.rand(n, 1, device=dev, generator=gen)
        h = c * v + 0.1 * torch.randn(n, 1, device=dev, generator=gen)
        g = torch.randn(n, 1, device=dev, generator=gen)
        pool.append((v, h, g))
        del c
    ctx = psgd.get_context(local)
    exchange = "none (single GPU)"
    if world > 1:
        exchange = partition.install_exchange(ctx) if not args.nccl_hook else (partition.install_allreduce(ctx) or "all-reduce hook")

    # CUDA graphs need a host-free chain: fine on one GPU and with the peer-memory exchange, not with the Python hook.
    # The timed region replays graphs at EVERY N (one measurement mode for the whole scaling curve); the per-kernel
    # roofline comes from an eager pass right after it, in which every launch is bracketed by CUDA events.
    use_graphs = (not args.no_graphs) and (world == 1 or exchange == "peer-memory")
    graphs = {}

    form = {"fused": "fused"}.get(args.uvd_form, "separate")

    def step(i, U, V, d, v, h, g):
        balance, update_U = (i % 100 == 99), (i % 2 == 0)
        if use_graphs:
            gs = graphs.get((U.data_ptr(), form))
            if gs is None:
                from psgd_tf_b200.graphs import UVdStepGraphs
                gs = graphs[(U.data_ptr(), form)] = UVdStepGraphs(U, V, d, 0.01, psgd._tiny, fused=(form == "fused"))
            return gs.step(v, h, g, balance, update_U)
        if form == "fused":       # the sequence UVd.step runs (psgd.py:732-748) as one call: three sweeps over U, V
            return psgd.update_precond_and_grad_UVd(U, V, d, v, h, g, 0.01, psgd._tiny, balance=balance, update_U=update_U)
        psgd.update_precond_UVd_math_(U, V, d, v, h, 0.01, psgd._tiny, balance=balance, update_U=update_U)
        return psgd.precond_grad_UVd_math(U, V, d, g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -----------------------------------------------------------------
    ctx.set_option("uvd_fused", 0 if args.uvd_form == "separate-3sweep" else 1)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    U, V, d = U0.clone(), V0.clone(), d0.clone()
    # Step k always uses input set k % POOL and coin flips (k % 100 == 99, k % 2 == 0), in the warm-up and in the timed
    # region alike, so that every CUDA graph the timed region replays (keyed by input buffers + coin flips) has been
    # captured during the warm-up.
    run_step = lambda k: step(k, U, V, d, *pool[k % POOL])
    n_warm = args.warmup + (2 * POOL if use_graphs else 0)
    for i in range(n_warm):
        run_step(i)
    if not use_graphs:
        ctx.set_option("profile", 1)
    ctx.profile_read()
    def timed_region(first_step):
        """EXACTLY args.steps steps between two events (barrier + synchronize on both sides); an event after every
        step as well, so that a one-off stall of the submitting host thread (another process holding the driver) shows
        up as one long step instead of silently inflating the mean."""
        barrier()
        clocks.mark()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        marks[0].record()
        out = None
        for i in range(args.steps):
            out = run_step(first_step + i)
            marks[i + 1].record()
        barrier()
        per = [marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)]
        return marks[0].elapsed_time(marks[-1]), per, out

    launches0 = ctx.launch_count
    graph_launches0 = sum(g.kernel_launches for g in graphs.values())
    ms, per_step, pre = timed_region(n_warm)
    remeasured = None
    med = float(np.median(per_step))
    flag = torch.tensor([1.0 if max(per_step) > 3.0 * med and ms > 1.15 * med * args.steps else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if flag.item() > 0:
        # a submission stall landed in the timed region: re-measure once (same rule as for a throttled run) and say so
        remeasured = dict(reason="one step took > 3x the median step (host/driver stall, not kernel time)",
                          discarded_ms_per_step=round(ms / args.steps, 4), discarded_max_step_ms=round(max(per_step), 3))
        ctx.profile_read()
        launches0 = ctx.launch_count
        graph_launches0 = sum(g.kernel_launches for g in graphs.values())
        ms, per_step, pre = timed_region(n_warm + args.steps)
    clk = clocks.stop() if rank == 0 else None
    launches = ctx.launch_count - launches0 + sum(g.kernel_launches for g in graphs.values()) - graph_launches0
    kernels_from = "CUDA events around every launch inside the timed region"
    if use_graphs:
        # the timed region replayed CUDA graphs (no per-launch events possible): per-kernel times come from a separate
        # eager pass over the same inputs, right after it
        use_graphs = False
        ctx.set_option("profile", 1)
        ctx.profile_read()
        for i in range(min(args.steps, 6)):
            run_step(n_warm + i)
        torch.cuda.synchronize()
        use_graphs = True
        kernels_from = "separate eager pass right after the (graph-replayed) timed region"
    prof = ctx.profile_read(cap=65536)
    ctx.set_option("profile", 0)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    assert torch.isfinite(pre).all(), "non-finite preconditioned gradient"
    value = args.steps / (ms / 1e3)

    # ---- per-kernel roofline ----------------------------------------------------------------------
    peaks = load_peaks()
    per_kernel_bytes, step_bytes, survey_bytes = uvd_bytes(n, r, form)
    agg = {}
    for kid, kms, _w in prof:
        a = agg.setdefault(kid, [0.0, 0])
        a[0] += kms; a[1] += 1
    kernels = []
    for kid, (tot, cnt) in sorted(agg.items()):
        avg = tot / cnt
        ach = per_kernel_bytes.get(kid, 0) / (avg * 1e-3) / 1e9
        kernels.append(dict(kernel=KERNEL_NAMES.get(kid, str(kid)), launches=cnt, avg_ms=round(avg, 4),
                            algorithmic_GB=round(per_kernel_bytes.get(kid, 0) / 1e9, 3), achieved_GBps=round(ach, 1),
                            frac=round(ach / peaks["hbm"], 4), share_of_step=round(tot / ms, 4)))
    dom = max(kernels, key=lambda k: k["avg_ms"] * k["launches"]) if kernels else None
    roofline = None
    if dom:
        traffic_tab, traffic_src = load_traffic()
        traffic = traffic_tab.get(dom["kernel"]) if (N == 100_000_000 and r == 10 and world == 1) else None
        roofline = dict(bound="hbm", kernel=dom["kernel"], achieved=dom["achieved_GBps"], peak=peaks["hbm"], unit="GB/s",
                        frac=dom["frac"], traffic=traffic, traffic_source=traffic_src if traffic else None,
                        peak_source=f"of {peaks['source']}",
                        step_achieved=round(step_bytes / (ms / args.steps * 1e-3) / 1e9, 1),
                        step_frac=round(step_bytes / (ms / args.steps * 1e-3) / 1e9 / peaks["hbm"], 4),
                        step_algorithmic_GB=round(step_bytes / 1e9, 3),
                        step_frac_survey=round(survey_bytes / (ms / args.steps * 1e-3) / 1e9 / peaks["hbm"], 4),
                        step_survey_GB=round(survey_bytes / 1e9, 3),
                        note="step_algorithmic_GB: bytes this call form has to move (see uvd_bytes); step_survey_GB: "
                             "SURVEY.md 8d's 4N(9r+12) model of update and apply as two separate calls")

    # ---- the same step through the reference's two separate calls (update_precond_UVd_math_, precond_grad_UVd_math) ----
    separate = None
    if form == "fused" and not args.no_separate and (world == 1 or use_graphs):
        form = "separate"
        n_warm2 = 3 + (2 * POOL if use_graphs else 0)
        for i in range(n_warm2):
            run_step(i)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(args.steps):
            run_step(n_warm2 + i)
        s1.record()
        barrier()
        t = torch.tensor([s0.elapsed_time(s1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sms = float(t.item()) / args.steps
        separate = dict(value=round(1e3 / sms, 3), unit=UNIT, ms_per_step=round(sms, 4),
                        step_frac_survey=round(survey_bytes / (sms * 1e-3) / 1e9 / peaks["hbm"], 4),
                        note="update_precond_UVd_math_ then precond_grad_UVd_math as two calls (4 sweeps + a pass over d)")
        form = "fused"

    if roofline is not None and separate is not None:
        roofline["two_call_form"] = dict(value=separate["value"], ms_per_step=separate["ms_per_step"],
                                         step_frac_survey=separate["step_frac_survey"],
                                         note="the reference's own call sequence (update_precond_UVd_math_, then "
                                              "precond_grad_UVd_math) against SURVEY 8d's 4N(9r+12) bytes")

    # ---- parity at the benchmarked configuration, outside every timed region --------------------------------
    # One more step from the state the timed steps left behind, on the GPU (the fused call the headline times) and
    # through the multi-threaded CPU twin of the oracle on the same inputs: relative Frobenius error per output.
    parity = None
    if world == 1 and not args.no_parity:
        parity = uvd_parity(psgd, U, V, d, *pool[0], step_index=args.warmup + args.steps)
        if roofline is not None:
            roofline["parity_rel_err"] = parity["max_rel_err"]

    # ---- end to end: host (pinned) inputs, host read-back, copies inside the timed region -----------
    e2e = None
    if not args.no_e2e:
        del pool
        torch.cuda.empty_cache()
        NB = 2
        hv = [torch.randn(n, 1).pin_memory() for _ in range(NB)]
        hh = [(1.3 * hv[k] + 0.1 * torch.randn(n, 1)).pin_memory() for k in range(NB)]
        hg = [torch.randn(n, 1).pin_memory() for _ in range(NB)]
        hout = [torch.empty(n, 1).pin_memory() for _ in range(NB)]
        dv = [torch.empty(n, 1, device=dev) for _ in range(NB)]
        dh = [torch.empty(n, 1, device=dev) for _ in range(NB)]
        dg = [torch.empty(n, 1, device=dev) for _ in range(NB)]
        s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        U, V, d = U0.clone(), V0.clone(), d0.clone()

        def e2e_loop(steps, base):
            in_ready = [None] * NB
            buf_free = [None] * NB
            out_done = [None] * NB
            pres = [None] * NB

            def upload(i):
                k = i % NB
                with torch.cuda.stream(s_h2d):
                    if buf_free[k] is not None:
                        s_h2d.wait_event(buf_free[k])
                    dv[k].copy_(hv[k], non_blocking=True); dh[k].copy_(hh[k], non_blocking=True)
                    dg[k].copy_(hg[k], non_blocking=True)
                    ev = torch.cuda.Event(); ev.record(s_h2d); in_ready[k] = ev

            upload(0)
            for i in range(steps):
                k = i % NB
                if i + 1 < steps:
                    upload(i + 1)
                cur.wait_event(in_ready[k])
                if out_done[k] is not None:
                    cur.wait_event(out_done[k])         # the previous result in this slot has left the device
                pres[k] = step(base + i, U, V, d, dv[k], dh[k], dg[k])
                ev = torch.cuda.Event(); ev.record(cur); buf_free[k] = ev
                with torch.cuda.stream(s_d2h):
                    s_d2h.wait_event(ev)
                    hout[k].copy_(pres[k], non_blocking=True)
                    ev2 = torch.cuda.Event(); ev2.record(s_d2h); out_done[k] = ev2
            cur.wait_stream(s_d2h)

        e2e_loop(max(2, min(args.warmup, 3)), 0)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        e2e_loop(args.steps, 2 * args.warmup)          # even base, like the warm-up's: the same two graph keys
        f1.record()
        barrier()
        ems = f0.elapsed_time(f1)
        t = torch.tensor([ems], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ems = float(t.item())
        assert np.isfinite(hout[0].numpy()).all()
        e2e = dict(value=round(args.steps / (ems / 1e3), 3), unit=UNIT, h2d_bytes_per_step=int(3 * 4 * N_total),
                   d2h_bytes_per_step=int(4 * N_total), ms_per_step=round(ems / args.steps, 3),
                   note="v,h,g uploaded from pinned host memory and pre_grad read back to host every step; U,V,d are "
                        "device-resident optimizer state; copies double-buffered on side streams")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_uvd(N_total, r, budget_s=20.0, steps=2, warmup=1)

    if rank != 0:
        return None
    return dict(
        metric=METRIC, value=round(value, 3), unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=round(ms / args.steps, 4), higher_is_better=True, scaling=args.scaling, vs_baseline=None,
        dtype="f32", data="synthetic",
        config=uvd_config(N_total, r, world),
        run=dict(rows_per_gpu=n,
                 call_form=("update_precond_and_grad_UVd (psgd_uvd_update_apply): update + apply of psgd.py:732-748 "
                            "fused into three sweeps" if form == "fused" else
                            "update_precond_UVd_math_ + precond_grad_UVd_math as two calls (" + args.uvd_form + ")"),
                 cross_gpu_exchange=exchange, cuda_graphs=bool(use_graphs), host_binding=getattr(args, "numa_binding", None)),
        parity=parity,
        step_ms_median=round(float(np.median(per_step)), 4), step_ms_max=round(max(per_step), 4), remeasured=remeasured,
        roofline=roofline, kernels=kernels, kernels_measured=kernels_from, separate_calls=separate, cpu_baseline=cpu, e2e=e2e,
        gpu_launches=int(launches), clocks=clk)


def uvd_parity(psgd, U, V, d, v, h, g, step_index):
    """CUDA (fused update+apply, the headline's kernels) against the CPU oracle at the FULL benchmarked size.

    The oracle runs twice on the same inputs: in float64 (the yardstick) and in float32 (the reference's arithmetic).
    At 1e8 rows the float32 op sequence is itself ill-conditioned in the rank-2 step of U / V (psgd.py:594-597: the
    normaliser is a difference of O(N) float32 sums; tests/test_gpu_bench_configs.py::test_uvd_at_2e7_rows), so the
    float32 oracle's own distance from float64 is reported next to ours: `max_rel_err` is CUDA vs the float64 twin."""
    import torch
    from oracle import psgd_oracle_torch as T
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    t0 = time.perf_counter()
    update_U = (step_index % 2 == 0)
    hU, hV, hd, hv, hh, hg = (x.cpu() for x in (U, V, d, v, h, g))
    pre = psgd.update_precond_and_grad_UVd(U, V, d, v, h, g, 0.01, psgd._tiny, balance=False, update_U=update_U)
    got = dict(U=U.cpu(), V=V.cpu(), d=d.cpu(), pre_grad=pre.cpu())
    Ur, Vr, dr = T.update_precond_UVd_math(hU, hV, hd, hv, hh, 0.01, balance=False, update_U=update_U)
    w32 = dict(U=Ur, V=Vr, d=dr, pre_grad=T.precond_grad_UVd_math(Ur, Vr, dr, hg))
    del Ur, Vr, dr
    dbl = lambda x: x.double()
    U6, V6, d6 = T.update_precond_UVd_math(dbl(hU), dbl(hV), dbl(hd), dbl(hv), dbl(hh), 0.01, balance=False, update_U=update_U)
    w64 = dict(U=U6, V=V6, d=d6, pre_grad=T.precond_grad_UVd_math(U6, V6, d6, dbl(hg)))
    del U6, V6, d6
    e64 = {k: rel_err_chunked(got[k], w64[k]) for k in got}
    e32 = {k: rel_err_chunked(got[k], w32[k]) for k in got}
    own = {k: rel_err_chunked(w32[k], w64[k]) for k in got}
    fmt = lambda dd: {k: float("%.3e" % e) for k, e in dd.items()}
    ok = all(e64[k] <= max(1e-5, own[k]) for k in got)
    return dict(max_rel_err=float("%.3e" % max(e64.values())), rel_err_vs_float64_oracle=fmt(e64),
                rel_err_vs_float32_oracle=fmt(e32), float32_oracle_vs_float64_oracle=fmt(own),
                tolerance=1e-5, passed=bool(ok), rows=int(U.shape[0]), branch="U" if update_U else "V",
                against="oracle/psgd_oracle_torch.py (multi-threaded CPU twin of the NumPy oracle; psgd.py:554-627) in float64 "
                        "and float32 on the same inputs, one step from the state the timed steps left; passed = every "
                        "output within 1e-5 of the float64 twin, or as close to it as the float32 oracle is",
                seconds=round(time.perf_counter() - t0, 1))


# ---------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle (port of the reference's TF op sequence) on host cores
# ---------------------------------------------------------------------------------------------
def _cpu_threads():
    """Host threads the CPU legs use: torch's intra-op pool (OpenMP element-wise kernels, MKL GEMM)."""
    import torch
    return int(torch.get_num_threads())


def _oracle_uvd_step(T, st, i):
    U, V, d, v, h, g = st
    U, V, d = T.update_precond_UVd_math(U, V, d, v, h, 0.01, balance=(i % 100 == 99), update_U=(i % 2 == 0))
    pre = T.precond_grad_UVd_math(U, V, d, g)
    return (U, V, d, v, h, g), pre


def _uvd_host_state(n, r, n_total, as_torch=False):
    """Same distributions as the GPU arm's synthetic state (run_uvd), generated on the host."""
    import torch
    gen = torch.Generator().manual_seed(2024)
    uv = (1.0 / (n_total * r)) ** 0.5
    U = torch.randn(n, r, generator=gen) * uv
    V = torch.randn(n, r, generator=gen) * uv
    d = torch.ones(n, 1)
    v = torch.randn(n, 1, generator=gen)
    h = (0.5 + 1.5 * torch.rand(n, 1, generator=gen)) * v + 0.1 * torch.randn(n, 1, generator=gen)
    g = torch.randn(n, 1, generator=gen)
    st = (U, V, d, v, h, g)
    if not as_torch:
        st = tuple(x.numpy() for x in st)
    return st


_CPU_PORT = ("multi-threaded torch-CPU port of psgd.py:554-627 (oracle/psgd_oracle_torch.py: the reference's op sequence on "
             "ATen/OpenMP/MKL kernels, checked against the NumPy oracle; TensorFlow is not installable in this image)")


def _time_cpu_uvd(n_total, r, budget_s, steps, warmup, cap_rows=20_000_000, full=False):
    """Time the CPU port.  full=True (the reference arm): every step runs ALL n_total rows.  Otherwise (the
    `cpu_baseline` leg of the GPU arm, bounded to ~20 s) a row sample is timed and scaled linearly in N -- every op of
    the path is O(N r^2)."""
    from oracle import psgd_oracle_torch as T
    if full:
        n_s = int(n_total)
    else:
        st = _uvd_host_state(200_000, r, n_total, as_torch=True)
        _oracle_uvd_step(T, st, 0)
        t0 = time.perf_counter(); _oracle_uvd_step(T, st, 0); per_row = (time.perf_counter() - t0) / 200_000
        n_s = int(min(n_total, cap_rows, max(200_000, budget_s / max(per_row, 1e-12) / max(steps + warmup, 1))))
    st = _uvd_host_state(n_s, r, n_total, as_torch=True)
    for i in range(warmup):
        st, _ = _oracle_uvd_step(T, st, i)
    t0 = time.perf_counter()
    for i in range(steps):
        st, pre = _oracle_uvd_step(T, st, warmup + i)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dt, n_s


def cpu_baseline_uvd(n_total, r, budget_s, steps, warmup):
    dt, n_s = _time_cpu_uvd(n_total, r, budget_s, steps, warmup)
    full = dt * (n_total / n_s)
    return dict(value=round(1.0 / full, 5), unit=UNIT, cores=_cpu_threads(), kind="port",
                sample=f"{_CPU_PORT} on {n_s:,} of {n_total:,} rows, {steps} timed steps, {dt * 1e3:.0f} ms/step on the "
                       f"sample, scaled linearly to the full vector; host has {os.cpu_count()} logical cores, "
                       f"{_cpu_threads()} threads used",
                ms_per_step_sample=round(dt * 1e3, 2), sample_rows=n_s)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  TensorFlow is not installable in this
    image (no network), so this is the op-for-op CPU port on all host threads; rank 0 only."""
    if rank != 0:
        return None
    import torch
    torch.set_num_threads(max(1, os.cpu_count() or 1))      # torchrun pins OMP_NUM_THREADS=1 per rank; this arm is rank 0 alone
    n_total, r = args.n, args.rank
    if args.workload == "kron":
        return run_reference_kron(args, full=True, world=world)
    if args.workload == "all":
        args.workload = "uvd"
        out = run_reference(args, rank, world)
        kr = run_reference_kron(args, full=True, world=world)
        out["kron"] = {k: kr[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "config", "cpu_baseline", "e2e")}
        args.workload = "all"
        return out
    # the stated configuration, not a sample: every one of the K timed steps runs all n_total rows (10 GB of host state
    # at 1e8 x rank 10, ~5 s per step on 16 threads)
    dt, n_s = _time_cpu_uvd(n_total, r, 0.0, args.steps, args.warmup, full=True)
    assert n_s == n_total
    full_ms = dt * 1e3
    value = 1e3 / full_ms
    sample = (f"{_CPU_PORT}; FULL size: all {n_total:,} rows in each of the {args.steps} timed steps "
              f"({args.warmup} warm-up), nothing scaled; {os.cpu_count()} logical cores, {_cpu_threads()} threads used")
    return dict(impl="reference", metric=METRIC, value=round(value, 5), unit=UNIT, n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=round(full_ms, 2), higher_is_better=True, scaling=args.scaling,
                vs_baseline=None, dtype="f32", data="synthetic",
                config=uvd_config(n_total, r, world),
                cpu_baseline=dict(value=round(value, 5), unit=UNIT, cores=_cpu_threads(), kind="port", sample=sample),
                e2e=dict(value=round(value, 5), unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))


def run_reference_kron(args, full=True, world=1):
    """The CPU port of psgd.py:156-192 on the Kron stack.  full=True (the reference arm): ALL `layers` layers at the
    full n x n size in every step -- at ~1.2 s per 4096^2 layer-step on 16 threads a 24-layer step takes ~30 s, so the
    step count is bounded (1 warm-up + at most 2 timed steps) to keep the whole arm within a few minutes; the figure
    is measured, not scaled.  full=False (the `cpu_baseline` leg of the GPU arm, ~10-20 s): ONE full-size layer per
    step, scaled by the layer count only (layers are independent and identical in cost)."""
    import torch
    from oracle import psgd_oracle_torch as T
    n = args.kron_n
    L = args.layers
    Lrun = L if full else 1
    steps = max(1, min(args.steps, 2))
    warmup = min(args.warmup, 1)
    gen = torch.Generator().manual_seed(1000)
    Ql = [torch.eye(n) for _ in range(Lrun)]
    Qr = [torch.eye(n) for _ in range(Lrun)]
    S = [0.5 + 1.5 * torch.rand(n, 1, generator=gen) for _ in range(Lrun)]
    Tm = [0.5 + 1.5 * torch.rand(1, n, generator=gen) for _ in range(Lrun)]
    dX = [torch.randn(n, n, generator=gen) for _ in range(Lrun)]
    dG = [s * x * t + 0.1 * torch.randn(n, n, generator=gen) for s, x, t in zip(S, dX, Tm)]
    G = [torch.randn(n, n, generator=gen) for _ in range(Lrun)]

    def one_step():
        for l in range(Lrun):
            Ql[l], Qr[l] = T.update_precond_dense_dense(Ql[l], Qr[l], dX[l], dG[l], 0.01)
            T.precond_grad_dense_dense(Ql[l], Qr[l], G[l])
    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / steps
    step_s = dt * (L / Lrun)
    value = 1.0 / step_s
    if full:
        sample = (f"multi-threaded torch-CPU port of psgd.py:156-192 (oracle/psgd_oracle_torch.py; TensorFlow unavailable); FULL "
                  f"size: all {L} layers of {n}x{n} in each of {steps} timed steps ({warmup} warm-up; the requested "
                  f"--steps {args.steps} --warmup {args.warmup} are capped for this nested workload), nothing scaled; "
                  f"{os.cpu_count()} logical cores, {_cpu_threads()} threads used")
    else:
        sample = (f"multi-threaded torch-CPU port of psgd.py:156-192 (oracle/psgd_oracle_torch.py; TensorFlow unavailable): one "
                  f"full-size {n}x{n} dense-dense layer per step ({dt * 1e3:.0f} ms, {steps} timed steps), times {L} "
                  f"independent layers; {os.cpu_count()} logical cores, {_cpu_threads()} threads used")
    return dict(impl="reference", metric=METRIC, value=round(value, 6), unit=UNIT, n_gpus=world, steps=steps,
                warmup=warmup, ms_per_step=round(step_s * 1e3, 1), higher_is_better=True, scaling=args.scaling,
                vs_baseline=None, dtype="f32", data="synthetic",
                config=kron_config(L, n, world),
                cpu_baseline=dict(value=round(value, 6), unit=UNIT, cores=_cpu_threads(), kind="port", sample=sample),
                e2e=dict(value=round(value, 6), unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))


# ---------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def _claim_stdout():
    """Rank 0 must print exactly ONE JSON line on stdout, but native libraries write there too (NCCL prints its version
    banner to stdout on the first communicator).  Point fd 1 at stderr for the whole run and keep the real stdout aside
    for the final line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def main():
    _claim_stdout()
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "uvd", "kron"],
                    help="all (default): UVd 100M is the headline line, the Kron stack result is nested under 'kron'")
    ap.add_argument("--n", type=int, default=100_000_000, help="UVd: parameters in the flattened vector")
    ap.add_argument("--rank", type=int, default=10)
    ap.add_argument("--layers", type=int, default=24)
    ap.add_argument("--kron-n", type=int, default=4096)
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--nccl-hook", action="store_true", help="UVd multi-GPU: use the torch.distributed all-reduce hook "
                    "instead of the peer-memory exchange kernel")
    ap.add_argument("--no-graphs", action="store_true", help="UVd: launch every kernel from Python instead of replaying "
                    "the step from CUDA graphs")
    ap.add_argument("--uvd-form", default="fused", choices=["fused", "separate", "separate-3sweep"],
                    help="UVd: fused update+apply call (default), the reference's two calls, or those with the three-sweep update")
    ap.add_argument("--no-separate", action="store_true", help="UVd: skip the extra timing of the two-call form")
    ap.add_argument("--kron-gather", default="auto", choices=["auto", "once", "slots", "peer", "peer-once"],
                    help="Kron multi-GPU, how the preconditioned gradients reach every rank: auto = slots at <= 3 layers per GPU, "
                         "else once; once = one in-place NCCL "
                         "all-gather after the batched apply; slots = NCCL all-gather per layer slot as its apply finishes; "
                         "peer = copy-engine pushes into IPC-mapped peer buffers per layer slot (no SMs); peer-once = the same "
                         "after the whole batched apply")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the CUDA-vs-oracle check at the benchmarked size")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank, world, local = dist_env()

    if args.impl == "reference":
        out = run_reference(args, rank, world)
        if out is not None:
            emit(out)
        return 0

    args.warmup = max(args.warmup, 3)       # timing rule: at least 3 warm-up steps
    import torch
    if not torch.cuda.is_available():
        emit({"error": "no CUDA device: psgd_tf_b200 has no CPU path"})
        return 1
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        args.numa_binding = bind_to_gpu_numa(local)
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        out = None
        if args.workload in ("uvd", "all"):
            out = run_uvd(args, rank, world, local)
            import gc
            gc.collect()
            torch.cuda.empty_cache()
        if args.workload in ("kron", "all"):
            from bench_kron import run_kron
            kr = run_kron(args, rank, world, local, METRIC, UNIT, load_peaks, ClockSampler)
            if args.workload == "kron":
                out = kr
            elif out is not None and kr is not None:
                out["kron"] = {k: kr[k] for k in ("value", "unit", "ms_per_step", "scaling", "dtype", "config", "run", "roofline",
                                                   "kernels", "parity", "cpu_baseline", "e2e", "gpu_launches", "clocks")}
                if out.get("roofline") is not None:       # the second north-star workload's headline inside the parsed object
                    out["roofline"]["kron_stack"] = dict(value=kr["value"], unit=kr["unit"], ms_per_step=kr["ms_per_step"],
                                                         kernel_frac=(kr["roofline"] or {}).get("frac"),
                                                         step_frac=(kr["roofline"] or {}).get("step_frac"),
                                                         parity_rel_err=(kr["parity"] or {}).get("max_rel_err"))
        if args.workload == "all" and world == 1 and out is not None:
            # the other streaming rows of SURVEY.md section 8(a): diagonal, X-shape, (norm,scale) Kron pair, dense apply
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            from bench_aux import run_aux
            out["aux"] = run_aux(load_peaks()["hbm"], steps=10)
        if args.workload == "all" and world > 1:
            # diagonal / X-shape on the sharded vector: every rank takes part, rank 0 reports
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            from bench_aux import run_aux_sharded
            rows = run_aux_sharded(load_peaks()["hbm"], rank, world)
            if out is not None:
                out["aux"] = rows
        if out is not None:
            emit(out)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
