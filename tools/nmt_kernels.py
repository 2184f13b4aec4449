"""Warm per-kernel device times (CUPTI through torch.profiler) of the cfg5 NMT layer list (or one of its layers):
update + apply through the batched calls, eager, groups on the side streams."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import psgd_tf_b200 as psgd
from bench_aux import _factor

nmt = [("scale", "dense", 9414, 256), ("norm", "scale", 1281, 1024), ("scale", "dense", 2048, 10),
       ("dense", "dense", 1, 10), ("scale", "dense", 4935, 256), ("norm", "scale", 2305, 1024),
       ("norm", "scale", 1025, 4935)]
if len(sys.argv) > 1:
    nmt = [nmt[int(i)] for i in sys.argv[1].split(",")]
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(11)
Ql = [_factor(torch, kl, M, dev) for kl, kr, M, N in nmt]
Qr = [_factor(torch, kr, N, dev) for kl, kr, M, N in nmt]
dX = [torch.randn(M, N, device=dev, generator=g) for _, _, M, N in nmt]
dG = [1.3 * x + 0.1 * torch.randn(x.shape, device=dev, generator=g) for x in dX]
G = [torch.randn(M, N, device=dev, generator=g) for _, _, M, N in nmt]


def step():
    global Ql, Qr
    new = psgd.update_precond_kron_batched(Ql, Qr, dX, dG, 0.01)
    Ql, Qr = [a for a, _ in new], [b for _, b in new]
    return psgd.precond_grad_kron_batched(Ql, Qr, G)


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10):
        step()
    torch.cuda.synchronize()
rows = [(e.device_time_total / 10.0, e.count / 10.0, e.key[:100]) for e in prof.key_averages() if e.device_time_total > 0]
for us, cnt, name in sorted(rows, reverse=True):
    print(f"{us:8.1f} us/step  x{cnt:5.1f}  {name}")
print("layers", nmt, "sum us", round(sum(r[0] for r in rows), 1))
if len(nmt) == 1:
    evs = sorted([e for e in prof.events() if e.device_type.name == "CUDA" and e.device_time > 0], key=lambda e: e.time_range.start)
    per = len(evs) // 10
    prev_end = None
    for e in evs[-per:]:
        gap = (e.time_range.start - prev_end) if prev_end is not None else 0.0
        print(f"  +{gap:6.1f} us gap  {e.device_time:7.1f} us  {e.name[:90]}")
        prev_end = e.time_range.end
