"""Warm per-kernel device times (CUPTI through torch.profiler) of one small Kron layer's update+apply."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import psgd_tf_b200 as psgd
from bench_aux import _factor

M, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (257, 120)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(11)
Ql, Qr = _factor(torch, "dense", M, dev), _factor(torch, "dense", N, dev)
dX = torch.randn(M, N, device=dev, generator=g); dG = 1.3 * dX + 0.1 * torch.randn(M, N, device=dev, generator=g)
G = torch.randn(M, N, device=dev, generator=g)


def step():
    global Ql, Qr
    Ql, Qr = psgd.update_precond_kron(Ql, Qr, dX, dG, 0.01)
    return psgd.precond_grad_kron(Ql, Qr, G)


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10):
        step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    if e.device_time_total > 0:
        rows.append((e.device_time_total / 10.0, e.count // 10, e.key[:90]))
for us, cnt, name in sorted(rows, reverse=True):
    print(f"{us:8.1f} us/step  x{cnt:2d}  {name}")
print("sum", sum(r[0] for r in rows))

# launch order of the last profiled step, with each kernel's own duration and the gap before it
evs = sorted([e for e in prof.events() if e.device_type.name == "CUDA" and e.device_time > 0], key=lambda e: e.time_range.start)
per = len(evs) // 10
last = evs[-per:]
prev_end = None
for e in last:
    gap = (e.time_range.start - prev_end) if prev_end is not None else 0.0
    print(f"  +{gap:6.1f} us gap  {e.device_time:7.1f} us  {e.name[:80]}")
    prev_end = e.time_range.end
