"""Warm per-kernel device times (CUPTI through torch.profiler) of one dense-preconditioner update (psgd.py:26-44)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import psgd_tf_b200 as psgd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(11)
Q = torch.triu(torch.randn(n, n, device=dev, generator=g)) * 0.01 + torch.eye(n, device=dev)
dx = [torch.randn(n, device=dev, generator=g)]
dg = [1.3 * dx[0] + 0.1 * torch.randn(n, device=dev, generator=g)]
for _ in range(3):
    Q = psgd.update_precond_dense(Q, dx, dg, 0.01)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    Q = psgd.update_precond_dense(Q, dx, dg, 0.01)
e1.record(); torch.cuda.synchronize()
print("wall ms per update", round(e0.elapsed_time(e1) / 5, 3))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        Q = psgd.update_precond_dense(Q, dx, dg, 0.01)
    torch.cuda.synchronize()
rows = [(e.device_time_total / 5.0, e.count // 5, e.key[:100]) for e in prof.key_averages() if e.device_time_total > 0]
for us, cnt, name in sorted(rows, reverse=True):
    print(f"{us:9.1f} us/step  x{cnt:3d}  {name}")
print("n", n, "sum us", round(sum(r[0] for r in rows), 1))
