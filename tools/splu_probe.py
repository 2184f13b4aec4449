"""SPLU update+apply at n = 5e7, r = 10 (the bench_aux row) in isolation: timing, or a target for ncu."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import psgd_tf_b200 as psgd

n, r = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000, 10
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(7)
L12 = torch.cat([torch.eye(r, device=dev), torch.zeros(n - r, r, device=dev)])
U12 = torch.cat([torch.eye(r, device=dev), torch.zeros(r, n - r, device=dev)], 1)
l3 = torch.ones(n - r, 1, device=dev); u3 = torch.ones(n - r, 1, device=dev)
dx = torch.randn(n, device=dev, generator=g); dg = 1.3 * dx + 0.1 * torch.randn(n, device=dev, generator=g)
gg = torch.randn(n, device=dev, generator=g)
st = [L12, l3, U12, u3]
for i in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = psgd.update_precond_splu(*st, [dx], [dg], 0.01)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    pre = psgd.precond_grad_splu(*st, [gg])
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"update {1e3 * (t1 - t0):.2f} ms  apply {1e3 * (t2 - t1):.2f} ms", flush=True)
