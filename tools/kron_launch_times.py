"""Per-launch CUDA-event times of one 24 x 4096^2 Kron update+apply step (library profile mode), in launch order."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import psgd_tf_b200 as psgd
L, n = int(os.environ.get("LAYERS", 24)), 4096
dev = "cuda"
ctx = psgd.get_context()
Ql = [torch.eye(n, device=dev) for _ in range(L)]; Qr = [torch.eye(n, device=dev) for _ in range(L)]
dX = [torch.randn(n, n, device=dev) for _ in range(L)]; dG = [1.3 * x + 0.1 * torch.randn(n, n, device=dev) for x in dX]
G = [torch.randn(n, n, device=dev) for _ in range(L)]
def step(Ql, Qr):
    new = psgd.update_precond_kron_batched(Ql, Qr, dX, dG, 0.01)
    Ql, Qr = [a for a, _ in new], [b for _, b in new]
    return Ql, Qr, psgd.precond_grad_kron_batched(Ql, Qr, G)
for _ in range(3):
    Ql, Qr, pre = step(Ql, Qr)
torch.cuda.synchronize()
ctx.set_option("profile", 1); ctx.profile_read(cap=1 << 16)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); Ql, Qr, pre = step(Ql, Qr); e1.record(); torch.cuda.synchronize()
prof = ctx.profile_read(cap=1 << 16)
print("step ms", e0.elapsed_time(e1), "profiled launches", len(prof), "sum ms", sum(p[1] for p in prof))
for i, (kid, ms, work) in enumerate(prof):
    if kid == 10:
        print(f"{i:4d} gemm {ms:8.3f} ms  {work / 1e9:9.1f} GFLOP(dense)  {work / ms / 1e9:7.1f} TFLOP/s")
    else:
        print(f"{i:4d} k{kid}   {ms:8.3f} ms")
