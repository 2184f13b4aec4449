"""One batched Kron update+apply step on a small ragged layer list (cfg1 LeNet5 or cfg5 NMT), a few times, for an ncu
launch list:  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:. -s <skip> python tools/small_probe.py lenet"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import psgd_tf_b200 as psgd
from bench_aux import _factor

which = sys.argv[1] if len(sys.argv) > 1 else "lenet"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(11)
if which == "lenet":
    layers = [("dense", "dense", M, N) for M, N in ((26, 6), (151, 16), (257, 120), (121, 84), (85, 10))]
else:
    layers = [("scale", "dense", 9414, 256), ("norm", "scale", 1281, 1024), ("scale", "dense", 2048, 10),
              ("dense", "dense", 1, 10), ("scale", "dense", 4935, 256), ("norm", "scale", 2305, 1024), ("norm", "scale", 1025, 4935)]
Ql = [_factor(torch, kl, M, dev) for kl, kr, M, N in layers]
Qr = [_factor(torch, kr, N, dev) for kl, kr, M, N in layers]
dX = [torch.randn(M, N, device=dev, generator=g) for _, _, M, N in layers]
dG = [1.3 * x + 0.1 * torch.randn(x.shape, device=dev, generator=g) for x in dX]
G = [torch.randn(M, N, device=dev, generator=g) for _, _, M, N in layers]
ctx = psgd.get_context()
if len(sys.argv) > 3:
    ctx.set_option("kron_streams", int(sys.argv[3]))
for _ in range(steps):
    new = psgd.update_precond_kron_batched(Ql, Qr, dX, dG, 0.01)
    Ql, Qr = [a for a, _ in new], [b for _, b in new]
    pre = psgd.precond_grad_kron_batched(Ql, Qr, G)
torch.cuda.synchronize()
print("ok", which, ctx.launch_count)
