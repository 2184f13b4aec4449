"""clock64() stamps of CTA 0 of the two panel solves of one small dense-dense Kron layer (csrc/linalg.cu
trsm_panel_kernel): cycles for the slab load, the block inversions and, per block step, the off-diagonal tiles and the
diagonal product."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import psgd_tf_b200 as psgd
from bench_aux import _factor

M, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (257, 120)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(11)
Ql, Qr = _factor(torch, "dense", M, dev), _factor(torch, "dense", N, dev)
dX = torch.randn(M, N, device=dev, generator=g); dG = 1.3 * dX
buf = torch.zeros(256, dtype=torch.int64, device=dev)
ctx = psgd.get_context()
for _ in range(3):
    psgd.update_precond_kron(Ql, Qr, dX, dG, 0.01)
ctx.set_option("stamp_ptr", buf.data_ptr())
psgd.update_precond_kron(Ql, Qr, dX, dG, 0.01)
torch.cuda.synchronize()
ctx.set_option("stamp_ptr", 0)
for name, off in (("left solve (n = M)", 0), ("right solve (n = N)", 128)):
    t = [x for x in buf[off:off + 128].cpu().tolist() if x != 0]
    if len(t) < 4:
        continue
    print(name, "total cycles", t[-1] - t[0], "| slab load", t[1] - t[0], "| inversion", t[2] - t[1])
    steps = t[3:]
    for i in range(0, len(steps) - 2, 3):
        a, b, c = steps[i:i + 3]
        print(f"  step {i // 3:2d}: (barrier+) tiles {b - a:6d} | reduce + diagonal product {c - b:6d}")
