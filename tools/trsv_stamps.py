"""clock64() stamps of the first panel of the dense preconditioner's vector solve (csrc/linalg.cu trsv_panel_kernel):
cycles per phase -- start, right-hand side in shared memory, diagonal blocks inverted, then per block step:
[step top, warp-0 product done, barrier passed, update + next loads issued]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import psgd_tf_b200 as psgd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(11)
Q = torch.triu(torch.randn(n, n, device=dev, generator=g)) * 0.01 + torch.eye(n, device=dev)
dx = [torch.randn(n, device=dev, generator=g)]
dg = [1.3 * dx[0] + 0.1 * torch.randn(n, device=dev, generator=g)]
buf = torch.zeros(256, dtype=torch.int64, device=dev)
ctx = psgd.get_context()
for _ in range(3):
    psgd.update_precond_dense(Q, dx, dg, 0.01)
ctx.set_option("stamp_ptr", buf.data_ptr())
psgd.update_precond_dense(Q, dx, dg, 0.01)
torch.cuda.synchronize()
ctx.set_option("stamp_ptr", 0)
t = buf.cpu().tolist()
t = [x for x in t if x != 0]
t0 = t[0]
print("n", n, "total cycles", t[-1] - t0)
print("rhs+prefetch", t[1] - t[0], "| inversion", t[2] - t[1], "| barrier", t[3] - t[2])
steps = t[4:-1]
for i in range(0, len(steps) - 3, 4):
    a, b, c, d = steps[i:i + 4]
    nxt = steps[i + 4] if i + 4 < len(steps) else t[-1]
    print(f"step {i // 4:2d}: warp0 product {b - a:6d} | barrier A {c - b:6d} | update+loads {d - c:6d} | barrier B {nxt - d:6d}")
