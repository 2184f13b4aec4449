"""One (normalization, scaling) update + apply at [8192, 8192] for an ncu launch list (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import psgd_tf_b200 as psgd
M = N = 8192
dev = "cuda"
ql = torch.stack([torch.ones(M, device=dev), torch.zeros(M, device=dev)])
qr = torch.ones(1, N, device=dev)
dX = torch.randn(M, N, device=dev); dG = 1.3 * dX + 0.1 * torch.randn(M, N, device=dev); G = torch.randn(M, N, device=dev)
for _ in range(3):
    a, b = psgd.update_precond_kron(ql, qr, dX, dG, 0.01)
    p = psgd.precond_grad_kron(a, b, G)
torch.cuda.synchronize()
