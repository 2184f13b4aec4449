"""Which tensor-core GEMM launch of a Kron update+apply goes wrong when it runs on the CTA-pair kernel?  Runs the layer
with the pair kernel enabled for ONE launch at a time (psgd_set_option("tc_pair_sel", k)) and prints the error against
the single-CTA result of the same library."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import psgd_tf_b200 as psgd
from tests import cases

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ctx = psgd.get_context()
ctx.set_option("gemm_path", 2)
c = cases.kron_case(7000 + 2 * n, "dense", "dense", n, n)
dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
ins = {k: dev(v) for k, v in c.items()}


def run():
    ql, qr = psgd.update_precond_kron(ins["Ql"], ins["Qr"], ins["dX"], ins["dG"], 0.01)
    pre = psgd.precond_grad_kron(ql, qr, ins["G"])
    torch.cuda.synchronize()
    return ql, qr, pre


rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
ctx.set_option("tc_pair", 0)
ref = run()
ctx.set_option("tc_pair", 1)
ctx.set_option("tc_pair_sel", -1)
allp = run()
print("all launches on the pair kernel:", [f"{rel(a, b):.2e}" for a, b in zip(allp, ref)], flush=True)
ctx.set_option("tc_pair_sel", 10 ** 6)
nlaunch = None
run()
import ctypes
for k in range(64):
    ctx.set_option("tc_pair_sel", k)
    got = run()
    e = [rel(a, b) for a, b in zip(got, ref)]
    flag = "  <-- BAD" if max(e) > 1e-5 else ""
    print(f"sel {k:2d}: " + " ".join(f"{x:.2e}" for x in e) + flag, flush=True)
