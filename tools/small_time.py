"""Where does a small ragged Kron step spend its time?  Graph-replayed update+apply of each LeNet5 layer alone and of the
whole list, with the groups on side streams and serialised."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import psgd_tf_b200 as psgd
from psgd_tf_b200.graphs import KronStepGraphs
from bench_aux import _factor, _time

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(11)
ctx = psgd.get_context()
shapes = ((26, 6), (151, 16), (257, 120), (121, 84), (85, 10))


def graph_us(layers, streams):
    ctx.set_option("kron_streams", streams)
    Ql = [_factor(torch, "dense", M, dev) for M, N in layers]
    Qr = [_factor(torch, "dense", N, dev) for M, N in layers]
    sets = []
    for _ in range(2):
        dX = [torch.randn(M, N, device=dev, generator=g) for M, N in layers]
        dG = [1.3 * x + 0.1 * torch.randn(x.shape, device=dev, generator=g) for x in dX]
        G = [torch.randn(M, N, device=dev, generator=g) for M, N in layers]
        sets.append((dX, dG, G))
    gr = KronStepGraphs(Ql, Qr, 0.01)
    k = [0]

    def step():
        dX, dG, G = sets[k[0] & 1]; k[0] += 1
        return gr.step(dX, dG, G)
    for _ in range(6):
        step()
    return _time(torch, step, 50) * 1e3


for s in shapes:
    print(f"layer {s}: {graph_us([s], 1):7.1f} us per graph-replayed step", flush=True)
print(f"all five, side streams: {graph_us(shapes, 1):7.1f} us", flush=True)
print(f"all five, one stream  : {graph_us(shapes, 0):7.1f} us", flush=True)
