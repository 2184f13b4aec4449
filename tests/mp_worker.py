"""Multi-process worker used by the distributed tests (launched with torch.distributed.run / mp.spawn).

mode "cpu-protocol": gloo, no GPU -- checks the host-side sharding logic: chunk bounds tile the vector, the layer
  assignment covers every layer once, ragged/uniform all-gathers return every layer on every rank, and the reduction
  protocol of the chunk-sharded UVd update (what is summed, what is maxed) reproduces the unsharded oracle.
mode "gpu-uvd" / "gpu-kron": nccl, one GPU per rank -- the real CUDA path with the all-reduce hook / layer sharding
  against the oracle on the full problem.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import psgd_oracle as O          # noqa: E402
from tests import cases                       # noqa: E402


def cpu_protocol(rank, world):
    from psgd_tf_b200 import partition
    # ---- chunk bounds tile [0, n) -------------------------------------------------------------------
    for n in (1, 255, 256, 1021, 100_003):
        bounds = [partition.chunk_bounds(n, world, r) for r in range(world)]
        assert bounds[0][0] == 0 and bounds[-1][1] == n
        for a, b in zip(bounds, bounds[1:]):
            assert a[1] == b[0] and (a[0] % 256 == 0 or a[0] == a[1])
    (lo, hi), (mlo, mhi) = partition.mirrored_chunk_bounds(1001, world, rank)
    assert (mlo, mhi) == (1001 - hi, 1001 - lo)
    # ---- X-shape on mirrored chunk pairs (+ the centre for odd n): the shards partition the vector, every local
    # problem is flip-symmetric, and local X-shape math with ONE max all-reduce reproduces the unsharded oracle -------
    for n in (1, 2, 7, 1000, 1001, 70_001):
        sl = [partition.xmat_shard_slices(n, world, k) for k in range(world)]
        seen = np.zeros(n, int)
        for parts in sl:
            for a, b in parts:
                seen[a:b] += 1
        assert (seen == 1).all(), (n, sl)
        idx = np.concatenate([np.arange(a, b) for a, b in sl[rank]]).astype(int)
        assert (idx + idx[::-1] == n - 1).all()                        # local flip == global flip
        c = cases.vec_case(n, n)
        la, lb, lv, lh = (c[k][idx].astype(np.float64) for k in ("a", "b", "v", "h"))
        flip = lambda x: x[::-1]
        Qh = la * lh + lb * flip(lh)
        iq = (flip(la) * lv - flip(lb) * flip(lv)) / (la * flip(la) - lb * flip(lb))
        na, nb = Qh * Qh - iq * iq, Qh * flip(Qh) - iq * flip(iq)
        if len(idx) % 2 == 1:
            nb[len(idx) // 2] = 0
        mx = torch.tensor([max(np.abs(na).max(), np.abs(nb).max()) if len(idx) else 0.0], dtype=torch.float64)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        mu = 0.01 / (mx.item() + float(O.TINY))
        an, bn = la - mu * (na * la + nb * flip(lb)), lb - mu * (na * lb + nb * flip(la))
        f64 = {k: c[k].astype(np.float64) for k in c}
        ar, br = O.update_precond_Xmat(f64["a"], f64["b"], f64["v"], f64["h"], 0.01)
        np.testing.assert_allclose(an, ar[idx], rtol=1e-10)
        np.testing.assert_allclose(bn, br[idx], rtol=1e-10, atol=1e-300)
    # ---- layer assignment + all-gather -----------------------------------------------------------------
    shapes = cases.LENET_SHAPES
    owned = partition.assign_layers([partition.kron_layer_cost(m, n) for m, n in shapes], world)
    assert sorted(i for o in owned for i in o) == list(range(len(shapes)))
    local = [torch.full(shapes[i], float(i)) for i in owned[rank]]
    full = partition.all_gather_layers(local, owned, shapes, rank)
    for i, t in enumerate(full):
        assert tuple(t.shape) == shapes[i] and torch.all(t == float(i))
    uni = [(8, 8)] * 4
    owned_u = partition.assign_layers([1.0] * 4, world)
    full = partition.all_gather_layers([torch.full((8, 8), float(i)) for i in owned_u[rank]], owned_u, uni, rank)
    for i, t in enumerate(full):
        assert torch.all(t == float(i))
    gb = partition.KronGatherBuffer(uni, owned_u, rank, torch.device("cpu")) if len({len(o) for o in owned_u}) == 1 else None
    for rep_ in range(3 if gb is not None else 0):          # two alternating buffers
        for j, o in enumerate(gb.local_outs()):
            o.fill_(float(owned_u[rank][j] + 10 * rep_))
        if rep_ == 1:                                        # slot by slot, as the overlapped form of the bench issues them
            for j in range(gb.per):
                gb.gather_slot(j)
            full = gb.finish()
        else:
            full = gb.gather()
        for i, t in enumerate(full):
            assert torch.all(t == float(i + 10 * rep_))
    # ---- reduction protocol of the sharded UVd update (mirrors psgd_tf_b200/csrc/uvd.cu) ------------------------
    n, r = 1021, 4
    c = cases.uvd_case(42, n, r)
    lo, hi = partition.chunk_bounds(n, world, rank)
    U, V, d, v, h = (c[k][lo:hi].astype(np.float64) for k in ("U", "V", "d", "v", "h"))
    Z = np.concatenate([U, V], 1)
    X = np.concatenate([d * h, v / d], 1)
    G = torch.from_numpy(Z.T @ np.concatenate([Z, X], 1))
    dist.all_reduce(G)                                                  # sweep 1: sums
    G = G.numpy()
    UtU, VtU, VtV = G[:r, :r], G[r:2 * r, :r], G[r:2 * r, r:2 * r]
    p, Utdh, Utw, Vtw = G[r:2 * r, 2 * r], G[:r, 2 * r], G[:r, 2 * r + 1], G[r:2 * r, 2 * r + 1]
    IpVtU = np.eye(r) + VtU
    t = Utdh + UtU @ p
    s1 = np.linalg.solve(IpVtU.T, Utw)
    s2 = np.linalg.solve(IpVtU, Vtw - VtV @ s1)
    Qh = (d * h)[:, 0] + U @ p
    Ph = d[:, 0] * (Qh + V @ t)
    b = (v / d)[:, 0] - V @ s1
    invPv = (b - U @ s2) / d[:, 0]
    nabla = Ph * h[:, 0] - v[:, 0] * invPv
    mx = torch.tensor([np.abs(nabla).max() if hi > lo else 0.0], dtype=torch.float64)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)                           # sweep 2: one max ...
    sums = torch.from_numpy(np.concatenate([[Qh @ Qh, b @ b, Qh @ b], Qh @ V, b @ V]))
    dist.all_reduce(sums)                                               # ... and 3 + 2r sums
    sums = sums.numpy()
    aa, bb, ab, atV, btV = sums[0], sums[1], sums[2], sums[3:3 + r], sums[3 + r:]
    norm = np.sqrt(abs(aa * (atV @ VtV @ atV) + bb * (btV @ VtV @ btV) - 2 * ab * (atV @ VtV @ btV)))
    tiny = float(O.TINY)
    d_new = d[:, 0] - 0.01 / (mx.item() + tiny) * d[:, 0] * nabla
    U_new = U - 0.01 / (norm + tiny) * (np.outer(Qh, atV @ IpVtU) - np.outer(b, btV @ IpVtU))
    f64 = {k: c[k].astype(np.float64) for k in c}
    Ur, Vr, dr = O.update_precond_UVd_math(f64["U"], f64["V"], f64["d"], f64["v"], f64["h"], 0.01, update_U=True)
    np.testing.assert_allclose(U_new, Ur[lo:hi], rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(d_new, dr[lo:hi, 0], rtol=1e-9)
    # ---- protocol of the FUSED forms (psgd_uvd_update with "uvd_fused", psgd_uvd_update_apply): exchange 1 = the
    # (2r+2)^2 Gram table, exchange 2 = max|nablaD| (+ 4r sums over the updated rows for the fused apply) -------------
    from tests import uvd_pipeline_model as M
    for update_U in (True, False):
        sh = {k: c[k][lo:hi] for k in c}
        G = torch.from_numpy(M.gram_table(sh["U"], sh["V"], sh["d"], sh["h"], sh["v"]))
        dist.all_reduce(G)                                              # exchange 1: sums
        k = M.small1(G.numpy(), r, update_U, 0.01)
        Un, Vn, nd = M.fused_map(sh["U"], sh["V"], sh["d"], sh["h"], sh["v"], k, update_U)
        x0 = (sh["d"] * sh["g"]).astype(np.float32)
        x1 = (x0 * nd).astype(np.float32)
        Z = np.concatenate([Un, Vn], 1).astype(np.float64)
        sums = torch.from_numpy(np.concatenate([Z.T @ x0, Z.T @ x1], 1))                 # [2r, 2]
        mx = torch.tensor([np.abs(nd).max() if hi > lo else 0.0], dtype=torch.float64)
        dist.all_reduce(sums)                                           # exchange 2: 4r sums ...
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)                       # ... and one max
        sums = sums.numpy()
        mu_d = np.float32(0.01) / (np.float32(mx.item()) + O.TINY)
        pp = sums[r:, 0] - np.float64(mu_d) * sums[r:, 1]
        tt = sums[:r, 0] - np.float64(mu_d) * sums[:r, 1] + M.updated_UtU(k, update_U) @ pp
        dn = (sh["d"] - (mu_d * sh["d"]) * nd).astype(np.float32)
        y = dn * sh["g"] + Un @ pp.astype(np.float32)[:, None]
        pre = dn * (y + Vn @ tt.astype(np.float32)[:, None])
        Ur, Vr, dr = O.update_precond_UVd_math(c["U"], c["V"], c["d"], c["v"], c["h"], 0.01, update_U=update_U)
        pr = O.precond_grad_UVd_math(Ur, Vr, dr, c["g"])
        for got, want in ((Un, Ur), (Vn, Vr), (dn, dr), (pre, pr)):
            if hi > lo:
                assert np.linalg.norm(got.astype(np.float64) - want[lo:hi]) / np.linalg.norm(want) < 2e-6


def gpu_uvd(rank, world, exchange="hook"):
    import psgd_tf_b200 as psgd
    from psgd_tf_b200 import partition
    torch.cuda.set_device(rank)
    ctx = psgd.get_context(rank)
    if exchange == "peer":
        partition.install_peer_exchange(ctx)      # must work on a multi-GPU box: no silent fallback in this test
    else:
        partition.install_allreduce(ctx)
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    for n, r in ((100_003, 10), (1021, 10), (5000, 16)):
        c = cases.uvd_case(600 + n, n, r)
        lo, hi = partition.chunk_bounds(n, world, rank)
        for kw in (dict(update_U=True, balance=False), dict(update_U=False, balance=True)):
            U, V, d = dev(c["U"][lo:hi]), dev(c["V"][lo:hi]), dev(c["d"][lo:hi])
            psgd.update_precond_UVd_math_(U, V, d, dev(c["v"][lo:hi]), dev(c["h"][lo:hi]), 0.01, psgd._tiny, **kw)
            pre = psgd.precond_grad_UVd_math(U, V, d, dev(c["g"][lo:hi]))
            Ur, Vr, dr = O.update_precond_UVd_math(c["U"], c["V"], c["d"], c["v"], c["h"], 0.01, **kw)
            pr = O.precond_grad_UVd_math(Ur, Vr, dr, c["g"])
            for got, want in ((U, Ur), (V, Vr), (d, dr), (pre, pr)):
                if hi > lo:
                    e = np.linalg.norm(got.cpu().numpy().astype(np.float64) - want[lo:hi]) / np.linalg.norm(want)
                    assert e < 1e-5, (n, r, kw, e)
            # the fused update+apply call: same results, two exchanges instead of three
            U, V, d = dev(c["U"][lo:hi]), dev(c["V"][lo:hi]), dev(c["d"][lo:hi])
            pre = psgd.update_precond_and_grad_UVd(U, V, d, dev(c["v"][lo:hi]), dev(c["h"][lo:hi]), dev(c["g"][lo:hi]),
                                                   0.01, psgd._tiny, **kw)
            for got, want in ((U, Ur), (V, Vr), (d, dr), (pre, pr)):
                if hi > lo:
                    e = np.linalg.norm(got.cpu().numpy().astype(np.float64) - want[lo:hi]) / np.linalg.norm(want)
                    assert e < 1e-5, ("fused", n, r, kw, e)
    # X-shape on mirrored chunk pairs, diagonal on plain chunks: only the max is exchanged
    n = 20_001
    c = cases.vec_case(7, n)
    lo, hi = partition.chunk_bounds(n, world, rank)
    q = dev(c["a"][lo:hi])
    psgd.update_precond_diag(q, dev(c["v"][lo:hi]), dev(c["h"][lo:hi]), 0.01)
    want = O.update_precond_diag(c["a"], c["v"], c["h"], 0.01)
    assert np.allclose(q.cpu().numpy(), want[lo:hi], rtol=1e-5)
    # X-shape on mirrored chunk pairs (odd n: the centre sits on the last rank), against the unsharded oracle
    n_x = 0
    for n in (20_001, 100_000):
        c = cases.vec_case(8 + n, n)
        idx = np.concatenate([np.arange(a, b) for a, b in partition.xmat_shard_slices(n, world, rank)]).astype(int)
        a_, b_ = dev(c["a"][idx]), dev(c["b"][idx])
        psgd.update_precond_Xmat(a_, b_, dev(c["v"][idx]), dev(c["h"][idx]), 0.01)
        ar, br = O.update_precond_Xmat(c["a"], c["b"], c["v"], c["h"], 0.01)
        assert np.allclose(a_.cpu().numpy(), ar[idx], rtol=1e-5, atol=1e-7), ("xmat a", n)
        assert np.allclose(b_.cpu().numpy(), br[idx], rtol=1e-5, atol=1e-7), ("xmat b", n)
        pre = psgd.precond_grad_Xmat(a_, b_, dev(c["g"][idx]))                      # no exchange: purely local
        assert np.allclose(pre.cpu().numpy(), O.precond_grad_Xmat(ar, br, c["g"])[idx], rtol=2e-5, atol=1e-6), ("xmat pre", n)
        n_x += 1
    if exchange == "peer":
        done = ctx.comm_status()                  # raises if any in-kernel wait timed out
        # per size: separate calls 2+1 exchanges (plain step) + 3+1 (balance step), fused call 2 + 3; + diag + X-shape
        assert done == 3 * (3 + 4 + 2 + 3) + 1 + n_x, done
        gpu_uvd_graph(rank, world, ctx)
        ctx.comm_detach()
    else:
        ctx.set_allreduce(None)


def gpu_uvd_graph(rank, world, ctx):
    """The sharded step replayed from CUDA graphs (peer exchange inside the graph) is bit-identical to the eager one."""
    import psgd_tf_b200 as psgd
    from psgd_tf_b200 import partition
    from psgd_tf_b200.graphs import UVdStepGraphs
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    n, r = 300_007, 10
    c = cases.uvd_case(77, n, r)
    lo, hi = partition.chunk_bounds(n, world, rank)
    ins = [tuple(dev(np.roll(c[k][lo:hi], s, 0)) for k in ("v", "h", "g")) for s in (0, 1)]
    Ue, Ve, de = dev(c["U"][lo:hi]), dev(c["V"][lo:hi]), dev(c["d"][lo:hi])
    Ug, Vg, dg = Ue.clone(), Ve.clone(), de.clone()
    gs = UVdStepGraphs(Ug, Vg, dg, 0.01)
    e0 = ctx.comm_status()
    for i in range(8):
        v, h, g = ins[i % 2]
        want = psgd.update_precond_and_grad_UVd(Ue, Ve, de, v, h, g, 0.01, psgd._tiny, balance=False, update_U=(i % 2 == 0))
        got = gs.step(v, h, g, False, i % 2 == 0)
        torch.cuda.synchronize()
        assert torch.equal(got, want) and torch.equal(Ug, Ue) and torch.equal(Vg, Ve) and torch.equal(dg, de), i
    assert gs.replays >= 5, gs.replays
    assert ctx.comm_status() - e0 == 8 * 2 * 2, (ctx.comm_status(), e0)   # 2 exchanges per fused update+apply, eager + graph


def gpu_kron(rank, world):
    import psgd_tf_b200 as psgd
    from psgd_tf_b200 import partition
    torch.cuda.set_device(rank)
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    for shapes in (cases.LENET_SHAPES, [(512, 512)] * 4):
        cs = [cases.kron_case(900 + i, "dense", "dense", M, N) for i, (M, N) in enumerate(shapes)]
        owned = partition.assign_layers([partition.kron_layer_cost(M, N) for M, N in shapes], world)
        mine = owned[rank]
        new = psgd.update_precond_kron_batched([dev(cs[i]["Ql"]) for i in mine], [dev(cs[i]["Qr"]) for i in mine],
                                               [dev(cs[i]["dX"]) for i in mine], [dev(cs[i]["dG"]) for i in mine], 0.01)
        pre = psgd.precond_grad_kron_batched([a for a, _ in new], [b for _, b in new], [dev(cs[i]["G"]) for i in mine])
        full = partition.all_gather_layers(pre, owned, shapes, rank)
        for i, c in enumerate(cs):
            ql, qr = O.update_precond_kron(c["Ql"], c["Qr"], c["dX"], c["dG"], 0.01)
            want = O.precond_grad_kron(ql, qr, c["G"])
            e = cases.rel_err(full[i].cpu().numpy(), want)
            assert e < 1e-5, (shapes[i], e)
        if len({tuple(sh) for sh in shapes}) != 1 or len({len(o) for o in owned}) != 1:
            continue
        # uniform stack: the in-place gather buffers (NCCL per slot; copy-engine pushes into IPC-mapped peer buffers), the
        # apply writing straight into them, three steps in a row so that both alternating buffers are used twice
        for cls, overlapped in ((partition.KronGatherBuffer, False), (partition.KronGatherBuffer, True),
                                (partition.KronPeerGather, False), (partition.KronPeerGather, True)):
            gb = cls(shapes, owned, rank, torch.device("cuda", rank))
            Ql, Qr = [a for a, _ in new], [b for _, b in new]
            for rep in range(3):
                G = [dev(cs[i]["G"] * (1.0 + rep)) for i in mine]
                outs = gb.local_outs()
                if overlapped:
                    for j in range(len(outs)):
                        psgd.precond_grad_kron_batched(Ql[j:j + 1], Qr[j:j + 1], G[j:j + 1], outs=outs[j:j + 1])
                        gb.push_slot(j) if cls is partition.KronPeerGather else gb.gather_slot(j)
                    full = gb.finish()
                else:
                    psgd.precond_grad_kron_batched(Ql, Qr, G, outs=outs)
                    full = gb.gather()
                for i, c in enumerate(cs):
                    ql, qr = O.update_precond_kron(c["Ql"], c["Qr"], c["dX"], c["dG"], 0.01)
                    want = O.precond_grad_kron(ql, qr, c["G"] * np.float32(1.0 + rep))
                    e = cases.rel_err(full[i].cpu().numpy(), want)
                    assert e < 1e-5, (cls.__name__, overlapped, rep, i, e)
            dist.barrier()


def main():
    mode = sys.argv[1]
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if mode == "cpu-protocol":
        dist.init_process_group("gloo")
        cpu_protocol(rank, world)
    else:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
        if mode == "gpu-uvd-peer":
            gpu_uvd(rank, world, exchange="peer")
        else:
            {"gpu-uvd": gpu_uvd, "gpu-kron": gpu_kron}[mode](rank, world)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print(f"MP_OK {mode} world={world}")


if __name__ == "__main__":
    main()
