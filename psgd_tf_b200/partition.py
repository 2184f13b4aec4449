"""Multi-GPU partitioners for one 8 x B200 box: one process per GPU, ``torch.distributed`` (NCCL over NVLink 5 /
NVSwitch) as the plumbing.

* **UVd / diagonal / X-shape** shard the flattened parameter vector by contiguous chunk.  Every big operand is indexed
  by parameter row, so the only exchange is an all-reduce of the r x r / r-length partial sums (float64, a few hundred
  values) and of one max per phase.  The C library calls back into :func:`install_allreduce`'s hook between its
  kernels, on its own stream, so the collective is stream-ordered with the sweeps and nothing syncs the host.
* **Kron stacks** shard layer-wise: layers are independent (mnist_with_lenet5.py:51-53), so each rank updates and
  applies only the layers it owns and the preconditioned gradients are all-gathered afterwards.
* **Dense full-matrix** preconditioners are replicas only (n < 1e4; not worth sharding).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


class _DevBuf:
    """A raw device pointer dressed up for ``torch.as_tensor`` (zero-copy via __cuda_array_interface__)."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": None}


def install_allreduce(ctx, group=None) -> None:
    """Route the library's cross-rank reductions through ``torch.distributed.all_reduce`` on ``group``."""
    import torch.distributed as dist

    cache = {}

    def hook(ptr: int, count: int, op: int, stream: int) -> int:
        key = (ptr, count, op)
        t = cache.get(key)
        if t is None:
            t = torch.as_tensor(_DevBuf(ptr, count, "<f8" if op == 0 else "<f4"), device=torch.device("cuda", ctx.device))
            cache[key] = t
        # the library's stream is torch's current stream (psgd.get_context sets it), so the collective is ordered
        # after the partial-sum kernel and before the kernel that consumes the result
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == 0 else dist.ReduceOp.MAX, group=group)
        return 0

    ctx.set_allreduce(hook)


def install_peer_exchange(ctx, group=None) -> None:
    """Attach the library's peer-memory exchange (csrc/comm.cu) across the ranks of ``group`` (one process per GPU
    on ONE NVSwitch box): every rank exports a small device slab through CUDA IPC, the 64-byte handles are
    all-gathered with torch.distributed (host plumbing, once), and from then on each cross-rank reduction of the
    sharded UVd / diagonal / X-shape paths is one tiny kernel that stores the partials straight into the peers'
    slabs over NVLink and reduces in fixed rank order -- no host callback, no NCCL launch, CUDA-graph capturable,
    bit-identical results on every rank.  Takes precedence over :func:`install_allreduce` while attached."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    handles = [None] * world
    dist.all_gather_object(handles, ctx.comm_export(), group=group)
    err = None
    try:
        ctx.comm_attach(rank, world, handles)
    except Exception as e:              # keep the collective sequence identical on every rank
        err = e
    ok = torch.tensor([0.0 if err else 1.0], device=torch.device("cuda", ctx.device))
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)   # also the barrier: every rank has mapped every slab
    if ok.item() < 1:
        ctx.comm_detach()
        raise RuntimeError(f"peer-memory exchange unavailable on this box: {err or 'CUDA IPC failed on another rank'}")


def install_exchange(ctx, group=None) -> str:
    """Peer-memory exchange when CUDA IPC between the ranks works, else the torch.distributed hook.  Returns which."""
    try:
        install_peer_exchange(ctx, group)
        return "peer-memory"
    except RuntimeError as e:
        print(f"psgd_tf_b200: {e}; using the torch.distributed all-reduce hook")
    install_allreduce(ctx, group)
    return "all-reduce hook"


def chunk_bounds(n: int, world: int, rank: int, align: int = 256) -> Tuple[int, int]:
    """Contiguous row chunk [lo, hi) of rank ``rank``; chunk starts are multiples of ``align`` rows so every shard's
    U/V/d base pointers stay 16-byte aligned for the TMA bulk copies."""
    per = -(-n // world)
    per = -(-per // align) * align
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def mirrored_chunk_bounds(n: int, world: int, rank: int, align: int = 256) -> Tuple[Tuple[int, int], Tuple[int, int]]:
    """X-shape sharding: element i only ever meets n-1-i, so rank k owns a chunk [lo, hi) of the first n // 2 elements
    and its mirror image [n-hi, n-lo) in the second half; no data exchange is needed, only the max all-reduce.  For odd
    n the centre element n // 2 (its own mirror image) belongs to neither half: see :func:`xmat_shard_slices`."""
    lo, hi = chunk_bounds(n // 2, world, rank, align)
    return (lo, hi), (n - hi, n - lo)


def xmat_shard_slices(n: int, world: int, rank: int, align: int = 256) -> List[Tuple[int, int]]:
    """Index ranges of the global vector that make up rank ``rank``'s LOCAL X-shape problem, in local order:
    ``[lo, hi)``, the centre element when n is odd and this rank owns it, then ``[n-hi, n-lo)``.  The local vector is
    itself a flip-symmetric problem (local element j meets local element n_local-1-j exactly when their global indices
    are mirror images), so each rank calls ``update_precond_Xmat`` / ``precond_grad_Xmat`` on its concatenated shard
    unchanged and the library only exchanges the max of |nabla| (csrc/elementwise.cu).  The centre goes to the last
    rank with a non-empty chunk (rank 0 when n < 2), whose local length is then odd with the centre in the middle, which
    is where the kernel zeroes nabla_b."""
    (lo, hi), (mlo, mhi) = mirrored_chunk_bounds(n, world, rank, align)
    parts = [(lo, hi)]
    if n % 2 == 1:
        owners = [k for k in range(world) if chunk_bounds(n // 2, world, k, align)[1] > chunk_bounds(n // 2, world, k, align)[0]]
        if rank == (owners[-1] if owners else 0):
            parts.append((n // 2, n // 2 + 1))
    parts.append((mlo, mhi))
    return parts


def xmat_shard(x: torch.Tensor, world: int, rank: int, align: int = 256) -> torch.Tensor:
    """This rank's local X-shape vector (a copy) cut from the global 1-D tensor ``x``."""
    return torch.cat([x[a:b] for a, b in xmat_shard_slices(x.numel(), world, rank, align)])


def assign_layers(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time bin packing of layers onto ranks by cost (e.g. the 26 n^3 flop count of a dense-dense
    layer).  Uniform stacks reduce to round-robin blocks."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world
    owned: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        k = min(range(world), key=lambda j: (loads[j], j))
        owned[k].append(i)
        loads[k] += costs[i]
    for o in owned:
        o.sort()
    return owned


def kron_layer_cost(M: int, N: int, kind_l: int = 0, kind_r: int = 0) -> float:
    """Dense flop count of one update+apply; structured factors cost O(MN)."""
    c = 20.0 * M * N
    if kind_l == 0:
        c += 2.0 * M * M * N * 4 + 2.0 * M ** 3
    if kind_r == 0:
        c += 2.0 * M * N * N * 4 + 2.0 * N ** 3
    return c


class KronGatherBuffer:
    """In-place all-gather of a UNIFORM layer-sharded Kron stack (every layer the same shape, every rank the same number
    of layers ``per``).  One ``[per, world, M, N]`` buffer per rank, slot-major: ``local_outs()`` are views of this
    rank's entries ``[j, rank]``, to be passed as ``outs=`` to :func:`psgd_tf_b200.precond_grad_kron_batched` so the
    apply's last product writes its result where NCCL sends it from -- no ``torch.stack``, no staging copy.

    ``gather_slot(j)`` issues the in-place ``all_gather_into_tensor`` of slot ``j`` (the j-th layer of every rank,
    contiguous in the buffer) asynchronously, as soon as the apply of that layer has been enqueued: NCCL's stream waits
    for the compute stream at that point only, so the transfer of slot j runs while slot j+1 is being applied.
    ``finish()`` joins the compute stream with all of them and returns every layer's result.  ``gather()`` = all slots,
    then ``finish()``.  Two buffers alternate so that the result of step t stays valid while step t+1 is being computed."""

    def __init__(self, shapes, owned, rank: int, device, nbuf: int = 2):
        if len({tuple(s) for s in shapes}) != 1 or len({len(o) for o in owned}) != 1:
            raise ValueError("KronGatherBuffer: needs a uniform stack (use all_gather_layers for ragged ones)")
        self.owned, self.rank, self.per, self.world = owned, rank, len(owned[0]), len(owned)
        self.shape = tuple(shapes[0])
        self.bufs = [torch.empty((self.per, self.world) + self.shape, device=device, dtype=torch.float32) for _ in range(nbuf)]
        self.cur = 0
        self._works = []

    def local_outs(self):
        self.cur = (self.cur + 1) % len(self.bufs)
        b = self.bufs[self.cur]
        return [b[j, self.rank] for j in range(self.per)]

    def gather_slot(self, j: int, group=None):
        import torch.distributed as dist
        b = self.bufs[self.cur]
        # flat views: concatenation along dim 0 is the one layout every backend accepts (gloo rejects the stacked form)
        self._works.append(dist.all_gather_into_tensor(b[j].view(-1), b[j, self.rank].view(-1), group=group, async_op=True))

    def finish(self):
        for w in self._works:
            w.wait()
        self._works = []
        b = self.bufs[self.cur]
        full = [None] * (self.world * self.per)
        for k, layer_ids in enumerate(self.owned):
            for j, li in enumerate(layer_ids):
                full[li] = b[j, k]
        return full

    def gather(self, group=None):
        for j in range(self.per):
            self.gather_slot(j, group)
        return self.finish()


class KronPeerGather:
    """All-gather of a uniform layer-sharded Kron stack by COPY ENGINES over NVLink: no SMs, so the transfer of layer j
    really runs under the apply of layer j+1 (an NCCL all-gather needs SMs that the persistent GEMM kernels hold, and a
    GEMM launch that finds some of its SMs taken ends late by the collective's duration).

    One process per GPU on one NVSwitch box.  Every rank allocates ``[world, per, M, N]`` buffers (rank-major: rank k's
    layers at ``[k]``), exports them with CUDA IPC (``torch.multiprocessing.reductions.reduce_tensor``; handles exchanged
    once through ``all_gather_object``) and maps every peer's buffers.  ``local_outs()`` are this rank's entries of its
    own buffer (``outs=`` of the batched apply); ``push_slot(j)`` enqueues, on side streams that wait for the compute
    stream at that point only, one contiguous device-to-device copy of layer j into every peer's buffer;
    ``finish()`` closes the step with a tiny all-reduce enqueued behind the copies -- no rank gets past it before every
    rank's copies have landed -- joins the compute stream and returns every layer's result.  Two buffers alternate: a peer
    that is one step ahead writes into the other one."""

    def __init__(self, shapes, owned, rank: int, device, nbuf: int = 2, group=None, push_streams: int = 2):
        import torch.distributed as dist
        from torch.multiprocessing.reductions import reduce_tensor
        if len({tuple(s) for s in shapes}) != 1 or len({len(o) for o in owned}) != 1:
            raise ValueError("KronPeerGather: needs a uniform stack (use all_gather_layers for ragged ones)")
        self.owned, self.rank, self.per, self.world, self.group = owned, rank, len(owned[0]), len(owned), group
        self.shape = tuple(shapes[0])
        self.bufs = [torch.empty((self.world, self.per) + self.shape, device=device, dtype=torch.float32) for _ in range(nbuf)]
        self.peers = []                                   # peers[b][k]: rank k's buffer b, mapped here
        for b in self.bufs:
            exported = [None] * self.world
            dist.all_gather_object(exported, reduce_tensor(b), group=group)
            self.peers.append([b if k == rank else fn(*a) for k, (fn, a) in enumerate(exported)])
        self.streams = [torch.cuda.Stream(device=device) for _ in range(push_streams)]
        self.flag = torch.zeros(1, device=device)
        self.cur = 0

    def local_outs(self):
        self.cur = (self.cur + 1) % len(self.bufs)
        b = self.bufs[self.cur]
        return [b[self.rank, j] for j in range(self.per)]

    def push_slot(self, j: int):
        ready = torch.cuda.current_stream().record_event()
        src = self.bufs[self.cur][self.rank, j]
        for s in self.streams:
            s.wait_event(ready)
        for i in range(self.world - 1):                   # staggered: not every rank starts on the same destination
            dst = self.peers[self.cur][(self.rank + 1 + i) % self.world]
            with torch.cuda.stream(self.streams[i % len(self.streams)]):
                dst[self.rank, j].copy_(src, non_blocking=True)

    def finish(self):
        import torch.distributed as dist
        s0 = self.streams[0]
        for s in self.streams[1:]:
            s0.wait_event(s.record_event())
        with torch.cuda.stream(s0):
            dist.all_reduce(self.flag, group=self.group)            # behind this rank's copies; completes when all ranks joined
            done = s0.record_event()
        torch.cuda.current_stream().wait_event(done)
        b = self.bufs[self.cur]
        full = [None] * (self.world * self.per)
        for k, layer_ids in enumerate(self.owned):
            for j, li in enumerate(layer_ids):
                full[li] = b[k, j]
        return full

    def gather(self):
        for j in range(self.per):
            self.push_slot(j)
        return self.finish()


def all_gather_layers(outs_local: Sequence[torch.Tensor], owned: List[List[int]], shapes: Sequence[Tuple[int, int]],
                      rank: int, group=None) -> List[torch.Tensor]:
    """All-gather of preconditioned gradients for a layer-sharded Kron stack: every rank ends up with every layer's
    result.  Ragged layer sizes are handled by one broadcast per layer from its owner (grouped into one NCCL group
    call by torch's coalescing manager when available)."""
    import torch.distributed as dist

    dev = outs_local[0].device if outs_local else torch.device("cuda", torch.cuda.current_device())
    full: List[torch.Tensor] = [None] * len(shapes)  # type: ignore
    for k, layer_ids in enumerate(owned):
        for j, li in enumerate(layer_ids):
            full[li] = outs_local[j] if k == rank else torch.empty(shapes[li], device=dev, dtype=torch.float32)
    uniform = len({tuple(s) for s in shapes}) == 1 and len({len(o) for o in owned}) == 1
    if uniform:
        # uniform stack (24 x 4096^2): one all_gather_into_tensor of the stacked local results
        per = len(owned[0])
        local = torch.stack([full[li] for li in owned[rank]]) if per else torch.empty(0, device=dev)
        gathered = torch.empty((len(owned) * per,) + tuple(local.shape[1:]), device=dev, dtype=torch.float32)
        dist.all_gather_into_tensor(gathered, local, group=group)          # rank k's layers land at rows [k*per, (k+1)*per)
        for k, layer_ids in enumerate(owned):
            for j, li in enumerate(layer_ids):
                full[li] = gathered[k * per + j]
        return full
    works = []
    for k, layer_ids in enumerate(owned):
        for li in layer_ids:
            works.append(dist.broadcast(full[li], src=k, group=group, async_op=True))
    for w in works:
        w.wait()
    return full
