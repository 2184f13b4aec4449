"""Peer copy bandwidth between two GPUs of one box: torch copy_ (same process), cudaMemcpyPeerAsync through cuda-python,
and torch copy_ into a tensor another process exported with CUDA IPC (what partition.KronPeerGather does)."""
import os, sys, time
import torch

def bw(fn, nbytes, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9

if len(sys.argv) > 1 and sys.argv[1] == "ipc":
    import torch.distributed as dist
    from torch.multiprocessing.reductions import reduce_tensor
    rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    buf = torch.zeros(64 << 20, device=f"cuda:{rank}")
    ex = [None, None]
    dist.all_gather_object(ex, reduce_tensor(buf))
    fn, a = ex[1 - rank]
    peer = fn(*a)
    src = torch.ones(64 << 20, device=f"cuda:{rank}")
    print(rank, "peer tensor device", peer.device, "can_access", torch.cuda.can_device_access_peer(rank, 1 - rank), flush=True)
    print(rank, "torch copy_ into IPC peer tensor: %.1f GB/s" % bw(lambda: peer.copy_(src, non_blocking=True), src.numel() * 4), flush=True)
    from cuda import cudart
    st = torch.cuda.current_stream().cuda_stream
    def raw():
        cudart.cudaMemcpyAsync(peer.data_ptr(), src.data_ptr(), src.numel() * 4, cudart.cudaMemcpyKind.cudaMemcpyDefault, st)
    print(rank, "cudaMemcpyAsync(default) into IPC peer pointer: %.1f GB/s" % bw(raw, src.numel() * 4), flush=True)
    def rawpeer():
        cudart.cudaMemcpyPeerAsync(peer.data_ptr(), 1 - rank, src.data_ptr(), rank, src.numel() * 4, st)
    print(rank, "cudaMemcpyPeerAsync into IPC peer pointer: %.1f GB/s" % bw(rawpeer, src.numel() * 4), flush=True)
    dist.barrier(); dist.destroy_process_group()
else:
    a = torch.ones(64 << 20, device="cuda:0"); b = torch.zeros(64 << 20, device="cuda:1")
    print("can_access", torch.cuda.can_device_access_peer(0, 1))
    torch.cuda.set_device(0)
    print("same process torch copy_ 0->1: %.1f GB/s" % bw(lambda: b.copy_(a, non_blocking=True), a.numel() * 4))
