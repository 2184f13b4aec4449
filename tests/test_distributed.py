"""World-size-2 tests of the multi-GPU host logic (gloo, CPU) and, on a multi-GPU box, of the real sharded CUDA paths."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(mode, world, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mp_worker.py"), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0 and f"MP_OK {mode}" in r.stdout, f"stdout:\n{r.stdout[-3000:]}\nstderr:\n{r.stderr[-6000:]}"


@pytest.mark.parametrize("world", [2, 3])
def test_partition_protocol_gloo(world):
    _launch("cpu-protocol", world)


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["gpu-uvd", "gpu-uvd-peer", "gpu-kron"])
def test_sharded_cuda_paths_two_gpus(mode):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    _launch(mode, 2)
