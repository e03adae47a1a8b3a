"""Multi-process parity (-m gpu). test_nccl_*: one process per GPU over NCCL + peer memory, launched through
torch.distributed.run exactly like bench.py, checked inside tests/mgpu_check.py (needs >= 2 devices; skipped on a 1-GPU
box -- bench.py's own `parity` gate covers the driver's N > 1 runs). test_peer_memory_exchange_between_processes_*: W
processes sharing ONE device, wired with CUDA IPC over gloo (tests/p2p_ipc_check.py) -- runs on any GPU box."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("W,path_gen_mode", [(2, 0), (2, 1), (4, 1), (8, 0)])
def test_nccl_exchange_and_reduce_match_oracle(gpu_required, W, path_gen_mode):
    if _ngpu() < W:
        pytest.skip(f"needs {W} GPUs, {_ngpu()} visible")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={W}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_check.py"), "--path-gen-mode", str(path_gen_mode)]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "MGPU_CHECK_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]


@pytest.mark.parametrize("W", [2, 3])
def test_peer_memory_exchange_between_processes_on_one_device(gpu_required, W):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={W}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "p2p_ipc_check.py")]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=420)
    assert p.returncode == 0 and "P2P_IPC_CHECK_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
