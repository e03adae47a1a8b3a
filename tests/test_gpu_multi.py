"""Real multi-GPU parity (-m gpu, needs >= 2 devices; skipped on a 1-GPU box): one process per GPU over NCCL,
launched through torch.distributed.run exactly like bench.py, checked inside tests/mgpu_check.py."""
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
