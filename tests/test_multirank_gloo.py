"""N>1 host path on CPU: world_size-2 (and 4) gloo process groups, one process per rank (not gpu)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from helpers import dprt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(nproc, script_args, timeout=600):
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port())] + script_args
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)


@pytest.mark.parametrize("W,path_gen_mode", [(2, 0), (2, 1), (4, 1)])
def test_exchange_protocol_over_gloo_matches_single_process_oracle(W, path_gen_mode):
    p = _torchrun(W, [os.path.join(ROOT, "tests", "gloo_worker.py"), str(path_gen_mode)])
    assert p.returncode == 0 and "GLOO_WORKER_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]


def test_plan_exchange_known_answers():
    # rank s row = exclusive offsets of its segments per destination
    M = np.array([[0, 5, 7, 7], [0, 0, 3, 4], [0, 2, 2, 9]], np.int32)
    sc, ro, rc, tot, loc = dprt.plan_exchange(M, 1)
    assert sc.tolist() == [0, 3, 1] and rc.tolist() == [2, 3, 0] and ro.tolist() == [0, 2, 5, 5] and tot == 5 and not loc
    sc, ro, rc, tot, loc = dprt.plan_exchange(M, 2)
    assert sc.tolist() == [2, 0, 7] and rc.tolist() == [0, 1, 7] and ro.tolist() == [0, 0, 1, 8] and tot == 8
    # all traffic local -> termination (renderer.cpp:1292-1298)
    L = np.array([[0, 4, 4], [0, 0, 6]], np.int32)
    assert dprt.plan_exchange(L, 0)[3:] == (4, True) and dprt.plan_exchange(L, 1)[3:] == (6, True)
    # empty world
    Z = np.zeros((3, 4), np.int32)
    assert dprt.plan_exchange(Z, 0)[3:] == (0, True)
    # malformed rows are rejected, not trusted
    with pytest.raises(dprt.DprtError):
        dprt.plan_exchange(np.array([[0, 3, 2], [0, 0, 0]], np.int32), 0)
    with pytest.raises(dprt.DprtError):
        dprt.plan_exchange(np.array([[1, 3, 4], [0, 0, 0]], np.int32), 0)
    with pytest.raises(dprt.DprtError):
        dprt.plan_exchange(np.zeros((2, 2), np.int32), 0)


def test_bench_reference_arm_under_torchrun_prints_one_line():
    """bench.py --impl reference at N=2: rank 0 alone works and prints the JSON line, rank 1 exits 0."""
    p = _torchrun(2, [os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                      "--tris", "2000", "--ref-scale", "24"])
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 2 and j["unit"] == "Mrays/s" and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "port" and j["e2e"]["h2d_bytes_per_step"] == 0
