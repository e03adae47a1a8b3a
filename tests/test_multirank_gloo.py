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


@pytest.mark.parametrize("W,path_gen_mode", [(2, 1), (4, 0)])
def test_settled_deque_protocol_over_gloo_matches_single_process_oracle(W, path_gen_mode):
    """The migrate loop of cfg.referenceMigrate = 0 (W+1-bucket partition, dprt_plan_exchange_deque, settled block growing
    at both ends) with one process per rank over gloo: same final buffers as the oracle's iterate-over-everything loop."""
    p = _torchrun(W, [os.path.join(ROOT, "tests", "gloo_worker.py"), str(path_gen_mode), "deque"])
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


# ---- settled deque: the host-side model of the migrate loop ---------------------------------------------------------
def _simulate_migrate(W, itineraries, start_rank, deque):
    """Paths with fixed itineraries (the ranks a path asks for after each trace; the last one is where it settles and is
    traced once more as a rider, like a path returning to the rank of its closest hit). Returns the final buffer of every
    rank (path ids in order) and the number of exchange iterations. deque=False: every iteration partitions every path
    (Work_Efficient_Scan + MPI_Alltoallv, renderer.cpp:1212-1318); deque=True: only travelling paths, planned by
    dprt_plan_exchange_deque."""
    step = {p: 0 for p in range(len(itineraries))}

    def target(p, here):
        it = itineraries[p]
        if step[p] < len(it):
            t = it[step[p]]
            step[p] += 1
            return t
        return here                                           # settled (or a rider): stays

    bufs = [[p for p in range(len(itineraries)) if start_rank[p] == r] for r in range(W)]
    settled = [[] for _ in range(W)]
    nl = [0] * W
    iters = 0
    while True:
        iters += 1
        if not deque:
            seg = [[[] for _ in range(W)] for _ in range(W)]      # seg[s][d]
            for s in range(W):
                for p in bufs[s]:
                    seg[s][target(p, s)].append(p)
            off = sum(len(seg[s][d]) for s in range(W) for d in range(W) if s != d)
            bufs = [[p for s in range(W) for p in seg[s][d]] for d in range(W)]
            if off == 0:
                return bufs, iters
        else:
            rows = np.zeros((W, W + 2), np.int32)
            buckets = [[[] for _ in range(W + 1)] for _ in range(W)]
            for s in range(W):
                for i, p in enumerate(bufs[s]):
                    t = target(p, s)
                    buckets[s][W if (t == s and i >= nl[s]) else t].append(p)
                rows[s, 1:] = np.cumsum([len(b) for b in buckets[s]])
            new_bufs = []
            done = None
            plans = [dprt.plan_exchange_deque(rows, r) for r in range(W)]
            for d in range(W):
                plan = plans[d]
                flat = [p for b in buckets[d] for p in b]
                oL, cL, oR, cR = plan["piece"]
                settled[d] = flat[oL:oL + cL] + settled[d] + flat[oR:oR + cR]
                recv = []
                for s in range(W):
                    if s != d:
                        assert plan["recv_count"][s] == len(buckets[s][d])
                        recv += buckets[s][d]
                assert plan["new_active"] == len(recv) and plan["new_nl"] == sum(len(buckets[s][d]) for s in range(d))
                # one-sided placement (what the peer-memory exchange does): every sender writes its bucket at its dst_offset
                placed = [None] * len(recv)
                for s in range(W):
                    if s != d:
                        o = int(plans[s]["dst_offset"][d])
                        placed[o:o + len(buckets[s][d])] = buckets[s][d]
                assert placed == recv
                new_bufs.append(recv)
                nl[d] = plan["new_nl"]
                done = plan["all_local"] if done is None else (done and plan["all_local"])
            bufs = new_bufs
            if done:
                assert all(len(b) == 0 for b in bufs)
                return settled, iters


@pytest.mark.parametrize("W,npaths,seed", [(2, 200, 0), (3, 500, 1), (8, 3000, 2), (16, 2000, 3), (31, 1500, 4)])
def test_settled_deque_reproduces_the_reference_buffers(W, npaths, seed):
    """Property behind cfg.referenceMigrate = 0: for arbitrary itineraries the deque schedule leaves every
    rank with exactly the records, in exactly the order, of the reference's iterate-over-everything schedule."""
    rng = np.random.default_rng(seed)
    its, start = [], []
    for _ in range(npaths):
        hops = int(rng.integers(0, min(W, 6)))
        order = rng.permutation(W)[:hops + 1]
        start.append(int(order[0]))
        it = [int(x) for x in order[1:]]
        if hops and rng.random() < 0.5:
            it.append(int(order[rng.integers(0, hops + 1)]))   # return to an earlier rank (the one that owns the closest hit)
        its.append(it)
    ref, it_ref = _simulate_migrate(W, its, start, deque=False)
    fast, it_fast = _simulate_migrate(W, its, start, deque=True)
    assert it_ref == it_fast
    assert ref == fast
    assert sum(len(b) for b in ref) == npaths


def test_p2p_plan_matches_deque_plan(tmp_path):
    """The plan the peer-memory exchange computes on the device from the gathered histograms (p2p_plan_numbers, compiled
    here for the host) equals dprt_plan_exchange_deque -- what the NCCL fallback executes -- on random count matrices."""
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    exe = str(tmp_path / "p2p_plan_check")
    csrc = os.path.join(ROOT, "pg2024-data-parallel-ray-tracing_b200", "csrc")
    cc = subprocess.run(["g++", "-O1", "-std=c++17", "-I", cuda_inc, "-I", csrc, "-I", os.path.join(ROOT, "include"),
                         os.path.join(ROOT, "tests", "p2p_plan_check.cpp"), "-o", exe], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-3000:]
    rng = np.random.default_rng(5)
    for W in (2, 3, 8, 31):
        counts = rng.integers(0, 50, (W, W + 1)) * (rng.random((W, W + 1)) < 0.6)
        if W == 3:
            counts[~np.eye(W, W + 1, dtype=bool)] = 0            # nothing crosses ranks: termination
            counts[:, W] = 7
        rows = np.zeros((W, W + 2), np.int32)
        rows[:, 1:] = np.cumsum(counts, axis=1)
        for me in {0, W // 2, W - 1}:
            inp = f"{W} {me}\n" + " ".join(str(int(v)) for v in counts.reshape(-1)) + "\n"
            out = subprocess.run([exe], input=inp, capture_output=True, text=True)
            assert out.returncode == 0
            v = [int(x) for x in out.stdout.split()]
            ref = dprt.plan_exchange_deque(rows, me)
            assert v[0:W] == ref["dst_offset"].tolist()
            cL, cR, new_nl, new_active, all_local, sent, total, max_arr = v[W:W + 8]
            assert (cL, cR) == (ref["piece"][1], ref["piece"][3])
            assert new_nl == ref["new_nl"] and new_active == ref["new_active"] and bool(all_local) == ref["all_local"]
            assert sent == int(ref["send_count"].sum()) and total == int(counts[me].sum())
            arrivals = [int(counts[:, d].sum() - counts[d, d]) for d in range(W)]
            assert max_arr == max(arrivals)
