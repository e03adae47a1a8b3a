"""world_size-W worker of tests/test_multirank_gloo.py (CPU, gloo): one process per rank, as deployed.

Each process owns ONE rank of the data-parallel job. The per-rank stage math comes from the oracle (there is no GPU
here), but everything between ranks goes through the product's host logic: libdprt's dprt_plan_exchange (send/recv
plan + termination rule) and host.exchange_host_records (the all-gather + send/recv protocol of dprt_exchange) over
a real torch.distributed process group. The result of the distributed run is compared, per rank, with the
single-process W-rank oracle.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch.distributed as dist
    rank, W = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    path_gen_mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    deque = len(sys.argv) > 2 and sys.argv[2] == "deque"       # the settled-deque migrate loop instead of the reference one
    dist.init_process_group("gloo")
    dprt = importlib.import_module("pg2024-data-parallel-ray-tracing_b200")
    D = dprt.ctypes_defs
    from oracle import oracle as O
    from helpers import assert_bits_equal, assert_records_equal

    w, h, bounces = 96, 54, 2
    chunks, mats, lights = dprt.scene.make_scene(W, 3000)
    cfg = dprt.make_config(w, h, spp=1, bounces=bounces, scene_size=W, proxy_mode=0, path_gen_mode=path_gen_mode)
    cam = dprt.scene.default_camera(w, h)

    def make_world():
        wd = O.World(cfg, W)
        for c in chunks:
            wd.add_object(c.index, c.desc(False), c.verts, c.normals, c.mats)
        wd.set_materials(mats); wd.set_lights(lights); wd.set_camera(cam)
        wd.reset_frame()
        return wd

    mine, whole = make_world(), make_world()
    whole.render_sample(0)                                  # the single-process W-rank answer

    # runSample (renderer.cpp:1457-1574) for THIS rank only, ranks coupled through the process group
    mine.begin_sample(0)
    mine.path_gen(rank)
    iters = sent = 0
    for bounce in range(bounces + 1):
        if deque:
            # only the travelling paths are traced / partitioned / sent; the settled block grows at both ends
            settled, n_lower = np.zeros(0, D.PATH_DTYPE), 0
            while True:
                mine.traverse(rank)                                        # the oracle's TraRay over the ACTIVE records only
                act = mine.download(rank, D.BUF_PATHS, mine.path_size(rank))
                buckets, row = dprt.host.partition_host_records_deque(act, rank, W, n_lower)
                active, n_lower, first, second, done = dprt.host.exchange_host_records_deque(buckets, row, rank, W, dist)
                settled = np.concatenate([first, settled, second])
                sent += int(row[W + 1] - (row[rank + 1] - row[rank]) - (row[W + 1] - row[W]))
                mine.upload(rank, D.BUF_PATHS, active)
                mine.set_path_size(rank, active.size)
                iters += 1
                if done:
                    break
            mine.upload(rank, D.BUF_PATHS, settled)
            mine.set_path_size(rank, settled.size)
            mine.shade(rank); mine.reset_nn(rank); mine.shadow_trace(rank); mine.frame_buffer_update(rank)
            continue
        while True:
            mine.traverse(rank)
            mine.partition(rank)
            off = mine.download(rank, D.BUF_TRANSFER_OFFSET, W + 1)
            tr = mine.download(rank, D.BUF_TRANSFER, int(off[W]))
            recv, done = dprt.host.exchange_host_records(tr, off, rank, W, dist)
            sent += int(off[W] - (off[rank + 1] - off[rank]))
            mine.upload(rank, D.BUF_PATHS, recv)
            mine.set_path_size(rank, recv.size)
            iters += 1
            if done:
                break
        mine.shade(rank); mine.reset_nn(rank); mine.shadow_trace(rank); mine.frame_buffer_update(rank)

    N, spc = w * h, cfg.shadowPathCount
    n = mine.path_size(rank)
    assert n == whole.path_size(rank), (rank, n, whole.path_size(rank))
    assert_records_equal(mine.download(rank, D.BUF_PATHS, n * (1 + spc)), whole.download(rank, D.BUF_PATHS, n * (1 + spc)), f"rank {rank} paths")
    assert_bits_equal(mine.download(rank, D.BUF_ENV, 3 * N), whole.download(rank, D.BUF_ENV, 3 * N), f"rank {rank} env")
    assert_bits_equal(mine.download(rank, D.BUF_DIRECT, 3 * N * spc), whole.download(rank, D.BUF_DIRECT, 3 * N * spc), f"rank {rank} direct")
    so = whole.stats(rank)
    assert iters == so["exchange_iters"] and sent == so["paths_sent_offrank"], (rank, iters, sent, so)
    import torch
    t = torch.tensor([sent], dtype=torch.int64)
    dist.all_reduce(t)
    assert int(t.item()) > 0, "nothing migrated"
    # image: MPI_Reduce(sum) to rank 0 (renderer.cpp:2052) == gloo reduce of the per-rank averages
    img = (mine.download(rank, D.BUF_DIRECT, 3 * N) + mine.download(rank, D.BUF_ENV, 3 * N)) / np.float32(cfg.spp)
    ti = torch.from_numpy(img.copy())
    dist.reduce(ti, 0)
    if rank == 0:
        ref = whole.image().reshape(-1)
        err = float(np.abs(ti.numpy() - ref).max() / max(1e-30, np.abs(ref).max()))
        assert err <= 1e-6, err
        print(f"GLOO_WORKER_OK world={W} iters={iters} migrated={int(t.item())}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
