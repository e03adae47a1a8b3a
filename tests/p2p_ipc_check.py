"""Two (or more) PROCESSES sharing ONE GPU, wired together with CUDA IPC handles exchanged over gloo -- no NCCL at all.

Exercises what the in-process group cannot: cudaIpcGetMemHandle / cudaIpcOpenMemHandle, stores into another process's
buffers, the flag protocol across contexts. (On one device the processes are time-sliced, so this is a functional
test, not a timing.) Launched by tests/test_gpu_multi.py as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P tests/p2p_ipc_check.py
Every rank compares its own path / env / direct buffers with the oracle bit for bit; rank 0 prints P2P_IPC_CHECK_OK.
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("DPRT_P2P_TIMEOUT_MS", "20000")


def main():
    import torch
    import torch.distributed as dist
    rank, W = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = int(os.environ.get("DPRT_TEST_DEVICE", "0"))          # all ranks on the same device
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo")
    dprt = importlib.import_module("pg2024-data-parallel-ray-tracing_b200")
    D = dprt.ctypes_defs
    from oracle import oracle as O
    from helpers import assert_bits_equal, assert_records_equal

    w, h, spp, bounces = 96, 54, 2, 2
    chunks, mats, lights = dprt.scene.make_scene(W, 4000)
    cfg = dprt.make_config(w, h, spp=spp, bounces=bounces, scene_size=W, proxy_mode=0, path_gen_mode=0, serial_stages=1)
    cam = dprt.scene.default_camera(w, h)
    R = dprt.Renderer(cfg, rank=rank, world=W, device=dev)          # no NCCL id: the host brings its own bootstrap
    world = O.World(cfg, W)
    for c in chunks:
        world.add_object(c.index, c.desc(False), c.verts, c.normals, c.mats)
        if c.node_id == rank:
            R.upload_chunk(c.index, c.desc(False), c.verts, c.normals, c.mats)
        else:
            R.upload_proxy(c.index, c.desc(True), None, None)
    for X in (R, world):
        X.set_materials(mats); X.set_lights(lights); X.set_camera(cam)

    handles = [None] * W
    dist.all_gather_object(handles, R.p2p_export())
    ok = torch.tensor([1 if R.p2p_connect(handles) else 0])
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    R.p2p_enable(bool(ok.item()))
    assert R.p2p_enabled, "CUDA IPC wiring failed"
    dist.barrier()

    R.reset_frame(); world.reset_frame()
    for s in range(spp):
        R.run_sample(s); world.render_sample(s)
    R.synchronize()
    N, spc = w * h, cfg.shadowPathCount
    n = R.path_size
    assert n == world.path_size(rank), (rank, n, world.path_size(rank))
    assert_records_equal(R.download(D.BUF_PATHS, n * (1 + spc)), world.download(rank, D.BUF_PATHS, n * (1 + spc)), f"rank {rank} paths")
    assert_bits_equal(R.download(D.BUF_ENV), world.download(rank, D.BUF_ENV, 3 * N), f"rank {rank} env")
    assert_bits_equal(R.download(D.BUF_DIRECT), world.download(rank, D.BUF_DIRECT, 3 * N * spc), f"rank {rank} direct")
    sg, so = R.stats(), world.stats(rank)
    for k in ("paths_sent_offrank", "exchange_iters"):
        assert sg[k] == so[k], (rank, k, sg[k], so[k])
    sent = torch.tensor([sg["paths_sent_offrank"]])
    dist.all_reduce(sent)
    assert int(sent.item()) > 0
    dist.barrier()
    R.close()
    if rank == 0:
        print(f"P2P_IPC_CHECK_OK world={W} migrated_paths={int(sent.item())}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
