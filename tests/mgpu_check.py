"""Multi-process, multi-GPU parity check of the NCCL data path (dprt_exchange, dprt_reduce_image).

Launched by tests/test_gpu_multi.py (or by hand) as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P tests/mgpu_check.py
One process per GPU, exactly like the product is deployed (renderer.cpp: one MPI rank per GPU). Every rank renders
its share through the C ABI with a real NCCL communicator; rank 0 compares the reduced image, and every rank its
own path buffer / lighting buffers / exchange statistics, with the oracle's W-rank world. Bit-exact (proxies off)
except the final image, whose ncclReduce sums ranks in an order NCCL chooses: compared at 1e-6 relative.
"""
import argparse
import importlib
import os
import sys

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")      # one hardware queue per stream (samples in flight)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tris", type=int, default=6000)
    ap.add_argument("--width", type=int, default=128)
    ap.add_argument("--height", type=int, default=72)
    ap.add_argument("--bounces", type=int, default=2)
    ap.add_argument("--path-gen-mode", type=int, default=0)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, W, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dprt = importlib.import_module("pg2024-data-parallel-ray-tracing_b200")
    D = dprt.ctypes_defs
    from oracle import oracle as O
    from helpers import assert_bits_equal, assert_records_equal

    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        t = torch.tensor(list(dprt.get_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    uid = bytes(t.cpu().tolist())

    chunks, mats, lights = dprt.scene.make_scene(W, args.tris)
    cfg = dprt.make_config(args.width, args.height, spp=2, bounces=args.bounces, scene_size=W, proxy_mode=0,
                           path_gen_mode=args.path_gen_mode)
    cam = dprt.scene.default_camera(args.width, args.height)
    R = dprt.Renderer(cfg, rank=rank, world=W, device=local, nccl_unique_id=uid)
    world = O.World(cfg, W)
    for c in chunks:
        world.add_object(c.index, c.desc(False), c.verts, c.normals, c.mats)
        if c.node_id == rank:
            R.upload_chunk(c.index, c.desc(False), c.verts, c.normals, c.mats)
        else:
            R.upload_proxy(c.index, c.desc(True), None, None)
    for X in (R, world):
        X.set_materials(mats); X.set_lights(lights); X.set_camera(cam)

    img = R.launch()
    img_o = world.launch()
    N, spc = args.width * args.height, cfg.shadowPathCount
    n = R.path_size
    assert n == world.path_size(rank), (rank, n, world.path_size(rank))
    assert_records_equal(R.download(D.BUF_PATHS, n * (1 + spc)), world.download(rank, D.BUF_PATHS, n * (1 + spc)), f"rank {rank} paths")
    assert_bits_equal(R.download(D.BUF_ENV), world.download(rank, D.BUF_ENV, 3 * N), f"rank {rank} env")
    assert_bits_equal(R.download(D.BUF_DIRECT), world.download(rank, D.BUF_DIRECT, 3 * N * spc), f"rank {rank} direct")
    sg, so = R.stats(), world.stats(rank)
    assert sg["rays_walked"] + sg["rays_shade_cached"] == so["rays_walked"], (sg, so)
    for k in ("rays_traverse", "rays_shade", "rays_shadow", "paths_sent_offrank", "exchange_iters"):
        assert sg[k] == so[k], (rank, k, sg[k], so[k])
    sent = torch.tensor([sg["paths_sent_offrank"]], dtype=torch.int64, device="cuda")
    dist.all_reduce(sent)
    assert int(sent.item()) > 0, "no path migrated: the scene does not exercise ncclSend/ncclRecv"
    # the same frame with two samples in flight per rank: a second context per rank on the same communicator, its own
    # mailboxes and receive buffers wired over the same peers (dprt_create_shared + dprt_adopt_scene)
    F = dprt.SamplesInFlight(R, 2)
    assert len(F.ctxs) == (2 if R.p2p_enabled else 1)
    img_f = F.launch()
    F.close()
    if rank == 0:
        assert np.isfinite(img).all() and img.max() > 0
        err = np.abs(img - img_o).max() / max(1e-30, float(np.abs(img_o).max()))
        assert err <= 1e-6, f"reduced image differs: {err}"
        err_f = np.abs(img_f - img_o).max() / max(1e-30, float(np.abs(img_o).max()))
        assert err_f <= 1e-6, f"image with two samples in flight differs: {err_f}"
        if W == 2:
            assert_bits_equal(img, img_o, "2-rank reduced image (a two-term fp32 sum is order-independent)")
        print(f"MGPU_CHECK_OK world={W} migrated_paths={int(sent.item())} image_rel_err={err:.2e} peer_memory_exchange={R.p2p_enabled}", flush=True)
    R.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
