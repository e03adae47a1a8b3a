#!/usr/bin/env python
"""BASELINE configs[3] -- the stated target: W chunk owners x T triangles (8 x 12.5 M = a 100 M-triangle scene), 1920x1080,
16 spp, bounces = 4, spc = 4, mc = 3, a TRAINED vis + depth proxy per remote chunk (7 per rank), one process per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \\
        profiles/run_config4.py [--tris 12500000] [--spp 16] [--epochs 8] [--train-rays 400000] [--inflight 3]

Every rank builds only its own chunk, trains the two proxies of that chunk on its own GPU (dprt_gen_train_data ->
proxy_train.train_chunk_proxies), the weight blobs are all-gathered, and the frame is rendered twice through the public call
sequence (reset_frame, spp samples in flight, accumulate, dprt_reduce_image): proxies ON (the reference's mode) and OFF
(sequential visiting only). One JSON line from rank 0: Mrays/s, samples/s, ms per sample, per-stage device times of a
profiled sample, all-to-all bytes and GB/s, image-reduce ms, MLP queries/s, and the HBM-resident traversal roofline (kernel
counters: the oracle cannot walk 100 M triangles here) -- device-timed, max over ranks."""
import argparse, importlib, json, os, sys, time
import numpy as np
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")     # one hardware queue per stream (samples in flight)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tris", type=int, default=12500000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--bounces", type=int, default=4)
    ap.add_argument("--epochs", type=int, default=8)
    ap.add_argument("--train-rays", type=int, default=400000)
    ap.add_argument("--inflight", type=int, default=3)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, W, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dprt = importlib.import_module("pg2024-data-parallel-ray-tracing_b200")
    import bench
    if W > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    def fresh_uid():
        """One NCCL unique id per communicator (a context creates its own): rank 0 draws it, everybody gets it."""
        if W == 1:
            return None
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t = torch.tensor(list(dprt.get_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())
    w, h = args.width, args.height
    N = w * h
    cam = dprt.scene.default_camera(w, h)
    t0 = time.time()
    chunks, mats, lights = dprt.scene.make_scene(W, args.tris, layout="slabs", camera=cam, only=[rank])
    mine = chunks[rank]
    t_scene = time.time() - t0
    pk = bench.peaks()

    def build(proxy_mode, blobs):
        cfg = dprt.make_config(w, h, spp=args.spp, bounces=args.bounces, scene_size=W, proxy_mode=proxy_mode, path_gen_mode=1 if W > 1 else 0, mlp_dtype=1)
        R = dprt.Renderer(cfg, rank=rank, world=W, device=local, nccl_unique_id=fresh_uid())
        for c in chunks:
            if c.index == rank:
                R.upload_chunk(c.index, c.desc(False), c.verts, c.normals, c.mats)
            else:
                vb, db = blobs.get(c.index, (None, None))
                R.upload_proxy(c.index, c.desc(True), vb, db)
        R.set_materials(mats); R.set_lights(lights); R.set_camera(cam)
        return R

    def allmax(x):
        if W == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    def allsum(x):
        if W == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.SUM); return float(t.item())

    def barrier():
        torch.cuda.synchronize()
        if W > 1:
            dist.barrier()

    # ---- proxies: every rank trains the networks of ITS chunk, the blobs travel to everybody -------------------------
    t0 = time.time()
    R0 = build(0, {})
    t_upload = time.time() - t0
    t0 = time.time()
    info = None
    blobs = {}
    if W > 1:
        vb, db, info = dprt.proxy_train.train_chunk_proxies(lambda rays: R0.gen_train_data(mine.index, rays), mine.aabb_min, mine.aabb_max,
                                                            n_rays=args.train_rays, epochs=args.epochs, seed=mine.index, device="cuda")
        every = [None] * W
        dist.all_gather_object(every, (mine.index, vb, db))
        blobs = {k: (a, b) for k, a, b in every}
    t_train = time.time() - t0

    def render(R, label):
        """spp samples in flight through the public call sequence; device-timed per context, max over contexts and ranks."""
        F = dprt.SamplesInFlight(R, args.inflight)
        F.reset_frame(); F.run_samples(0, 2 * len(F.ctxs))                 # warm-up
        for X in F.ctxs:
            X.synchronize()
        barrier()
        for X in F.ctxs:
            X.reset_stats()
        F.reset_frame()
        for X in F.ctxs:
            X.timer_start()
        F.run_samples(0, args.spp)
        ms = max(X.timer_stop() for X in F.ctxs)
        t1 = time.perf_counter()
        F.accumulate()
        img = R.reduce_image(0)
        reduce_wall_ms = (time.perf_counter() - t1) * 1e3
        barrier()
        st = F.stats()
        ms = allmax(ms)
        rays = allsum(float(st["rays_walked"]))
        # one profiled sample (serial stages, one context) for the per-stage device times and the traversal counters
        R.reset_stats(); R.stage_profile(True); R.enable_counters(True)
        R.run_sample(args.spp); R.synchronize()
        _ = R.reduce_image(0)
        stage = R.stage_times(); cnt = R.counters(); sst = R.stats()
        R.stage_profile(False); R.enable_counters(False)
        walked = {"traverse": sst["walked_traverse"], "shade": sst["walked_shade"], "shadow_trace": sst["walked_shadow"], "secondary_trace": sst["walked_secondary"]}
        stages = {}
        for name, (t_ms, ln) in stage.items():
            if not ln:
                continue
            e = {"ms": t_ms, "launches": ln}
            if name in walked and walked[name]:
                nodes, tris = cnt[name]
                alg = walked[name] * bench.RECORD_BYTES[name] + nodes * bench.NODE_BYTES + tris * bench.TRI_BYTES
                e.update(rays=walked[name], nodes_per_ray=nodes / walked[name], tris_per_ray=tris / walked[name], GBps=alg / (t_ms * 1e-3) / 1e9,
                         frac_of_hbm=alg / (t_ms * 1e-3) / 1e9 / pk["hbm"])
            stages[name] = e
        out = {"label": label, "Mrays_per_s": rays / (ms * 1e-3) / 1e6, "samples_per_s": N * args.spp / (ms * 1e-3), "ms_per_sample": ms / args.spp,
               "frame_ms": ms, "rays_per_sample": rays / args.spp, "samples_in_flight": len(F.ctxs),
               "alltoall": {"bytes_per_sample": allsum(float(st["bytes_alltoall"])) / args.spp, "exchange_iters_per_sample": st["exchange_iters"] / args.spp,
                            "GBps_per_gpu_in_exchange_kernels": sst["bytes_alltoall"] / max(1e-9, (stage["exchange"][0] + stage["partition"][0]) * 1e-3) / 1e9},
               "image_reduce": {"device_ms": stage["image"][0] / max(1, stage["image"][1]), "accumulate_plus_reduce_wall_ms": reduce_wall_ms, "bytes": N * 12},
               "nn_queries_per_sample": allsum(float(st["nn_queries"])) / args.spp,
               "mlp": None if not stage["proxy_mlp"][1] else {"ms_per_sample": stage["proxy_mlp"][0], "launches_per_sample": stage["proxy_mlp"][1],
                       "queries_per_sample_this_rank": sst["nn_queries"], "TFLOPs": sst["nn_queries"] * bench.MLP_FLOP_PER_QUERY / (stage["proxy_mlp"][0] * 1e-3) / 1e12,
                       "frac_of_tensor_peak": sst["nn_queries"] * bench.MLP_FLOP_PER_QUERY / (stage["proxy_mlp"][0] * 1e-3) / 1e12 / pk["tensor_burst"]},
               "stages_profiled_sample_rank0": stages, "exchange_data_plane": "peer memory (CUDA IPC)" if R.p2p_enabled else ("nccl" if W > 1 else "single rank")}
        F.close()
        return img, out

    results = []
    img_off, r_off = render(R0, "proxies off (sequential visiting)")
    results.append(r_off)
    img_on = None
    if W > 1:
        R0.close()
        R1 = build(1, blobs)
        img_on, r_on = render(R1, "proxies on (trained vis + depth network per remote chunk)")
        results.append(r_on)
        R1.close()
    else:
        R0.close()
    if rank == 0:
        line = {"config": f"BASELINE configs[3]: {W} chunks x {mine.ntris} triangles ({W * mine.ntris / 1e6:.1f} M), {w}x{h}, {args.spp} spp, bounces={args.bounces}, spc=4, mc=3",
                "n_gpus": W, "scene_seconds": t_scene, "upload_seconds_incl_bvh8_build": t_upload, "proxy_training": {"seconds": t_train, "rays": args.train_rays, "epochs": args.epochs,
                "this_rank": None if info is None else {"hit_fraction": info["hit_fraction"], "vis_test_loss_first_last": [info["vis_test_loss"][0], info["vis_test_loss"][-1]],
                                                         "depth_test_loss_first_last": [info["depth_test_loss"][0], info["depth_test_loss"][-1]]}},
                "peaks": pk, "runs": results}
        if img_on is not None:
            line["image_proxy_on_vs_off"] = {"rel_mse": float(np.mean((img_on - img_off) ** 2 / (img_off ** 2 + 1e-2))), "mean_on": float(img_on.mean()), "mean_off": float(img_off.mean()),
                                              "note": "off = remote chunks do not occlude shadow rays (no proxies to ask); on = the proxies answer for them"}
        print(json.dumps(line), flush=True)
        if args.out:
            open(args.out, "w").write(json.dumps(line, indent=1))
    if W > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
