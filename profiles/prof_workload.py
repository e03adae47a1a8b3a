#!/usr/bin/env python
"""Single-GPU workload that launches every kernel of the hot path, for ncu captures (never a bench number):
a 2-chunk world as a RankGroup on one GPU with neural proxies on (all trace modes, both partition kernels, the
epilogues, the fused proxy MLP inside the pipeline), then a 2^20-query proxy-MLP batch.
usage: python profiles/prof_workload.py [width height tris]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dprt = importlib.import_module("pg2024-data-parallel-ray-tracing_b200")
import torch  # noqa: E402

w, h, tris = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080, 1000000)
W = 2
chunks, mats, lights = dprt.scene.make_scene(W, tris)
blobs = {}
for k in range(W):
    torch.manual_seed(19990201 + k)
    vis = dprt.proxy.spread_output_(dprt.proxy.make_proxy(256, 4).eval(), gain=3.0, seed=k)
    dep = dprt.proxy.spread_output_(dprt.proxy.make_proxy(256, 4).eval(), gain=1.0, seed=100 + k)
    blobs[k] = (dprt.proxy.pack_module(vis), dprt.proxy.pack_module(dep))
cfg = dprt.make_config(w, h, spp=1, bounces=2, scene_size=W, proxy_mode=1, path_gen_mode=1, mlp_dtype=0)
cam = dprt.scene.default_camera(w, h)
rs = []
for r in range(W):
    R = dprt.Renderer(cfg, rank=r, world=W, device=0)
    for c in chunks:
        if c.node_id == r:
            R.upload_chunk(c.index, c.desc(False), c.verts, c.normals, c.mats)
        else:
            R.upload_proxy(c.index, c.desc(True), *blobs[c.index])
    R.set_materials(mats); R.set_lights(lights); R.set_camera(cam)
    rs.append(R)
G = dprt.RankGroup(rs)
for R in rs:
    R.reset_frame()
for R in rs:
    R.stage_profile(True)
for s in range(2):
    G.run_sample(s)
for R in rs:
    R.synchronize()
print("stats", rs[0].stats())
print("stage ms (rank 0, 2 samples):", {k: (round(t, 3), n) for k, (t, n) in rs[0].stage_times().items() if n})
for R in rs:
    R.stage_profile(False)
P = rs[0]
n = 1 << 20
x = np.random.default_rng(0).random((n, 5)).astype(np.float16).view(np.uint16)
dx, dy = P.device_alloc(x.nbytes), P.device_alloc(n * 2)
P.h2d(dx, x)
for _ in range(3):
    P.mlp_infer_device(1, 0, dx, n, dy)
P.synchronize()
print("prof_workload done")
