#!/usr/bin/env python
"""Calibrates the x-cuts of bench.py's N > 1 landscape on ALL-BOUNCE load: the oracle renders one sample of the W-chunk scene
(coarse mesh of the same analytic landscape, reduced frame of the benchmark camera), the rays every chunk owner walks are
counted, the cuts move so that every slab carries the same share (load taken as uniform inside a slab), and so on for a few
rounds. Prints the table for scene.CALIBRATED_SLAB_CUTS. CPU only (the oracle is the measuring device here, not the product).
usage: python profiles/calibrate_slabs.py [rounds] [tris_per_chunk] [frame_scale]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dprt = importlib.import_module("pg2024-data-parallel-ray-tracing_b200")
from oracle import oracle as O
sys.path.insert(0, ROOT)
import bench

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 5
tris = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
scale = int(sys.argv[3]) if len(sys.argv) > 3 else 6
O.lib(); O.use_all_host_threads()


def loads(W, cuts, path_gen_mode=1):
    fw, fh = bench.frame_for(W)
    w, h = fw // scale, fh // scale
    cam = dprt.scene.default_camera(w, h)
    chunks, mats, lights = dprt.scene.make_scene(W, tris, layout="slabs", camera=cam, cuts=cuts)
    cfg = dprt.make_config(w, h, spp=1, bounces=4, scene_size=W, proxy_mode=0, path_gen_mode=path_gen_mode)
    world = O.World(cfg, W)
    for c in chunks:
        world.add_object(c.index, c.desc(False), c.verts, c.normals, c.mats)
    world.set_materials(mats); world.set_lights(lights); world.set_camera(cam)
    world.reset_frame(); world.render_sample(0); world.render_sample(1)
    out = np.array([world.stats(r)["rays_walked"] for r in range(W)], np.float64)
    world.close()
    return out


table = {}
for W in (2, 4, 8):
    cam = dprt.scene.default_camera(*bench.frame_for(W))
    cells = dprt.scene.balanced_slab_layout(W, cam, 0)
    cuts = [float(c[0][0]) for c in cells] + [1.0]
    best = None
    for it in range(rounds + 1):
        L = loads(W, cuts)
        imb = L.max() / L.mean()
        print(f"W={W} round {it}: max/mean {imb:.3f}  loads(M) {np.round(L / 1e6, 2).tolist()}  cuts {np.round(cuts, 4).tolist()}", flush=True)
        if best is None or imb < best[0]:
            best = (imb, list(cuts))
        if it == rounds:
            break
        # piecewise-uniform load density -> cumulative load -> new cuts at the k/W quantiles (damped)
        cum = np.concatenate([[0.0], np.cumsum(L)]) / L.sum()
        new = [0.0] + [float(np.interp(k / W, cum, cuts)) for k in range(1, W)] + [1.0]
        cuts = [0.5 * a + 0.5 * b for a, b in zip(cuts, new)]
    table[W] = [round(x, 5) for x in best[1]]
    print(f"W={W}: best max/mean {best[0]:.3f}", flush=True)
print("CALIBRATED_SLAB_CUTS =", table)
