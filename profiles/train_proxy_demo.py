#!/usr/bin/env python
"""Closes the loop of SURVEY.md 8f rows 1-2 on one GPU: for a 2-chunk scene, generate Vis-pipeline samples with
dprt_gen_train_data, train the vis and depth proxies of every chunk (proxy_train.train_chunk_proxies), export the blobs
and render with proxies on; compare with the exact (proxy-off, ray-migrating) image and with untrained proxies.
usage: python profiles/train_proxy_demo.py [tris] [rays] [epochs]"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dprt = importlib.import_module("pg2024-data-parallel-ray-tracing_b200")
import torch  # noqa: E402

tris = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
nrays = int(sys.argv[2]) if len(sys.argv) > 2 else 400000
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 15
W, w, h, spp, bounces = 2, 256, 144, 4, 2
chunks, mats, lights = dprt.scene.make_scene(W, tris)
cam = dprt.scene.default_camera(w, h)


def render(proxy_mode, blobs):
    cfg = dprt.make_config(w, h, spp=spp, bounces=bounces, scene_size=W, proxy_mode=proxy_mode, path_gen_mode=1, mlp_dtype=0)
    rs = []
    for r in range(W):
        R = dprt.Renderer(cfg, rank=r, world=W, device=0)
        for c in chunks:
            if c.node_id == r:
                R.upload_chunk(c.index, c.desc(False), c.verts, c.normals, c.mats)
            else:
                vb, db = blobs.get(c.index, (None, None))
                R.upload_proxy(c.index, c.desc(True), vb, db)
        R.set_materials(mats); R.set_lights(lights); R.set_camera(cam)
        rs.append(R)
    img = dprt.RankGroup(rs).launch()
    st = [R.stats() for R in rs]
    for R in rs:
        R.close()
    return img, st


def relmse(a, b):
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


exact, st_exact = render(0, {})
untrained = {}
for k in range(W):
    torch.manual_seed(19990201 + k)
    m = dprt.proxy.make_proxy(256, 4).eval()
    untrained[k] = (dprt.proxy.pack_module(m), dprt.proxy.pack_module(m))
img_u, st_u = render(1, untrained)
trained, infos = {}, {}
t0 = time.time()
cfg1 = dprt.make_config(16, 16, scene_size=W)
for c in chunks:
    R = dprt.Renderer(cfg1, rank=c.node_id, world=W, device=0)
    R.upload_chunk(c.index, c.desc(False), c.verts, c.normals, c.mats)
    vb, db, info = dprt.proxy_train.train_chunk_proxies(lambda rays: R.gen_train_data(c.index, rays), c.aabb_min, c.aabb_max,
                                                        n_rays=nrays, epochs=epochs, seed=c.index, device="cuda")
    trained[c.index], infos[c.index] = (vb, db), info
    R.close()
train_s = time.time() - t0
img_t, st_t = render(1, trained)
out = {"tris_per_chunk": tris, "train_rays": nrays, "epochs": epochs, "train_seconds": train_s,
       "relmse_untrained_vs_exact": relmse(img_u, exact), "relmse_trained_vs_exact": relmse(img_t, exact),
       "paths_sent_offrank": {"exact": sum(s["paths_sent_offrank"] for s in st_exact), "proxy_trained": sum(s["paths_sent_offrank"] for s in st_t)},
       "nn_queries_trained": sum(s["nn_queries"] for s in st_t),
       "chunks": {k: {"hit_fraction": v["hit_fraction"], "vis_loss_first_last": [v["vis_test_loss"][0], v["vis_test_loss"][-1]],
                      "depth_loss_first_last": [v["depth_test_loss"][0], v["depth_test_loss"][-1]]} for k, v in infos.items()}}
print(json.dumps(out, indent=1))
