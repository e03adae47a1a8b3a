"""Freezes outputs of the ORACLE for a small two-chunk scene (run by hand when the arithmetic specification is changed on
purpose; never at test time):

  render_golden.npz   image [36, 64, 3] f32 after 2 samples x 3 bounces with ray migration (proxies off), the primary-ray
                      hit primitive ids / t bits of chunk 0, per-rank final path counts and exchange statistics, and the
                      Vis-pipeline training samples (features / labels) of 256 fixed rays.

The parity tests compare libdprt with the live oracle; this fixture pins the oracle itself, so that an accidental
change of the specification (a reordered float operation, a different tie rule) cannot go unnoticed on either side.
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

dprt = importlib.import_module("pg2024-data-parallel-ray-tracing_b200")

W, TRIS, WIDTH, HEIGHT, SPP, BOUNCES = 2, 3000, 64, 36, 2, 3


def build():
    chunks, mats, lights = dprt.scene.make_scene(W, TRIS, water_frac=0.03)
    cam = dprt.scene.default_camera(WIDTH, HEIGHT)
    cfg = dprt.make_config(WIDTH, HEIGHT, spp=SPP, bounces=BOUNCES, scene_size=W, proxy_mode=0, path_gen_mode=1)
    world = O.World(cfg, W)
    for c in chunks:
        world.add_object(c.index, c.desc(False), c.verts, c.normals, c.mats)
    world.set_materials(mats); world.set_lights(lights); world.set_camera(cam)
    return world, chunks, cam, cfg


def main():
    world, chunks, cam, cfg = build()
    img = world.launch()
    rays = dprt.scene.camera_rays(cam)
    hits = world.trace_closest(0, rays)
    trays = dprt.proxy_train.sample_training_rays(chunks[1].aabb_min, chunks[1].aabb_max, 256, seed=7)
    feat, lab = world.gen_train_data(1, trays)
    st = [world.stats(r) for r in range(W)]
    import zlib
    crc = np.array([zlib.crc32(np.ascontiguousarray(c.verts).tobytes()) for c in chunks], np.uint32)
    np.savez_compressed(os.path.join(HERE, "render_golden.npz"), image=img, prim=hits["primID"], t_bits=hits["t"].view(np.uint32),
                        train_rays=trays.view(np.uint8), scene_crc=crc, train_feat=feat, train_label=lab,
                        path_size=np.array([world.path_size(r) for r in range(W)]),
                        sent=np.array([s["paths_sent_offrank"] for s in st]), iters=np.array([s["exchange_iters"] for s in st]),
                        walked=np.array([s["rays_walked"] for s in st]))
    print("render_golden.npz written: image max", float(img.max()), "hits", int((hits["primID"] >= 0).sum()), "sent", [s["paths_sent_offrank"] for s in st])


if __name__ == "__main__":
    main()
