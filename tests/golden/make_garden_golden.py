"""Freezes outputs of the ORACLE for the real-scene front end (run by hand when the specification is
changed on purpose, never at test time):

  garden_golden.npz   the procedural garden (real_scene.make_garden: indexed meshes, three instancing levels, leaf-card
                      cut-outs, checker albedo map, water, HDR environment map) on one rank: image [36, 64, 3] f32 after
                      2 samples x 3 bounces, primary-ray primitive ids / t bits WITH the alpha cut-outs applied, CRCs of the
                      flattened streams (vertices, normals, texture coordinates) and of the textures.

Pins the oracle's flatten, texture filter, cut-out rule and environment look-up, so that neither side can drift unnoticed.
"""
import importlib
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402

dprt = importlib.import_module("pg2024-data-parallel-ray-tracing_b200")

WIDTH, HEIGHT, SPP, BOUNCES = 64, 36, 2, 3


def build():
    from helpers import build_garden_pair
    _, world, g = build_garden_pair(O, 1, WIDTH, HEIGHT, spp=SPP, bounces=BOUNCES, gpu=False)
    return world, g


def scene_crc(g):
    ob = g["objects"][0]
    v, n, uv, m = O.flatten_instances(ob.meshes, ob.instances)
    parts = [v, n, uv, m] + [g["textures"][k] for k in sorted(g["textures"])] + [g["env_map"]]
    return np.array([zlib.crc32(np.ascontiguousarray(p).tobytes()) for p in parts], np.uint32)


def main():
    world, g = build()
    img = world.launch()
    hits = world.trace_closest(0, dprt.scene.camera_rays(dprt.scene.default_camera(WIDTH, HEIGHT)))
    np.savez_compressed(os.path.join(HERE, "garden_golden.npz"), image=img, prim=hits["primID"], t_bits=hits["t"].view(np.uint32),
                        scene_crc=scene_crc(g), walked=np.array([world.stats(0)["rays_walked"]]))
    print("garden_golden.npz written: image max", float(img.max()), "hits", int((hits["primID"] >= 0).sum()))


if __name__ == "__main__":
    main()
