#!/usr/bin/env python
"""Per-launch device time of every stage of one runSample at N=1 (CUDA events around each C-ABI stage call).
usage: python profiles/per_launch.py [tris] [width height]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dprt = importlib.import_module("pg2024-data-parallel-ray-tracing_b200")
tris = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
chunks, mats, lights = dprt.scene.make_scene(1, tris)
cfg = dprt.make_config(w, h, spp=1, bounces=4, scene_size=1)
R = dprt.Renderer(cfg)
c = chunks[0]
R.upload_chunk(0, c.desc(False), c.verts, c.normals, c.mats)
R.set_materials(mats); R.set_lights(lights); R.set_camera(dprt.scene.default_camera(w, h))
R.reset_frame()
for s in range(3):
    R.run_sample(s)
R.synchronize()
def timed(name, fn, n):
    R.timer_start(); fn(); ms = R.timer_stop()
    print(f"  {name:14s} n={n:9d}  {ms * 1e3:8.1f} us  {n / ms / 1e3 if ms > 0 else 0:8.1f} M/s")
    return ms
tot = 0.0
R.begin_sample(3); R.path_gen()
for b in range(5):
    print(f"bounce {b}")
    tot += timed("traverse", R.traverse, R.path_size)
    R.partition(); R.exchange()
    tot += timed("shade", R.shade, R.path_size)
    R.reset_nn()
    tot += timed("shadow_trace", R.shadow_trace, R.shadow_path_size)
    R.frame_buffer_update()
print("trace stages total ms", tot)
