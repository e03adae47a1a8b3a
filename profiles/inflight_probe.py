#!/usr/bin/env python
"""Probe: K samples on one GPU rendered by 1, 2 or 3 contexts in flight (one host thread each, alternate samples).
usage: python profiles/inflight_probe.py [steps]"""
import importlib, os, sys, threading, time
import numpy as np
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")     # one hardware queue per stream (samples in flight)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dprt = importlib.import_module("pg2024-data-parallel-ray-tracing_b200")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
chunks, mats, lights = dprt.scene.make_scene(1, 1000000)
c = chunks[0]
cfg = dprt.make_config(1920, 1080, spp=1, bounces=4, scene_size=1)
cam = dprt.scene.default_camera(1920, 1080)
def make():
    R = dprt.Renderer(cfg)
    R.upload_chunk(0, c.desc(False), c.verts, c.normals, c.mats)
    R.set_materials(mats); R.set_lights(lights); R.set_camera(cam)
    R.reset_frame()
    for s in range(3):
        R.run_sample(s)
    R.synchronize()
    return R
Rs = [make() for _ in range(3)]
for k in (1, 2, 3, 1, 2):
    use = Rs[:k]
    for R in use:
        R.synchronize(); R.reset_stats()
    def work(R, j):
        for s in range(j, steps, k):
            R.run_sample(3 + s)
        R.synchronize()
    t0 = time.perf_counter()
    for R in use:
        R.timer_start()
    th = [threading.Thread(target=work, args=(R, j)) for j, R in enumerate(use)]
    [t.start() for t in th]; [t.join() for t in th]
    ms = max(R.timer_stop() for R in use)
    wall = (time.perf_counter() - t0) * 1e3
    rays = sum(R.stats()["rays_walked"] for R in use)
    print(f"contexts in flight {k}: {steps} samples in {ms:.2f} ms (wall {wall:.2f}) -> {ms / steps:.3f} ms/sample, {rays / ms / 1e3:.1f} Mrays/s", flush=True)
