"""One-GPU measurement of the real-scene front end at benchmark size: a procedural garden of ~1 M
INSTANCED triangles (30 K leaf instances of 5 indexed meshes, three instancing levels composed by the host), 1920x1080,
bounces = 4, spc = 4 -- rendered (a) with albedo / opacity maps (alpha cut-outs inside every trace, texture-mapped base
colour) and the lat-long environment map, (b) the same geometry with the material -> texture table cleared and the analytic
sky. Reports upload time (flatten + BVH8 build + copy), ms per sample and Mrays/s (rays that walked the BVH / device time,
CUDA events on the context's stream) for both, i.e. what the any-hit program and the texture look-ups cost on a textured scene.
Parity of this path is tests/test_gpu_real_scene.py (bit-exact against the oracle, including this size).
    python profiles/real_scene_bench.py [--clusters 9000] [--spp 4]"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dprt = importlib.import_module("pg2024-data-parallel-ray-tracing_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clusters", type=int, default=9000)
    ap.add_argument("--spp", type=int, default=4)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    a = ap.parse_args()
    g = dprt.real_scene.make_garden(1, clusters=a.clusters, ground=(300, 300), cluster_scale=0.004)
    ob = g["objects"][0]
    cfg = dprt.make_config(a.width, a.height, spp=a.spp, bounces=4, scene_size=1, proxy_mode=0)
    R = dprt.Renderer(cfg)
    t0 = time.time()
    R.upload_instanced_chunk(0, ob.desc(False), ob.meshes, ob.instances)
    upload_s = time.time() - t0
    R.set_materials(g["materials"]); R.set_lights(g["lights"]); R.set_camera(dprt.scene.default_camera(a.width, a.height))
    for slot, t in g["textures"].items():
        R.set_texture(slot, t)
    out = {"workload": f"garden: {ob.ntris} instanced triangles ({len(ob.instances)} instances of {len(ob.meshes)} meshes), "
                       f"{a.width}x{a.height}, bounces=4, spc=4, {a.spp} samples timed", "upload_s": round(upload_s, 3)}
    for name, mt, env in (("textured_cutouts_envmap", g["material_textures"], g["env_map"]),
                          ("untextured_analytic_sky", np.full(len(g["materials"]), -1, np.int32), None)):
        R.set_material_textures(mt)
        R.set_env_map(env, g["env_rotation"])
        R.reset_frame(); R.run_sample(0); R.synchronize()                  # warm-up
        R.reset_frame(); R.reset_stats()
        R.timer_start()
        for s in range(a.spp):
            R.run_sample(s)
        ms = R.timer_stop()
        st = R.stats()
        img = R.reduce_image(0)
        out[name] = {"ms_per_sample": ms / a.spp, "Mrays_per_s": st["rays_walked"] / (ms * 1e-3) / 1e6,
                     "rays_walked_per_sample": st["rays_walked"] / a.spp, "image_mean": float(img.mean()), "finite": bool(np.isfinite(img).all())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
