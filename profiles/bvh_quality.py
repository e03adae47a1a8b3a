"""BVH8 build quality without a GPU: nodes fetched and triangles tested per ray by a scalar exact-tbest walker over the
blob dprt_bvh8_build produces, for every traversal stage of one sample of the benchmark workload (the counts bench.py's
roofline bills). Compare builder variants with DPRT_LIB=<variant .so>:
    python profiles/bvh_quality.py [--tris 1000000] [--scale 8]
`cost` = 4 x nodes + 1 x triangles per ray: a node step costs the trace kernel about four triangle tests' worth of issue
slots (280 SASS instructions at 21 of 32 lanes against ~100 at 32 lanes, profiles/r2_ncu_trace.txt)."""
import argparse
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dprt = importlib.import_module("pg2024-data-parallel-ray-tracing_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tris", type=int, default=1000000)
    ap.add_argument("--scale", type=int, default=8)
    ap.add_argument("--bounces", type=int, default=4)
    a = ap.parse_args()
    from oracle import oracle as O        # measurement script: the checker's walker counts, nothing is shipped from here
    O.lib(); O.use_all_host_threads()
    w, h = 1920 // a.scale, 1080 // a.scale
    chunks, mats, lights = dprt.scene.make_scene(1, a.tris)
    cfg = dprt.make_config(w, h, spp=1, bounces=a.bounces, scene_size=1, proxy_mode=0)
    world = O.World(cfg, 1)
    c = chunks[0]
    world.add_object(c.index, c.desc(False), c.verts, c.normals, c.mats)
    t0 = time.time()
    nodes, tris, depth = dprt.build_bvh8(c.verts, c.mats)
    build_s = time.time() - t0
    world.set_bvh8(c.index, nodes, tris)
    world.set_materials(mats); world.set_lights(lights); world.set_camera(dprt.scene.default_camera(w, h))
    world.count_bvh8(True); world.reset_frame(); world.render_sample(0)
    out = {"lib": os.environ.get("DPRT_LIB", "default"), "nodes": int(nodes.size), "tris": int(tris.size), "depth": depth, "build_s": round(build_s, 2)}
    tot_n = tot_t = tot_r = 0
    for name in ("traverse", "shade", "shadow_trace"):
        n, t, r = world.bvh8_counters(0)[name]
        if r:
            out[name] = {"nodes_per_ray": round(n / r, 3), "tris_per_ray": round(t / r, 3), "rays": r}
        if name != "shade":
            tot_n += n; tot_t += t; tot_r += r
    out["cost_per_ray"] = round((4 * tot_n + tot_t) / tot_r, 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
