"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE (see oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import importlib
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
REF_RNG_PATH = os.path.join(_HERE, "_ref", "librefrng.so")

D = importlib.import_module("pg2024-data-parallel-ray-tracing_b200.ctypes_defs")

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle`")
        L = C.CDLL(LIB_PATH)
        L.orc_world_create.restype = C.c_void_p
        L.orc_world_create.argtypes = [C.POINTER(D.Config), C.c_int]
        L.orc_world_destroy.argtypes = [C.c_void_p]
        L.orc_world_add_object.argtypes = [C.c_void_p, C.c_int, C.POINTER(D.ObjectDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_world_add_object_uv.argtypes = [C.c_void_p, C.c_int, C.POINTER(D.ObjectDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_world_add_instanced_object.argtypes = [C.c_void_p, C.c_int, C.POINTER(D.ObjectDesc), C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
        L.orc_flatten_count.restype = C.c_int64
        L.orc_flatten_count.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
        L.orc_flatten.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.orc_world_set_texture.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.orc_world_set_material_textures.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_world_set_env_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float]
        L.orc_texture_sample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.orc_env_lookup.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_world_set_model.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.orc_world_set_materials.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_world_set_lights.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_world_set_camera.argtypes = [C.c_void_p, C.POINTER(D.Camera)]
        L.orc_enable_hit_prim.argtypes = [C.c_void_p, C.c_int]
        for name in ("orc_reset_frame",):
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("orc_begin_sample", "orc_path_gen", "orc_traverse", "orc_partition", "orc_shade", "orc_reset_nn",
                     "orc_shadow_trace", "orc_secondary_trace", "orc_frame_buffer_update", "orc_depth_buffer_update",
                     "orc_target_node_update", "orc_render_sample"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int]
        L.orc_exchange.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.orc_bucket_queries.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.orc_proxy_infer.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_image.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_get_path_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_set_path_size.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_set_query_total.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_get_stats.argtypes = [C.c_void_p, C.c_int, C.POINTER(D.Stats)]
        L.orc_download.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_upload.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_trace_closest.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        L.orc_gen_train_data.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_world_set_bvh8.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        L.orc_count_bvh8.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_bvh8_counters.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_gen_precom_data.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_bvh8_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_mlp_forward.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_tea4.restype = C.c_uint32
        L.orc_tea4.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_rnd_sequence.argtypes = [C.c_uint32, C.c_int, C.c_void_p]
        L.orc_sincos2pi.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_acos.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_atan2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_f2h.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def tea4(a, b):
    return int(lib().orc_tea4(a & 0xFFFFFFFF, b & 0xFFFFFFFF))


def rnd_sequence(seed, n):
    out = np.zeros(n, np.float32)
    lib().orc_rnd_sequence(seed & 0xFFFFFFFF, n, _p(out))
    return out


def num_threads():
    return int(lib().orc_num_threads())


def use_all_host_threads():
    """The timed CPU arms use every hardware thread the process may run on, whatever OMP_NUM_THREADS says."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(int(n))
    return num_threads()


def mlp_forward(blob, x_half):
    x = np.ascontiguousarray(x_half, np.uint16).reshape(-1, 5)
    b = np.frombuffer(blob, np.uint8)
    yf = np.zeros(x.shape[0], np.float32)
    yh = np.zeros(x.shape[0], np.uint16)
    rc = lib().orc_mlp_forward(_p(b), b.size, _p(x), x.shape[0], _p(yf), _p(yh))
    if rc:
        raise RuntimeError("orc_mlp_forward: bad blob")
    return yf, yh


def bvh8_trace(nodes, tris, rays):
    """Walk the product's BVH8 blob on the CPU: (hits, nodes_visited, tris_tested)."""
    nodes = np.ascontiguousarray(nodes, D.NODE_DTYPE)
    tris = np.ascontiguousarray(tris, D.TRI_DTYPE)
    rays = np.ascontiguousarray(rays, D.RAY_DTYPE)
    hits = np.zeros(rays.size, D.HIT_DTYPE)
    nv, tt = C.c_int64(), C.c_int64()
    lib().orc_bvh8_trace(_p(nodes), _p(tris), _p(rays), rays.size, _p(hits), C.byref(nv), C.byref(tt))
    return hits, int(nv.value), int(tt.value)


def flatten_instances(meshes, instances):
    """The oracle's own flatten of indexed, instanced meshes: (verts9, normals9, uv6 or None, mat_ids)."""
    md, nm, ins, ni, keep = D.pack_meshes(meshes, instances)
    n = lib().orc_flatten_count(md, nm, ins, ni)
    if n <= 0:
        raise RuntimeError("orc_flatten_count: invalid description")
    v, nr, uv, mats = np.zeros((n, 9), np.float32), np.zeros((n, 9), np.float32), np.zeros((n, 6), np.float32), np.zeros(n, np.int32)
    has = C.c_int(0)
    if lib().orc_flatten(md, nm, ins, ni, _p(v), _p(nr), _p(uv), _p(mats), C.byref(has)):
        raise RuntimeError("orc_flatten failed")
    del keep
    return v, nr, (uv if has.value else None), mats


def texture_sample(rgba, u, v, clamp_v=False):
    t = np.ascontiguousarray(rgba, np.float32)
    uu, vv = np.ascontiguousarray(u, np.float32), np.ascontiguousarray(v, np.float32)
    out = np.zeros((uu.size, 4), np.float32)
    assert lib().orc_texture_sample(_p(t), t.shape[1], t.shape[0], _p(uu), _p(vv), uu.size, int(clamp_v), _p(out)) == 0
    return out


def env_lookup(rgba, rotation, dirs):
    t = np.ascontiguousarray(rgba, np.float32)
    d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
    out = np.zeros((d.shape[0], 3), np.float32)
    assert lib().orc_env_lookup(_p(t), t.shape[1], t.shape[0], float(rotation), _p(d), d.shape[0], _p(out)) == 0
    return out


class World:
    """All scene objects + W simulated ranks (the in-process stand-in for the MPI job)."""

    def __init__(self, cfg, W=1):
        self.L = lib()
        self.cfg, self.W = cfg, W
        self.N = cfg.width * cfg.height
        self.h = C.c_void_p(self.L.orc_world_create(C.byref(cfg), W))

    def close(self):
        if self.h:
            self.L.orc_world_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_object(self, scene_index, desc, verts9, normals9, mat_ids):
        v = np.ascontiguousarray(verts9, np.float32).reshape(-1, 9)
        n = None if normals9 is None else np.ascontiguousarray(normals9, np.float32)
        m = None if mat_ids is None else np.ascontiguousarray(mat_ids, np.int32)
        assert self.L.orc_world_add_object(self.h, scene_index, C.byref(desc), _p(v), _p(n), _p(m), v.shape[0]) == 0

    def add_object_uv(self, scene_index, desc, verts9, normals9, uv6, mat_ids):
        v = np.ascontiguousarray(verts9, np.float32).reshape(-1, 9)
        n = None if normals9 is None else np.ascontiguousarray(normals9, np.float32)
        u = None if uv6 is None else np.ascontiguousarray(uv6, np.float32)
        m = None if mat_ids is None else np.ascontiguousarray(mat_ids, np.int32)
        assert self.L.orc_world_add_object_uv(self.h, scene_index, C.byref(desc), _p(v), _p(n), _p(u), _p(m), v.shape[0]) == 0

    def add_instanced_object(self, scene_index, desc, meshes, instances):
        md, nm, ins, ni, keep = D.pack_meshes(meshes, instances)
        assert self.L.orc_world_add_instanced_object(self.h, scene_index, C.byref(desc), md, nm, ins, ni) == 0
        del keep

    def set_texture(self, texture_index, rgba):
        if rgba is None:
            assert self.L.orc_world_set_texture(self.h, texture_index, None, 0, 0) == 0
            return
        t = np.ascontiguousarray(rgba, np.float32)
        assert self.L.orc_world_set_texture(self.h, texture_index, _p(t), t.shape[1], t.shape[0]) == 0

    def set_material_textures(self, texture_index):
        t = np.ascontiguousarray(texture_index, np.int32)
        assert self.L.orc_world_set_material_textures(self.h, _p(t), t.size) == 0

    def set_env_map(self, rgba, rotation_offset=0.0):
        if rgba is None:
            assert self.L.orc_world_set_env_map(self.h, None, 0, 0, 0.0) == 0
            return
        t = np.ascontiguousarray(rgba, np.float32)
        assert self.L.orc_world_set_env_map(self.h, _p(t), t.shape[1], t.shape[0], float(rotation_offset)) == 0

    def set_model(self, scene_index, kind, blob):
        b = np.frombuffer(blob, np.uint8)
        assert self.L.orc_world_set_model(self.h, scene_index, kind, _p(b), b.size) == 0

    def set_materials(self, mats):
        m = np.ascontiguousarray(mats, D.MATERIAL_DTYPE)
        self.L.orc_world_set_materials(self.h, _p(m), m.size)

    def set_lights(self, lights):
        l = np.ascontiguousarray(lights, D.LIGHT_DTYPE)
        self.L.orc_world_set_lights(self.h, _p(l), l.size)

    def set_camera(self, cam):
        self.L.orc_world_set_camera(self.h, C.byref(cam))

    def enable_hit_prim(self, enable=True):
        self.L.orc_enable_hit_prim(self.h, int(enable))

    def reset_frame(self):
        self.L.orc_reset_frame(self.h)

    def begin_sample(self, s):
        self.L.orc_begin_sample(self.h, s)

    def path_gen(self, rank=0):
        assert self.L.orc_path_gen(self.h, rank) == 0

    def traverse(self, rank=0):
        assert self.L.orc_traverse(self.h, rank) == 0

    def partition(self, rank=0):
        assert self.L.orc_partition(self.h, rank) == 0

    def exchange(self):
        d = C.c_int(0)
        self.L.orc_exchange(self.h, C.byref(d))
        return bool(d.value)

    def shade(self, rank=0):
        assert self.L.orc_shade(self.h, rank) == 0

    def reset_nn(self, rank=0):
        assert self.L.orc_reset_nn(self.h, rank) == 0

    def shadow_trace(self, rank=0):
        assert self.L.orc_shadow_trace(self.h, rank) == 0

    def secondary_trace(self, rank=0):
        assert self.L.orc_secondary_trace(self.h, rank) == 0

    def bucket_queries(self, rank, which, inside_only):
        t = C.c_int(0)
        assert self.L.orc_bucket_queries(self.h, rank, which, int(inside_only), C.byref(t)) == 0
        return int(t.value)

    def proxy_infer(self, rank, kind, pred_offset=0):
        assert self.L.orc_proxy_infer(self.h, rank, kind, pred_offset) == 0

    def frame_buffer_update(self, rank=0):
        assert self.L.orc_frame_buffer_update(self.h, rank) == 0

    def depth_buffer_update(self, rank=0):
        assert self.L.orc_depth_buffer_update(self.h, rank) == 0

    def target_node_update(self, rank=0):
        assert self.L.orc_target_node_update(self.h, rank) == 0

    def render_sample(self, s):
        self.L.orc_render_sample(self.h, s)

    def launch(self):
        self.reset_frame()
        for s in range(self.cfg.spp):
            self.render_sample(s)
        return self.image()

    def image(self):
        out = np.zeros((self.cfg.height, self.cfg.width, 3), np.float32)
        self.L.orc_image(self.h, _p(out))
        return out

    def path_size(self, rank=0):
        a, b = C.c_int(), C.c_int()
        self.L.orc_get_path_size(self.h, rank, C.byref(a), C.byref(b))
        return int(a.value)

    def shadow_path_size(self, rank=0):
        a, b = C.c_int(), C.c_int()
        self.L.orc_get_path_size(self.h, rank, C.byref(a), C.byref(b))
        return int(b.value)

    def set_path_size(self, rank, n):
        self.L.orc_set_path_size(self.h, rank, int(n))

    def set_query_total(self, rank, n):
        self.L.orc_set_query_total(self.h, rank, int(n))

    def set_bvh8(self, scene_index, nodes, tris):
        """A copy of the PRODUCT's BVH8 blob (dprt.build_bvh8 / dprt_bvh8_copy): walked by the counting walker only."""
        nodes = np.ascontiguousarray(nodes, D.NODE_DTYPE); tris = np.ascontiguousarray(tris, D.TRI_DTYPE)
        assert self.L.orc_world_set_bvh8(self.h, scene_index, _p(nodes), nodes.size, _p(tris), tris.size) == 0

    def count_bvh8(self, enable=True):
        self.L.orc_count_bvh8(self.h, 1 if enable else 0)

    def bvh8_counters(self, rank=0, reset=False):
        """{stage: (nodes fetched, triangles tested, rays walked)} of the scalar exact-tbest walk over the product's BVH8."""
        out = np.zeros(3 * D.STAGE_COUNT, np.int64)
        assert self.L.orc_get_bvh8_counters(self.h, rank, _p(out), 1 if reset else 0) == 0
        return {D.STAGE_NAMES[s]: (int(out[3 * s]), int(out[3 * s + 1]), int(out[3 * s + 2])) for s in range(D.STAGE_COUNT)}

    def stats(self, rank=0):
        s = D.Stats()
        self.L.orc_get_stats(self.h, rank, C.byref(s))
        return s.as_dict()

    def download(self, rank, buf, count, offset=0):
        dt = D.BUFFER_DTYPES[buf]
        out = np.zeros(count, dt)
        if count:
            assert self.L.orc_download(self.h, rank, buf, offset * dt.itemsize, _p(out), out.nbytes) == 0, "oracle download range"
        return out

    def upload(self, rank, buf, array, offset=0):
        dt = D.BUFFER_DTYPES[buf]
        a = np.ascontiguousarray(array, dt)
        if a.size:
            assert self.L.orc_upload(self.h, rank, buf, offset * dt.itemsize, _p(a), a.nbytes) == 0, "oracle upload range"

    def gen_train_data(self, scene_index, rays):
        r = np.ascontiguousarray(rays, D.RAY_DTYPE)
        feat, lab = np.zeros((r.size, 5), np.float32), np.zeros(r.size, np.float32)
        assert self.L.orc_gen_train_data(self.h, scene_index, _p(r), r.size, _p(feat), _p(lab)) == 0
        return feat, lab

    def gen_precom_data(self, scene_index, rays):
        r = np.ascontiguousarray(rays, D.RAY_DTYPE)
        feat, lab, valid = np.zeros((r.size, 5), np.float32), np.zeros(r.size, np.float32), np.zeros(r.size, np.uint8)
        assert self.L.orc_gen_precom_data(self.h, scene_index, _p(r), r.size, _p(feat), _p(lab), _p(valid)) == 0
        return feat, lab, valid

    def trace_closest(self, rank, rays, brute=False):
        r = np.ascontiguousarray(rays, D.RAY_DTYPE)
        hits = np.zeros(r.size, D.HIT_DTYPE)
        assert self.L.orc_trace_closest(self.h, rank, _p(r), r.size, _p(hits), 1 if brute else 0) == 0
        return hits


def ref_rng():
    """The reference's own optix/random.hpp compiled into oracle/_ref (None when not built)."""
    if not os.path.exists(REF_RNG_PATH):
        return None
    R = C.CDLL(REF_RNG_PATH)
    R.ref_tea4.restype = C.c_uint32
    R.ref_tea4.argtypes = [C.c_uint32, C.c_uint32]
    R.ref_rnd_sequence.argtypes = [C.c_uint32, C.c_int, C.c_void_p]
    return R
