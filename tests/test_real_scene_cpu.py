"""CPU tests of the real-scene front end: the host-side flatten of indexed, instanced meshes and the
texture / environment look-ups of the PRODUCT (libdprt.so, host-only entry points compiled from the kernels' own source)
against the oracle bit for bit, both against a float64 numpy model, and the oracle's alpha cut-out / textured shading.
No GPU needed; the GPU parity of the same scene is tests/test_gpu_real_scene.py."""
import os

import numpy as np
import pytest

from helpers import D, build_garden_pair, dprt, random_rays


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_flatten_product_equals_oracle_and_numpy(oracle):
    g = dprt.real_scene.make_garden(W=2, clusters=20)
    for ob in g["objects"]:
        v, n, uv, m = dprt.flatten_instances(ob.meshes, ob.instances)
        vo, no, uvo, mo = oracle.flatten_instances(ob.meshes, ob.instances)
        assert v.shape[0] == ob.ntris and uv is not None and uvo is not None
        assert np.array_equal(_bits(v), _bits(vo)) and np.array_equal(_bits(n), _bits(no))
        assert np.array_equal(_bits(uv), _bits(uvo)) and np.array_equal(m, mo)
        # float64 model: corners M p, normals A^-T n, attributes gathered through their own index arrays
        base = 0
        for mi, M in ob.instances:
            me = ob.meshes[mi]
            M = np.asarray(M, np.float64)
            idx = np.asarray(me["indices"]).reshape(-1, 3)
            P = np.asarray(me["positions"], np.float32).astype(np.float64).reshape(-1, 3)[idx] @ M[:, :3].T + M[:, 3]
            G = np.linalg.inv(M[:, :3]).T
            Nn = np.asarray(me["normals"], np.float32).astype(np.float64).reshape(-1, 3)[np.asarray(me["normal_indices"]).reshape(-1, 3)] @ G.T
            k = idx.shape[0]
            assert np.allclose(v[base:base + k].reshape(k, 3, 3), P, rtol=0, atol=2e-6)
            assert np.allclose(n[base:base + k].reshape(k, 3, 3), Nn, rtol=1e-5, atol=1e-5)
            if me["texcoords"] is not None:
                T = np.asarray(me["texcoords"], np.float32).reshape(-1, 2)[np.asarray(me["texcoord_indices"]).reshape(-1, 3)]
                assert np.array_equal(uv[base:base + k].reshape(k, 3, 2), T)
            else:
                assert not uv[base:base + k].any()
            assert (m[base:base + k] == me["material"]).all()
            base += k
        assert base == v.shape[0]
    # a mirrored instance exists (det < 0): its normals must still point along the transformed surface normal
    dets = [np.linalg.det(np.asarray(M, np.float64)[:, :3]) for _, M in g["objects"][0].instances]
    assert min(dets) < 0 < max(dets)


def test_flatten_rejects_bad_descriptions(oracle):
    g = dprt.real_scene.make_garden(W=1, clusters=4)
    ob = g["objects"][0]
    with pytest.raises(dprt.DprtError):
        dprt.flatten_instances(ob.meshes, [(len(ob.meshes), np.eye(4)[:3])])                    # mesh index out of range
    with pytest.raises(dprt.DprtError):
        dprt.flatten_instances(ob.meshes, [(0, np.zeros((3, 4), np.float32))])                  # singular transform
    bad = dict(ob.meshes[1]); bad["normal_indices"] = np.asarray(bad["normal_indices"]) + 10 ** 6
    with pytest.raises(dprt.DprtError):
        dprt.flatten_instances([bad], [(0, np.eye(4)[:3])])                                     # attribute index out of range
    with pytest.raises(RuntimeError):
        oracle.flatten_instances([bad], [(0, np.eye(4)[:3])])


def _numpy_bilinear(tex, u, v, clamp_v):
    """float64 model of the look-up: texel centres at (i + 0.5) / n, wrap in u, wrap or clamp in v."""
    h, w, _ = tex.shape
    u = np.asarray(u, np.float64); v = np.asarray(v, np.float64)
    u = u - np.floor(u)
    v = np.clip(v, 0, 1) if clamp_v else v - np.floor(v)
    x, y = u * w - 0.5, v * h - 0.5
    x0, y0 = np.floor(x), np.floor(y)
    fx, fy = (x - x0)[:, None], (y - y0)[:, None]
    x0, y0 = x0.astype(int), y0.astype(int)
    xa, xb = x0 % w, (x0 + 1) % w
    ya, yb = (np.clip(y0, 0, h - 1), np.clip(y0 + 1, 0, h - 1)) if clamp_v else (y0 % h, (y0 + 1) % h)
    t = tex.astype(np.float64)
    top = t[ya, xa] * (1 - fx) + t[ya, xb] * fx
    bot = t[yb, xa] * (1 - fx) + t[yb, xb] * fx
    return top * (1 - fy) + bot * fy


@pytest.mark.parametrize("shape", [(32, 32), (7, 13), (1, 1), (2, 64)])
@pytest.mark.parametrize("clamp_v", [False, True])
def test_texture_lookup_product_equals_oracle_and_numpy(oracle, shape, clamp_v):
    rng = np.random.default_rng(shape[0] * 100 + shape[1] + int(clamp_v))
    tex = rng.uniform(0, 4, (shape[0], shape[1], 4)).astype(np.float32)
    u = np.concatenate([rng.uniform(-3, 3, 4000), [0.0, 1.0, -1e-9, 1.0 - 1e-7, 0.5, 0.5 / shape[1], 1e31, np.nan, np.inf, -np.inf]]).astype(np.float32)
    v = np.concatenate([rng.uniform(-3, 3, 4000), [0.0, 1.0, 1.0, -1e-9, 0.5, 0.5 / shape[0], 0.25, 0.25, np.nan, 0.75]]).astype(np.float32)
    a = dprt.spec_texture_sample(tex, u, v, clamp_v)
    b = oracle.texture_sample(tex, u, v, clamp_v)
    assert np.array_equal(_bits(a), _bits(b)), "libdprt (dprt_math.cuh, host compile) and the oracle disagree"
    fin = np.isfinite(u) & np.isfinite(v) & (np.abs(u) < 1e30) & (np.abs(v) < 1e30)
    u, v = np.where(fin, u, np.float32(0.3)), np.where(fin, v, np.float32(0.3))
    fin &= np.abs(np.float64(u) - np.round(np.float64(u))) > 1e-6          # at an integer u the float32 wrap may land on either side
    if not clamp_v:
        fin &= np.abs(np.float64(v) - np.round(np.float64(v))) > 1e-6
    m = _numpy_bilinear(tex, u[fin], v[fin], clamp_v)
    assert np.allclose(a[fin], m, rtol=0, atol=2e-4 * max(shape))        # u W - 0.5 in float32: weights to ~1e-5 W
    assert np.isfinite(a).all()                                          # NaN / inf coordinates are defined (texel (0, 0) side), never UB


def test_env_lookup_product_equals_oracle(oracle):
    env = dprt.real_scene.sky_env_map(64, 32)
    rng = np.random.default_rng(5)
    d = rng.normal(size=(5000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = np.concatenate([d, [[0, 0, 1], [0, 0, -1], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [-1, -1e-8, 0]]]).astype(np.float32)
    for rot in (0.0, 0.7, 6.0):
        a = dprt.spec_env_lookup(env, rot, d)
        b = oracle.env_lookup(env, rot, d)
        assert np.array_equal(_bits(a), _bits(b))
        # float64 model of the direction -> (u, v) mapping, then the bilinear model
        phi = np.mod(np.arctan2(np.float64(d[:, 1]), np.float64(d[:, 0])), 2 * np.pi) + rot
        phi = np.where(phi > 2 * np.pi, phi - 2 * np.pi, phi)
        theta = np.arccos(np.clip(np.float64(d[:, 2]), -1, 1))
        m = _numpy_bilinear(env, phi / (2 * np.pi), theta / np.pi, True)[:, :3]
        edge = np.abs(phi / (2 * np.pi) - np.round(phi / (2 * np.pi))) < 1e-5
        assert np.allclose(a[~edge], m[~edge], rtol=2e-3, atol=2e-3)
    # the zenith row sees the sky, the nadir row the darkened ground half
    assert dprt.spec_env_lookup(env, 0.0, [[0, 0, 1.0]])[0, 2] > dprt.spec_env_lookup(env, 0.0, [[0, 0, -1.0]])[0, 2]


def test_oracle_alpha_cutout_semantics(oracle):
    """Cut-outs drop candidates inside the traversal (closest hit moves to what lies behind), identically in the BVH and the
    brute-force walker; without textures the same rays stop at the leaf cards."""
    _, world, g = build_garden_pair(oracle, 1, 64, 36, gpu=False)
    _, plain, _ = build_garden_pair(oracle, 1, 64, 36, gpu=False, textures=False)
    rays = random_rays(20000, 3, lo=0.0, hi=1.0)
    rays["origin"][:, 2] = 0.9
    rays["direction"][:, 2] = -np.abs(rays["direction"][:, 2]) - 0.3
    rays["direction"] /= np.linalg.norm(rays["direction"], axis=1, keepdims=True)
    h_bvh, h_brute = world.trace_closest(0, rays), world.trace_closest(0, rays, brute=True)
    assert np.array_equal(h_bvh["primID"], h_brute["primID"]) and np.array_equal(_bits(h_bvh["t"]), _bits(h_brute["t"]))
    h_plain = plain.trace_closest(0, rays)
    ob = g["objects"][0]
    _, _, _, mats = oracle.flatten_instances(ob.meshes, ob.instances)
    leaf = np.isin(mats, [dprt.real_scene.MAT_LEAF, dprt.real_scene.MAT_LEAF2])
    hit_leaf_plain = (h_plain["primID"] >= 0) & leaf[np.maximum(h_plain["primID"], 0)]
    hit_leaf_tex = (h_bvh["primID"] >= 0) & leaf[np.maximum(h_bvh["primID"], 0)]
    passed = hit_leaf_plain & (h_bvh["primID"] != h_plain["primID"])
    assert hit_leaf_plain.sum() > 200 and passed.sum() > 50, (hit_leaf_plain.sum(), passed.sum())   # cut-outs are exercised
    assert 0 < hit_leaf_tex.sum() < hit_leaf_plain.sum()                                             # the opaque ellipse still stops rays
    assert (h_bvh["t"][passed] > h_plain["t"][passed]).all()                                         # ... and what passes lands farther away
    same = ~hit_leaf_plain
    assert np.array_equal(h_bvh["primID"][same & ~hit_leaf_tex], h_plain["primID"][same & ~hit_leaf_tex])


def test_oracle_garden_render_uses_textures_and_env_map(oracle):
    _, world, _ = build_garden_pair(oracle, 1, 96, 54, spp=2, bounces=2, gpu=False)
    _, flat, _ = build_garden_pair(oracle, 1, 96, 54, spp=2, bounces=2, gpu=False, textures=False, env_map=False)
    a, b = world.launch(), flat.launch()
    assert np.isfinite(a).all() and a.max() > 0 and np.isfinite(b).all()
    assert not np.allclose(a, b)
    # the untextured ground is magenta (baseColor 0.9, 0.1, 0.9) under orange lights; the checker map is green-ish
    assert a[..., 1].mean() / a[..., 0].mean() > 1.5 * b[..., 1].mean() / b[..., 0].mean()


def test_oracle_garden_two_ranks_migrate(oracle):
    _, w2, _ = build_garden_pair(oracle, 2, 64, 36, spp=1, bounces=2, gpu=False)
    img = w2.launch()
    assert np.isfinite(img).all() and img.max() > 0
    assert sum(w2.stats(r)["paths_sent_offrank"] for r in range(2)) > 0


def test_obj_round_trip_keeps_the_three_index_streams(oracle, tmp_path):
    """Wavefront OBJ in and out (the mesh format behind the reference's scene objects): positions / normals / texture
    coordinates keep their own index arrays, groups map to materials, polygons are fanned, relative indices resolve."""
    g = dprt.real_scene.make_garden(W=1, clusters=3)
    ob = g["objects"][0]
    names = ["ground", "leaf", "rock", "water", "leaf2"]
    p = str(tmp_path / "garden.obj")
    dprt.real_scene.save_obj(p, ob.meshes, names)
    meshes, inv = dprt.real_scene.load_obj(p, {n: i for i, n in enumerate(names)})
    assert inv == names and len(meshes) == len(ob.meshes)
    ident = [(k, np.eye(4)[:3]) for k in range(len(ob.meshes))]
    a = oracle.flatten_instances(ob.meshes, ident)
    b = oracle.flatten_instances(meshes, ident)
    assert np.array_equal(_bits(a[0]), _bits(b[0])) and np.array_equal(_bits(a[1]), _bits(b[1])) and np.array_equal(a[3], b[3])
    assert np.array_equal(_bits(a[2]), _bits(b[2]))
    assert meshes[3]["texcoords"] is None and meshes[1]["texcoords"] is not None        # the rock has no vt, the bush does
    # hand-written file: quad fan, negative indices, missing normals, unknown material name
    q = tmp_path / "quad.obj"
    q.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nusemtl brick\nf -4/1 -3/2 -2/3 -1/4\n")
    m, inv = dprt.real_scene.load_obj(str(q))
    assert inv == ["brick"] and m[0]["indices"].tolist() == [[0, 1, 2], [0, 2, 3]] and m[0]["texcoord_indices"].tolist() == [[0, 1, 2], [0, 2, 3]]
    assert np.allclose(m[0]["normals"][m[0]["normal_indices"]], [0, 0, 1])


def test_texture_loader_flips_like_stbi(tmp_path):
    img = np.zeros((3, 2, 3), np.float32); img[0] = 1.0          # top row white
    dprt.scene.save_pfm(str(tmp_path / "t.pfm"), img)
    t = dprt.real_scene.load_texture(str(tmp_path / "t.pfm"))
    assert t.shape == (3, 2, 4) and (t[..., 3] == 1).all()
    assert (t[2, :, :3] == 1).all() and (t[0, :, :3] == 0).all()  # row 0 = v 0 = the image's bottom row


def _read_scene_v2(path):
    """Minimal reader of the DPRTSCN2 layout documented in csrc/dprt_render.cpp (the C++ host has the real one)."""
    import ctypes as C
    b = open(path, "rb").read()
    assert b[:8] == b"DPRTSCN2"
    nobj, nmat, nlight, ntex = np.frombuffer(b, "<i4", 4, 8)
    o = 24 + C.sizeof(D.Camera)
    mats = np.frombuffer(b, D.MATERIAL_DTYPE, nmat, o); o += mats.nbytes
    mat_tex = np.frombuffer(b, "<i4", nmat, o); o += 4 * nmat
    lights = np.frombuffer(b, D.LIGHT_DTYPE, nlight, o); o += lights.nbytes
    textures = {}
    for _ in range(ntex):
        slot, w, h = np.frombuffer(b, "<i4", 3, o); o += 12
        textures[int(slot)] = np.frombuffer(b, "<f4", 4 * w * h, o).reshape(h, w, 4); o += 16 * w * h
    ew, eh = np.frombuffer(b, "<i4", 2, o); o += 8
    rot = float(np.frombuffer(b, "<f4", 1, o)[0]); o += 4
    env = None
    if ew:
        env = np.frombuffer(b, "<f4", 4 * ew * eh, o).reshape(eh, ew, 4); o += 16 * ew * eh
    objs = []
    for _ in range(nobj):
        o += C.sizeof(D.ObjectDesc)
        nt = int(np.frombuffer(b, "<i8", 1, o)[0]); o += 8
        has_uv = int(np.frombuffer(b, "<i4", 1, o)[0]); o += 4
        v = np.frombuffer(b, "<f4", 9 * nt, o).reshape(nt, 9); o += 36 * nt
        o += 36 * nt
        uv = None
        if has_uv:
            uv = np.frombuffer(b, "<f4", 6 * nt, o).reshape(nt, 6); o += 24 * nt
        m = np.frombuffer(b, "<i4", nt, o); o += 4 * nt
        for _ in range(2):
            o += 8 + int(np.frombuffer(b, "<i8", 1, o)[0])
        objs.append((v, uv, m))
    assert o == len(b)
    return mats, mat_tex, lights, textures, env, rot, objs


def test_obj2scene_converter(tmp_path):
    """OBJ + MTL + texture image -> DPRTSCN2 scene file (host only): materials, BSDF types, texture slots, chunks by x-slab; the
    C++ host parses the result."""
    import importlib.util
    import subprocess
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("obj2scene", os.path.join(root, "pg2024-data-parallel-ray-tracing_b200", "obj2scene.py"))
    o2s = importlib.util.module_from_spec(spec); spec.loader.exec_module(o2s)
    g = dprt.real_scene.make_garden(W=1, clusters=3, ground=(12, 12))
    names = ["ground", "leaf", "rock", "water", "leaf2"]
    obj = tmp_path / "garden.obj"
    dprt.real_scene.save_obj(str(obj), g["objects"][0].meshes, names)
    obj.write_text("mtllib garden.mtl\n" + obj.read_text())
    (tmp_path / "garden.mtl").write_text(
        "newmtl ground\nKd 0.5 0.5 0.5\nmap_Kd checker.pfm\nnewmtl leaf\nKd 0.1 0.6 0.1\nmap_Kd checker.pfm\n"
        "newmtl rock\nKd 0.55 0.5 0.45\nnewmtl water\nNi 1.33\nillum 7\nnewmtl leaf2\nKd 0.2 0.7 0.2\n")
    dprt.scene.save_pfm(str(tmp_path / "checker.pfm"), dprt.real_scene.checker_texture(8)[..., :3])
    dprt.scene.save_pfm(str(tmp_path / "sky.pfm"), dprt.real_scene.sky_env_map(16, 8)[..., :3])
    out = str(tmp_path / "garden.dprt")
    info = o2s.convert(str(obj), out, world=2, width=64, height=36, env=str(tmp_path / "sky.pfm"), env_rotation=0.5)
    ntris = sum(np.asarray(m["indices"]).reshape(-1, 3).shape[0] for m in g["objects"][0].meshes)
    order = ["ground", "leaf", "leaf2", "rock", "water"]           # material indices follow first use in the OBJ
    assert info["triangles"] == ntris and info["materials"] == order and sum(info["chunks"]) == ntris and len(info["chunks"]) == 2
    mats, mat_tex, lights, textures, env, rot, objs = _read_scene_v2(out)
    assert mat_tex.tolist() == [0, 0, -1, -1, -1] and list(textures) == [0] and textures[0].shape == (8, 8, 4)
    assert mats["bsdfType"].tolist() == [0, 0, 0, 0, 1] and np.allclose(mats["baseColor"][3], (0.55, 0.5, 0.45))
    assert env.shape == (8, 16, 4) and abs(rot - 0.5) < 1e-7 and lights.size == 2
    # row 0 of a loaded texture is the image's bottom row (stbi flip): save_pfm wrote the checker top-down
    assert np.array_equal(textures[0][::-1, :, :3], dprt.real_scene.checker_texture(8)[..., :3])
    assert sum(o[0].shape[0] for o in objs) == ntris and all(o[1] is not None for o in objs)
    assert objs[0][0].reshape(-1, 3, 3)[:, :, 0].mean(1).max() <= objs[1][0].reshape(-1, 3, 3)[:, :, 0].mean(1).min() + 1e-6      # x-slabs
    # light normals face the scene (down along `up`)
    for L in lights:
        assert np.cross(L["p1"] - L["p0"], L["p2"] - L["p0"])[2] < 0 and L["p0"][2] > objs[0][0].reshape(-1, 3)[:, 2].max()
    binp = os.path.join(os.path.dirname(dprt.host.LIB_PATH), "dprt_render")
    if not torch.cuda.is_available():
        p = subprocess.run([binp, "--scene", out, "--world", "2"], capture_output=True, text=True)
        assert p.returncode == 1 and "no CUDA device" in p.stderr and "malformed" not in p.stderr
