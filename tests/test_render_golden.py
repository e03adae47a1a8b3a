"""Frozen oracle outputs (tests/golden/render_golden.npz, written by tests/golden/make_render_golden.py): the oracle must
still reproduce them bit for bit (CPU), and so must libdprt through the C ABI (-m gpu)."""
import importlib.util
import os
import zlib

import numpy as np
import pytest

from helpers import D, assert_bits_equal, dprt

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_render_golden", os.path.join(HERE, "golden", "make_render_golden.py"))
G = importlib.util.module_from_spec(spec)
spec.loader.exec_module(G)


def _golden(chunks):
    g = np.load(os.path.join(HERE, "golden", "render_golden.npz"))
    crc = np.array([zlib.crc32(np.ascontiguousarray(c.verts).tobytes()) for c in chunks], np.uint32)
    if not np.array_equal(crc, g["scene_crc"]):
        pytest.skip("the numpy scene generator rounds differently on this platform: the fixture's inputs are not reproduced")
    return g


def test_oracle_reproduces_frozen_outputs(oracle):
    world, chunks, cam, cfg = G.build()
    g = _golden(chunks)
    assert_bits_equal(world.launch(), g["image"], "oracle image vs frozen fixture")
    hits = world.trace_closest(0, dprt.scene.camera_rays(cam))
    assert_bits_equal(hits["primID"], g["prim"], "primary-ray primitive ids")
    assert_bits_equal(hits["t"].view(np.uint32), g["t_bits"], "primary-ray t")
    feat, lab = world.gen_train_data(1, g["train_rays"].view(D.RAY_DTYPE))
    assert_bits_equal(feat, g["train_feat"], "training features")
    assert_bits_equal(lab, g["train_label"], "training labels")
    st = [world.stats(r) for r in range(G.W)]
    assert [world.path_size(r) for r in range(G.W)] == g["path_size"].tolist()
    assert [s["paths_sent_offrank"] for s in st] == g["sent"].tolist() and [s["exchange_iters"] for s in st] == g["iters"].tolist()
    assert [s["rays_walked"] for s in st] == g["walked"].tolist()


@pytest.mark.gpu
def test_libdprt_reproduces_frozen_outputs(gpu_required):
    chunks, mats, lights = dprt.scene.make_scene(G.W, G.TRIS, water_frac=0.03)
    g = _golden(chunks)
    cam = dprt.scene.default_camera(G.WIDTH, G.HEIGHT)
    cfg = dprt.make_config(G.WIDTH, G.HEIGHT, spp=G.SPP, bounces=G.BOUNCES, scene_size=G.W, proxy_mode=0, path_gen_mode=1)
    rs = []
    for r in range(G.W):
        R = dprt.Renderer(cfg, rank=r, world=G.W)
        for c in chunks:
            if c.node_id == r:
                R.upload_chunk(c.index, c.desc(False), c.verts, c.normals, c.mats)
            else:
                R.upload_proxy(c.index, c.desc(True), None, None)
        R.set_materials(mats); R.set_lights(lights); R.set_camera(cam)
        rs.append(R)
    assert_bits_equal(dprt.RankGroup(rs).launch(), g["image"], "libdprt image vs frozen fixture")
    hits = rs[0].trace_closest(dprt.scene.camera_rays(cam))
    assert_bits_equal(hits["primID"], g["prim"], "primary-ray primitive ids")
    assert_bits_equal(hits["t"].view(np.uint32), g["t_bits"], "primary-ray t")
    feat, lab = rs[1].gen_train_data(1, g["train_rays"].view(D.RAY_DTYPE))
    assert_bits_equal(feat, g["train_feat"], "training features")
    assert_bits_equal(lab, g["train_label"], "training labels")
    assert [R.path_size for R in rs] == g["path_size"].tolist()
    assert [R.stats()["paths_sent_offrank"] for R in rs] == g["sent"].tolist()
    # the oracle re-traces MainRay, libdprt answers it from the hit cache: walked + cached on this side == walked in the fixture
    assert [R.stats()["rays_walked"] + R.stats()["rays_shade_cached"] for R in rs] == g["walked"].tolist()
    for R in rs:
        R.close()


# ---- real-scene front end: the frozen garden (tests/golden/make_garden_golden.py) ---------------------------------
spec2 = importlib.util.spec_from_file_location("make_garden_golden", os.path.join(HERE, "golden", "make_garden_golden.py"))
GG = importlib.util.module_from_spec(spec2)
spec2.loader.exec_module(GG)


def _garden_golden(g):
    f = np.load(os.path.join(HERE, "golden", "garden_golden.npz"))
    if not np.array_equal(GG.scene_crc(g), f["scene_crc"]):
        pytest.skip("the numpy scene generator rounds differently on this platform: the fixture's inputs are not reproduced")
    return f


def test_oracle_reproduces_frozen_garden(oracle):
    world, g = GG.build()
    f = _garden_golden(g)
    assert_bits_equal(world.launch(), f["image"], "oracle garden image vs frozen fixture")
    hits = world.trace_closest(0, dprt.scene.camera_rays(dprt.scene.default_camera(GG.WIDTH, GG.HEIGHT)))
    assert_bits_equal(hits["primID"], f["prim"], "primary-ray primitive ids with cut-outs")
    assert_bits_equal(hits["t"].view(np.uint32), f["t_bits"], "primary-ray t")
    assert world.stats(0)["rays_walked"] == int(f["walked"][0])


@pytest.mark.gpu
def test_libdprt_reproduces_frozen_garden(gpu_required, oracle):
    from helpers import build_garden_pair
    rs, _, g = build_garden_pair(oracle, 1, GG.WIDTH, GG.HEIGHT, spp=GG.SPP, bounces=GG.BOUNCES)
    f = _garden_golden(g)
    R = rs[0]
    assert_bits_equal(R.launch(), f["image"], "libdprt garden image vs frozen fixture")
    hits = R.trace_closest(dprt.scene.camera_rays(dprt.scene.default_camera(GG.WIDTH, GG.HEIGHT)))
    assert_bits_equal(hits["primID"], f["prim"], "primary-ray primitive ids with cut-outs")
    assert_bits_equal(hits["t"].view(np.uint32), f["t_bits"], "primary-ray t")
    assert R.stats()["rays_walked"] + R.stats()["rays_shade_cached"] == int(f["walked"][0])
