"""Edge cases of the hot path on the GPU (-m gpu), each against the oracle: degenerate geometry and rays, watertight
edges, empty wavefronts, extreme configurations, argument errors."""
import numpy as np
import pytest

from helpers import D, assert_bits_equal, assert_records_equal, build_pair, dprt

pytestmark = pytest.mark.gpu


def _single_object_pair(oracle, verts, w=32, h=18):
    verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 9)
    n = verts.shape[0]
    normals = np.tile(np.array([0, 0, 1] * 3, np.float32), (n, 1))
    mats = np.zeros(n, np.int32)
    mn, mx = verts.reshape(-1, 3).min(0) - 1e-3, verts.reshape(-1, 3).max(0) + 1e-3
    desc = dprt.make_object_desc(0, mn, mx)
    cfg = dprt.make_config(w, h)
    R = dprt.Renderer(cfg)
    R.upload_chunk(0, desc, verts, normals, mats)
    world = oracle.World(cfg, 1)
    world.add_object(0, desc, verts, normals, mats)
    return R, world


def _rays(o, d, tmin=1e-3, tmax=np.finfo(np.float32).max):
    o, d = np.asarray(o, np.float32).reshape(-1, 3), np.asarray(d, np.float32).reshape(-1, 3)
    r = np.zeros(o.shape[0], D.RAY_DTYPE)
    r["origin"], r["direction"], r["tMin"], r["tMax"] = o, d, tmin, tmax
    return r


def test_watertight_grid_edges_and_vertices(gpu_required, oracle):
    """Rays aimed exactly at the shared edges and vertices of a regular grid: every ray hits (no cracks), and the
    primitive chosen among the triangles that share the edge is the oracle's (lowest id on a tie in t)."""
    n = 16
    xs = np.arange(n + 1, dtype=np.float32) / n
    tris = []
    for i in range(n):
        for j in range(n):
            p00, p10, p01, p11 = (xs[i], xs[j], 0), (xs[i + 1], xs[j], 0), (xs[i], xs[j + 1], 0), (xs[i + 1], xs[j + 1], 0)
            tris += [p00 + p10 + p11, p00 + p11 + p01]
    R, world = _single_object_pair(oracle, np.array(tris, np.float32))
    gx, gy = np.meshgrid(xs[1:-1], xs[1:-1], indexing="ij")                      # interior vertices
    mid = (xs[:-1] + xs[1:]) / 2
    ex, ey = np.meshgrid(xs[1:-1], mid, indexing="ij")                           # points on vertical edges
    dx, dy = np.meshgrid(mid, mid, indexing="ij")                                # points on the quad diagonals
    px = np.concatenate([gx.ravel(), ex.ravel(), ey.ravel(), dx.ravel()])
    py = np.concatenate([gy.ravel(), ey.ravel(), ex.ravel(), dy.ravel()])
    o = np.stack([px, py, np.full_like(px, 1.0)], 1)
    for d in ([0, 0, -1], [0.0, 0.0, -3.0]):                                     # axis-aligned: two zero direction components
        rays = _rays(o, np.tile(np.array(d, np.float32), (o.shape[0], 1)))
        hg, ho = R.trace_closest(rays), world.trace_closest(0, rays)
        assert (hg["primID"] >= 0).all(), "a ray slipped through a shared edge or vertex"
        assert_bits_equal(hg["primID"], ho["primID"], "primitive on shared edges / vertices")
        assert_bits_equal(hg["t"], ho["t"], "t on shared edges / vertices")
    # oblique rays through the same points
    tgt = np.stack([px, py, np.zeros_like(px)], 1)
    src = np.array([0.3, -0.7, 0.9], np.float32)
    dirs = tgt - src
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    rays = _rays(np.tile(src, (o.shape[0], 1)), dirs)
    hg, ho = R.trace_closest(rays), world.trace_closest(0, rays)
    assert_bits_equal(hg["primID"], ho["primID"], "oblique rays at edges: primitive")
    assert_bits_equal(hg["t"], ho["t"], "oblique rays at edges: t")
    R.close()


def test_degenerate_triangles_duplicates_and_ray_ranges(gpu_required, oracle):
    """Zero-area triangles never hit; coincident duplicates resolve to the lowest primitive id; empty and inverted
    [tMin, tMax] ranges and far-away origins behave like the oracle."""
    a, b, c = (0, 0, 0), (1, 0, 0), (0, 1, 0)
    tris = [a + a + a,                       # point
            a + b + b,                       # segment
            a + b + c, a + b + c, a + b + c,  # three coincident copies (ids 2, 3, 4)
            (0, 0, -1) + (1, 0, -1) + (0, 1, -1)]
    R, world = _single_object_pair(oracle, np.array(tris, np.float32))
    o = np.array([[0.2, 0.2, 1.0]] * 6 + [[0.2, 0.2, 1e6], [5.0, 5.0, 1.0]], np.float32)
    d = np.array([[0, 0, -1]] * 8, np.float32)
    rays = _rays(o, d)
    rays["tMin"][1], rays["tMax"][1] = 0.5, 0.9          # range ends before the first surface
    rays["tMin"][2], rays["tMax"][2] = 1.5, 3.0          # range starts behind the first surface: the z = -1 triangle
    rays["tMin"][3], rays["tMax"][3] = 2.5, 1.0          # inverted
    rays["tMin"][4], rays["tMax"][4] = 1.0, 1.0          # empty at exactly the hit distance
    rays["tMin"][5], rays["tMax"][5] = 0.0, np.nextafter(np.float32(1.0), np.float32(2.0))
    hg, ho = R.trace_closest(rays), world.trace_closest(0, rays)
    assert_bits_equal(hg["primID"], ho["primID"], "primitive ids")
    assert_bits_equal(hg["t"], ho["t"], "t")
    assert hg["primID"][0] == 2 and hg["primID"][1] == -1 and hg["primID"][2] == 5 and hg["primID"][3] == -1 and hg["primID"][7] == -1
    R.close()


def test_all_rays_miss_empty_wavefronts(gpu_required, oracle):
    """A camera that looks away from the scene: every path dies in the first TraRay, all later stages run on empty
    wavefronts (n = 0 launches, empty partition, empty exchange) and the image is the environment alone."""
    for W in (1, 2):
        chunks, mats, lights = dprt.scene.make_scene(W, 3000)
        cam = dprt.make_camera((0.5, 0.5, 3.0), (0.5, 0.5, 9.0), (0.0, 1.0, 0.0), 30.0, 64, 36)   # looking up, away from the cube
        cfg = dprt.make_config(64, 36, spp=2, bounces=3, scene_size=W, path_gen_mode=1 if W > 1 else 0)
        world = oracle.World(cfg, W)
        rs = []
        for c in chunks:
            world.add_object(c.index, c.desc(False), c.verts, c.normals, c.mats)
        world.set_materials(mats); world.set_lights(lights); world.set_camera(cam)
        for r in range(W):
            R = dprt.Renderer(cfg, rank=r, world=W)
            for c in chunks:
                if c.node_id == r:
                    R.upload_chunk(c.index, c.desc(False), c.verts, c.normals, c.mats)
                else:
                    R.upload_proxy(c.index, c.desc(True), None, None)
            R.set_materials(mats); R.set_lights(lights); R.set_camera(cam)
            rs.append(R)
        img_g = rs[0].launch() if W == 1 else dprt.RankGroup(rs).launch()
        img_o = world.launch()
        assert_bits_equal(img_g, img_o, f"environment-only image, W={W}")
        assert img_g.min() > 0
        for r, R in enumerate(rs):
            assert R.path_size == 0 == world.path_size(r)
            assert R.stats()["rays_shade"] == 0 and R.stats()["rays_shadow"] == 0
            R.close()


@pytest.mark.parametrize("spc,mc,W,proxy", [(1, 1, 1, 0), (7, 2, 2, 0), (2, 5, 2, 1), (1, 3, 2, 1)])
def test_shadow_path_and_march_counts_other_than_default(gpu_required, oracle, spc, mc, W, proxy):
    """renderer.cpp:1602-1603 fixes shadowPathCount = 4 and maxCount = 3; the buffers and kernels are sized by them."""
    import torch
    chunks, mats, lights = dprt.scene.make_scene(W, 5000)
    cam = dprt.scene.default_camera(96, 54)
    cfg = dprt.make_config(96, 54, spp=2, bounces=2, spc=spc, mc=mc, scene_size=W, proxy_mode=proxy, path_gen_mode=1 if W > 1 else 0, mlp_dtype=1)
    blobs = {}
    if proxy:
        for k in range(W):
            torch.manual_seed(19990201 + k)
            b = dprt.proxy.pack_module(dprt.proxy.make_proxy(256, 4).eval())      # random init: predicts ~0.05, no decision near 0.5
            blobs[k] = (b, b)
    world = oracle.World(cfg, W)
    for c in chunks:
        world.add_object(c.index, c.desc(False), c.verts, c.normals, c.mats)
        if c.index in blobs:
            world.set_model(c.index, 0, blobs[c.index][0]); world.set_model(c.index, 1, blobs[c.index][1])
    world.set_materials(mats); world.set_lights(lights); world.set_camera(cam)
    rs = []
    for r in range(W):
        R = dprt.Renderer(cfg, rank=r, world=W)
        for c in chunks:
            if c.node_id == r:
                R.upload_chunk(c.index, c.desc(False), c.verts, c.normals, c.mats)
            else:
                R.upload_proxy(c.index, c.desc(True), *blobs.get(c.index, (None, None)))
        R.set_materials(mats); R.set_lights(lights); R.set_camera(cam)
        rs.append(R)
    img_g = rs[0].launch() if W == 1 else dprt.RankGroup(rs).launch()
    img_o = world.launch()
    assert_bits_equal(img_g, img_o, f"image spc={spc} mc={mc} W={W} proxy={proxy}")
    for r, R in enumerate(rs):
        n = R.path_size
        assert n == world.path_size(r)
        assert_records_equal(R.download(D.BUF_PATHS, n * (1 + spc)), world.download(r, D.BUF_PATHS, n * (1 + spc)), f"rank {r} paths")
        R.close()


def test_sixteen_chunk_owners(gpu_required, oracle):
    """visitedMask is a 32-bit set (distributed_traversal_kernel.cu:29-31): 16 owners in x-slabs, long migration chains."""
    W = 16
    rs, world, _ = build_pair(oracle, W, 1500, 96, 54, spp=1, bounces=1, proxy_mode=0, path_gen_mode=1)
    img_g = dprt.RankGroup(rs).launch()
    img_o = world.launch()
    assert_bits_equal(img_g, img_o, "16-rank image")
    hops = sum(R.stats()["paths_sent_offrank"] for R in rs)
    iters = rs[0].stats()["exchange_iters"]
    assert hops > 0 and iters == world.stats(0)["exchange_iters"] and iters >= 3
    for r, R in enumerate(rs):
        n = R.path_size
        assert n == world.path_size(r)
        assert_records_equal(R.download(D.BUF_PATHS, n * 5), world.download(r, D.BUF_PATHS, n * 5), f"rank {r} paths")
        R.close()


def test_argument_and_state_errors(gpu_required):
    """Every entry returns a negative dprt_error and a message instead of aborting (SURVEY.md 8b error convention)."""
    with pytest.raises(dprt.DprtError):
        dprt.Renderer(dprt.make_config(0, 10))                       # invalid frame
    with pytest.raises(dprt.DprtError):
        dprt.Renderer(dprt.make_config(8, 8, spc=99))                # shadowPathCount out of range
    with pytest.raises(dprt.DprtError):
        dprt.Renderer(dprt.make_config(8, 8), rank=3, world=2)       # rank outside the world
    R = dprt.Renderer(dprt.make_config(16, 16))
    chunks, mats, lights = dprt.scene.make_scene(1, 500)
    c = chunks[0]
    R.upload_chunk(0, c.desc(False), c.verts, c.normals, c.mats)
    R.set_materials(mats); R.set_camera(dprt.scene.default_camera(16, 16))
    R.reset_frame(); R.begin_sample(0); R.path_gen(); R.traverse(); R.partition(); R.exchange()
    with pytest.raises(dprt.DprtError) as e:
        R.shade()                                                     # no lights set
    assert "light" in str(e.value)
    with pytest.raises(dprt.DprtError):
        R.upload(D.BUF_PATHS, np.zeros(16 * 16 * 5 + 1, D.PATH_DTYPE))   # one record beyond pathDataBuffer
    with pytest.raises(dprt.DprtError):
        R.secondary_trace()                                           # needs proxyMode = 1
    R.set_lights(lights)
    R.shade()                                                         # the context is still usable after the errors
    R.close()
