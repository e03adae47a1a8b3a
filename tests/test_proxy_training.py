"""Proxy training pipeline: the Vis-pipeline operator dprt_gen_train_data / its oracle, the
dataset rules of trainingcode/datasets.py and the training loop of trainingcode/main.py."""
import numpy as np
import pytest

from helpers import D, assert_bits_equal, build_pair, dprt

PT = dprt.proxy_train


def _world(oracle, tris=4000):
    chunks, mats, lights = dprt.scene.make_scene(1, tris)
    cfg = dprt.make_config(16, 16)
    world = oracle.World(cfg, 1)
    c = chunks[0]
    world.add_object(0, c.desc(False), c.verts, c.normals, c.mats)
    return world, c


def test_training_rays_start_on_the_box_and_point_inward():
    mn, mx = np.array([0.1, 0.2, 0.3]), np.array([0.9, 0.7, 0.5])
    rays = PT.sample_training_rays(mn, mx, 20000, seed=3)
    o, d = rays["origin"].astype(np.float64), rays["direction"].astype(np.float64)
    on_face = (np.isclose(o, mn, atol=1e-6) | np.isclose(o, mx, atol=1e-6)).any(1)
    assert on_face.all()
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-5)
    inside = ((o + 1e-3 * d) >= mn - 1e-7).all(1) & ((o + 1e-3 * d) <= mx + 1e-7).all(1)
    assert inside.mean() > 0.99                       # a step along the ray stays in the box (edge-grazing rays aside)
    # area-uniform: the two largest faces (z = const: 0.8 x 0.5) get the most samples
    zface = np.isclose(o[:, 2], mn[2], atol=1e-6) | np.isclose(o[:, 2], mx[2], atol=1e-6)
    assert 0.55 < zface.mean() < 0.75                 # 2*0.40 / (2*0.40 + 2*0.10 + 2*0.16) = 0.606


def test_oracle_train_data_encoding(oracle):
    world, c = _world(oracle)
    rays = PT.sample_training_rays(c.aabb_min, c.aabb_max, 20000, seed=1)
    feat, lab = world.gen_train_data(0, rays)
    assert feat.shape == (20000, 5) and np.isfinite(feat).all()
    assert feat.min() >= -1e-4 and feat.max() <= 1.0 + 1e-4           # (o - min)/(max - min), phi/2pi, theta/pi
    hits = world.trace_closest(0, rays)
    hit = hits["primID"] >= 0
    assert 0.2 < hit.mean() < 0.98
    assert (lab[~hit] == 1.0).all()                                     # label 1.0 == miss (vis_ray_kernel.cu:160)
    maxlen = np.float32(c.desc(False).maxLength)
    assert_bits_equal(lab[hit], (hits["t"][hit] / maxlen).astype(np.float32), "depth label = t / maxLength")
    xv, yv = PT.vis_dataset(feat, lab)
    nh = int(hit.sum())
    assert int(yv.sum()) == nh and yv.size == nh + min(int(1.5 * nh), int((~hit).sum()))   # radio = 1.5 (datasets.py:153)
    xd, yd = PT.depth_dataset(feat, lab)
    assert yd.size == nh and (yd < 1.0).all() and (yd > 0).all()


def _precom_rays(c, n_side=160):
    """Camera paths as the Precom pipeline receives them (precom_ray_kernel.cu:205-212): tMin 1e-2, tMax = path.tMax = FLT_MAX;
    the benchmark camera (outside every chunk's AABB) plus a camera INSIDE the AABB, so both AABB faces occur."""
    cam = dprt.scene.default_camera(n_side, n_side * 9 // 16)
    a = dprt.scene.camera_rays(cam)
    b = a.copy()
    b["origin"] = (0.5 * (np.asarray(c.aabb_min) + np.asarray(c.aabb_max)) + np.float32([0.0, 0.0, 0.2 * (c.aabb_max[2] - c.aabb_min[2])])).astype(np.float32)
    rays = np.concatenate([a, b])
    rays["tMin"] = np.float32(1e-2)
    rays["tMax"] = np.finfo(np.float32).max
    return rays


def test_oracle_precom_data_encoding(oracle):
    """Precom pipeline (optix/precom_ray_kernel.cu:193-299): features at the proxy-AABB hit, label = geometry depth behind it."""
    world, c = _world(oracle)
    rays = _precom_rays(c)
    feat, lab, valid = world.gen_precom_data(0, rays)
    v = valid.astype(bool)
    assert 0.2 < v.mean() <= 1.0 and np.isfinite(feat).all() and np.isfinite(lab).all()
    assert (feat[~v] == 0).all() and (lab[~v] == 1.0).all()
    # the AABB hit point lies ON the box: one normalised coordinate is 0 or 1 (fp32 rounding of the transform aside)
    on_face = (np.abs(feat[v, :3]) < 1e-4) | (np.abs(feat[v, :3] - 1.0) < 1e-4)
    assert on_face.any(axis=1).all()
    assert feat[v, 3:].min() >= -1e-6 and feat[v, 3:].max() <= 1.0 + 1e-6
    # label: geometry hit -> (t_geo - t_aabb) / maxLength, never negative for a ray that enters from outside
    hits = world.trace_closest(0, rays)
    geo = hits["primID"] >= 0
    assert (lab[v & ~geo] == 1.0).all()
    n_out = rays.size // 2                                              # first half: camera outside the box
    assert (lab[:n_out][v[:n_out] & geo[:n_out]] >= 0).all() and (lab[v & geo] < 1.0).all()
    assert (v[n_out:]).all()                                            # a camera inside the box always meets its back face


def test_training_loop_learns_visibility(oracle):
    """A small trunk on 60 k oracle samples: the test loss must fall well below the constant predictor's."""
    import torch
    torch.set_num_threads(4)
    world, c = _world(oracle)
    rays = PT.sample_training_rays(c.aabb_min, c.aabb_max, 60000, seed=2)
    feat, lab = world.gen_train_data(0, rays)
    xv, yv = PT.vis_dataset(feat, lab)
    model, hist = PT.train_proxy(xv, yv, "vis", width=64, nres=2, epochs=12, lr=2e-3, batch=2048, device="cpu")
    base = float(yv.mean() * (1 - yv.mean()))                           # MSE of predicting the mean
    assert hist[-1] < 0.6 * base, (hist, base)
    blob = dprt.proxy.pack_module(model)
    assert len(blob) > 16 and dprt.proxy.unpack_blob(blob) is not None


@pytest.mark.gpu
def test_gpu_precom_data_matches_oracle(gpu_required, oracle):
    rs, world, chunks = build_pair(oracle, 2, 20000, 32, 18, proxy_mode=0)
    for r, R in enumerate(rs):
        c = chunks[r]
        rays = _precom_rays(c, 320)
        fg, lg, vg = R.gen_precom_data(c.index, rays)
        fo, lo, vo = world.gen_precom_data(c.index, rays)
        assert_bits_equal(vg, vo, f"valid flags of chunk {r}")
        assert_bits_equal(fg, fo, f"Precom features of chunk {r}")
        assert_bits_equal(lg, lo, f"Precom labels of chunk {r}")
        assert vg.mean() > 0.2 and ((lg != 1.0) & (vg == 1)).mean() > 0.05
        with pytest.raises(dprt.DprtError):
            R.gen_precom_data(chunks[1 - r].index, rays[:8])
        f0, l0, v0 = R.gen_precom_data(c.index, rays[:0])
        assert f0.shape == (0, 5) and l0.shape == (0,) and v0.shape == (0,)


@pytest.mark.gpu
def test_gpu_train_data_matches_oracle(gpu_required, oracle):
    rs, world, chunks = build_pair(oracle, 2, 20000, 32, 18, proxy_mode=0)
    for r, R in enumerate(rs):
        c = chunks[r]
        rays = PT.sample_training_rays(c.aabb_min, c.aabb_max, 50000, seed=10 + r)
        fg, lg = R.gen_train_data(c.index, rays)
        fo, lo = world.gen_train_data(c.index, rays)
        assert_bits_equal(fg, fo, f"features of chunk {r}")
        assert_bits_equal(lg, lo, f"labels of chunk {r}")
        assert 0.1 < (lg != 1.0).mean() < 0.99
        with pytest.raises(dprt.DprtError):
            R.gen_train_data(chunks[1 - r].index, rays[:8])             # a proxy on this rank: no geometry here
        f0, l0 = R.gen_train_data(c.index, rays[:0])
        assert f0.shape == (0, 5) and l0.shape == (0,)


@pytest.mark.gpu
def test_gpu_trained_proxies_beat_untrained_ones(gpu_required):
    """The whole loop on the GPU box: dprt_gen_train_data -> proxy_train (torch, cuda) -> blob -> proxy-on render. With
    trained vis/depth proxies the image must be closer to the exact (proxy-off, ray-migrating) one than with random-init
    networks (which predict "nothing there" everywhere)."""
    import torch
    W, w, h = 2, 192, 108
    chunks, mats, lights = dprt.scene.make_scene(W, 20000)
    cam = dprt.scene.default_camera(w, h)

    def render(proxy_mode, blobs):
        cfg = dprt.make_config(w, h, spp=4, bounces=2, scene_size=W, proxy_mode=proxy_mode, path_gen_mode=1, mlp_dtype=1)
        rs = []
        for r in range(W):
            R = dprt.Renderer(cfg, rank=r, world=W, device=0)
            for c in chunks:
                if c.node_id == r:
                    R.upload_chunk(c.index, c.desc(False), c.verts, c.normals, c.mats)
                else:
                    R.upload_proxy(c.index, c.desc(True), *blobs.get(c.index, (None, None)))
            R.set_materials(mats); R.set_lights(lights); R.set_camera(cam)
            rs.append(R)
        img = dprt.RankGroup(rs).launch()
        nnq = sum(R.stats()["nn_queries"] for R in rs)
        for R in rs:
            R.close()
        return img, nnq

    exact, _ = render(0, {})
    untrained = {}
    for k in range(W):
        torch.manual_seed(19990201 + k)
        m = dprt.proxy.make_proxy(256, 4).eval()
        untrained[k] = (dprt.proxy.pack_module(m), dprt.proxy.pack_module(m))
    img_u, _ = render(1, untrained)
    trained = {}
    for c in chunks:
        R = dprt.Renderer(dprt.make_config(16, 16, scene_size=W), rank=c.node_id, world=W, device=0)
        R.upload_chunk(c.index, c.desc(False), c.verts, c.normals, c.mats)
        vb, db, info = PT.train_chunk_proxies(lambda rays: R.gen_train_data(c.index, rays), c.aabb_min, c.aabb_max, n_rays=200000,
                                              epochs=10, seed=c.index, device="cuda")
        R.close()
        assert info["vis_test_loss"][-1] < 0.6 * info["vis_test_loss"][0]
        trained[c.index] = (vb, db)
    img_t, nnq = render(1, trained)
    assert nnq > 0

    def relmse(a, b):
        return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))
    eu, et = relmse(img_u, exact), relmse(img_t, exact)
    assert et < 0.8 * eu, (et, eu)
