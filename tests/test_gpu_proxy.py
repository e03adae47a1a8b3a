"""GPU parity of the neural-proxy branch of the per-bounce loop (-m gpu): ShadowRay / SecondaryRay proxy-AABB
march, Work_Efficient_Scan_For_NN[_HIT_INSIDE], the batched proxy forward and the three NN epilogues, all called
through the C ABI and compared with the oracle.

Thresholds (fixed before any GPU number, BASELINE.json north_star):
  * queries, features, bucket offsets, packed order, epilogue outputs: bit-exact;
  * proxy predictions: |gpu - oracle| <= 1e-3 abs on the random-init network (fp16 operands), 2e-2 on the
    "spread" network whose outputs are ~40x that scale (tests/test_gpu_mlp.py);
  * because a prediction inside that band around 0.5 may flip a decision, the stage-wise test copies the GPU's
    predictions into the oracle after each forward, so that every later stage is compared bit-exactly on equal
    inputs; the end-to-end test bounds the image instead: relative MSE <= 1e-2 at 1 spp, and the two images must
    be bit-identical when the predictions are nowhere near a threshold (random-init network: all outputs < 0.5).
"""
import numpy as np
import pytest

from helpers import D, assert_bits_equal, assert_records_equal, build_pair, dprt

pytestmark = pytest.mark.gpu


def _models(W, spread, nres=4):
    import torch
    out = {}
    for i in range(W):
        torch.manual_seed(19990201 + i)
        vis = dprt.proxy.make_proxy(256, nres).eval()
        dep = dprt.proxy.make_proxy(256, nres).eval()
        if spread:
            dprt.proxy.spread_output_(vis, gain=3.0, seed=1 + i)
            dprt.proxy.spread_output_(dep, gain=1.5, seed=11 + i)
        out[i] = (dprt.proxy.pack_module(vis), dprt.proxy.pack_module(dep))
    return out


def _compare_queries(R, world, r, nslots, what):
    """Unpacked query slots: the key (hitAABBID) of every slot, the whole record and feature row of live slots."""
    qg, qo = R.download(D.BUF_NN_QUERY, nslots), world.download(r, D.BUF_NN_QUERY, nslots)
    assert_bits_equal(qg["hitAABBID"], qo["hitAABBID"], f"{what}: query keys")
    live = qo["hitAABBID"] != 0
    assert_records_equal(qg[live], qo[live], f"{what}: live query records")
    xg = R.download(D.BUF_NN_INPUT, nslots * 5).reshape(-1, 5)
    xo = world.download(r, D.BUF_NN_INPUT, nslots * 5).reshape(-1, 5)
    assert_bits_equal(xg[live], xo[live], f"{what}: fp16 features")
    return int(live.sum())


def _bucket_and_compare(R, world, r, which, inside, S, what):
    tg, to = R.bucket_queries(which, inside), world.bucket_queries(r, which, inside)
    assert tg == to, f"{what}: total {tg} vs {to}"
    assert_bits_equal(R.download(D.BUF_SCENE_OFFSET, S + 1), world.download(r, D.BUF_SCENE_OFFSET, S + 1), f"{what}: sceneOffset")
    assert_records_equal(R.download(D.BUF_NN_PACKED_QUERY, tg), world.download(r, D.BUF_NN_PACKED_QUERY, to), f"{what}: packed queries")
    assert_bits_equal(R.download(D.BUF_NN_PACKED_INPUT, tg * 5), world.download(r, D.BUF_NN_PACKED_INPUT, to * 5), f"{what}: packed features")
    return tg


def _infer_and_sync(R, world, r, kind, off, total, tol, what):
    R.proxy_infer(kind, off); world.proxy_infer(r, kind, off)
    pg = R.download(D.BUF_PRED, total, offset=off)
    po = world.download(r, D.BUF_PRED, total, offset=off)
    if total:
        err = np.abs(pg.view(np.float16).astype(np.float32) - po.view(np.float16).astype(np.float32)).max()
        assert err <= tol, f"{what}: max |gpu - oracle| = {err}"
    world.upload(r, D.BUF_PRED, pg, offset=off)      # equal inputs for the bit-exact epilogue comparison
    return pg


@pytest.mark.parametrize("spread,tol,W,w,h,tris", [(False, 1e-3, 2, 96, 54, 4000), (True, 2e-2, 2, 96, 54, 4000),
                                                   (False, 1e-3, 4, 320, 180, 40000)])      # 4 owners, 3 proxies each, ~10^5 queries per stage
def test_proxy_branch_stagewise(gpu_required, oracle, spread, tol, W, w, h, tris):
    oracle.use_all_host_threads()
    bounces = 2
    rs, world, _ = build_pair(oracle, W, tris, w, h, bounces=bounces, proxy_mode=1, models=_models(W, spread), mlp_dtype=1)
    G = dprt.RankGroup(rs)
    N, spc, mc, S = w * h, rs[0].cfg.shadowPathCount, rs[0].cfg.maxCount, W
    for R in rs:
        R.reset_frame()
    world.reset_frame()
    for R in rs:
        R.begin_sample(0)
    world.begin_sample(0)
    for r, R in enumerate(rs):
        R.path_gen(); world.path_gen(r)
    n_queries = {"shadow": 0, "shadow_inside": 0, "secondary": 0}
    for bounce in range(bounces + 1):
        if bounce > 0:
            for r, R in enumerate(rs):
                R.reset_nn(); world.reset_nn(r)
                R.secondary_trace(); world.secondary_trace(r)
                n = R.path_size
                assert n == world.path_size(r)
                assert_records_equal(R.download(D.BUF_PATHS, n), world.download(r, D.BUF_PATHS, n), f"b{bounce} r{r} paths after secondary trace")
                assert_bits_equal(R.download(D.BUF_ENV), world.download(r, D.BUF_ENV, 3 * N), f"b{bounce} r{r} env after secondary trace")
                _compare_queries(R, world, r, mc * n, f"b{bounce} r{r} secondary")
                total = _bucket_and_compare(R, world, r, 1, False, S, f"b{bounce} r{r} secondary bucket")
                n_queries["secondary"] += total
                _infer_and_sync(R, world, r, 0, 0, total, tol, f"b{bounce} r{r} secondary vis")
                _infer_and_sync(R, world, r, 1, total, total, tol, f"b{bounce} r{r} secondary depth")
                R.target_node_update(); world.target_node_update(r)
                assert_records_equal(R.download(D.BUF_PATHS, n), world.download(r, D.BUF_PATHS, n), f"b{bounce} r{r} paths after target node update")
        while True:
            for r, R in enumerate(rs):
                R.traverse(); world.traverse(r)
                R.partition(); world.partition(r)
            dg, do = G.exchange(), world.exchange()
            assert dg == do
            if dg:
                break
        for r, R in enumerate(rs):
            n = R.path_size
            assert n == world.path_size(r)
            R.shade(); world.shade(r)
            R.reset_nn(); world.reset_nn(r)
            R.shadow_trace(); world.shadow_trace(r)
            assert_records_equal(R.download(D.BUF_PATHS, n * (1 + spc)), world.download(r, D.BUF_PATHS, n * (1 + spc)), f"b{bounce} r{r} paths after shadow trace")
            assert_bits_equal(R.download(D.BUF_DIRECT), world.download(r, D.BUF_DIRECT, 3 * N * spc), f"b{bounce} r{r} direct after shadow trace")
            _compare_queries(R, world, r, mc * spc * n, f"b{bounce} r{r} shadow")
            tin = _bucket_and_compare(R, world, r, 0, True, S, f"b{bounce} r{r} shadow inside bucket")
            n_queries["shadow_inside"] += tin
            _infer_and_sync(R, world, r, 1, 0, tin, tol, f"b{bounce} r{r} shadow depth")
            R.depth_buffer_update(); world.depth_buffer_update(r)
            qg, qo = R.download(D.BUF_NN_QUERY, mc * spc * n), world.download(r, D.BUF_NN_QUERY, mc * spc * n)
            live = qo["hitAABBID"] != 0
            assert_bits_equal(qg["normalizedT"][live], qo["normalizedT"][live], f"b{bounce} r{r} normalizedT after depth update")
            tall = _bucket_and_compare(R, world, r, 0, False, S, f"b{bounce} r{r} shadow bucket")
            n_queries["shadow"] += tall
            _infer_and_sync(R, world, r, 0, 0, tall, tol, f"b{bounce} r{r} shadow vis")
            R.frame_buffer_update(); world.frame_buffer_update(r)
            assert_bits_equal(R.download(D.BUF_DIRECT), world.download(r, D.BUF_DIRECT, 3 * N * spc), f"b{bounce} r{r} direct after frame buffer update")
    print("queries compared:", n_queries)
    assert n_queries["shadow"] > 0 and n_queries["secondary"] > 0, "the scene does not exercise the proxy branch"
    assert_bits_equal(G.reduce_image(0), world.image(), "image after stage-wise run with synchronised predictions")


def test_proxy_image_random_init_bit_exact(gpu_required, oracle):
    """Random-init proxies predict ~0.03-0.09 everywhere (far below 0.5): no decision can flip, so the composite
    modules of both sides must produce the very same image."""
    W = 2
    rs, world, _ = build_pair(oracle, W, 6000, 128, 72, spp=2, bounces=2, proxy_mode=1, models=_models(W, False), mlp_dtype=1)
    img_g = dprt.RankGroup(rs).launch()
    img_o = world.launch()
    assert np.isfinite(img_g).all() and img_g.max() > 0
    assert_bits_equal(img_g, img_o, "proxy-on image, random-init networks")
    for r, R in enumerate(rs):
        sg, so = R.stats(), world.stats(r)
        for k in ("rays_traverse", "rays_shade", "rays_shadow", "rays_secondary", "nn_queries", "paths_sent_offrank"):
            assert sg[k] == so[k], (r, k, sg[k], so[k])
    assert sum(R.stats()["nn_queries"] for R in rs) > 0


@pytest.mark.parametrize("W,mlp_dtype", [(2, 1), (4, 0)])
def test_proxy_image_spread_networks_relmse(gpu_required, oracle, W, mlp_dtype):
    """Networks whose outputs straddle the thresholds: decisions inside the error band may flip; the image is
    bounded in relative MSE (stated bound 1e-2 at 1 spp; fp16 and bf16 operands)."""
    rs, world, _ = build_pair(oracle, W, 5000, 128, 72, spp=1, bounces=2, proxy_mode=1, models=_models(W, True), mlp_dtype=mlp_dtype)
    img_g = dprt.RankGroup(rs).launch()
    img_o = world.launch()
    assert np.isfinite(img_g).all()
    mse = float(np.mean((img_g.astype(np.float64) - img_o) ** 2))
    rel = mse / float(np.mean(img_o.astype(np.float64) ** 2))
    diff_px = float((np.abs(img_g - img_o).max(axis=2) > 0).mean())
    print(f"W={W} dtype={'bf16' if mlp_dtype == 0 else 'fp16'} relMSE={rel:.3e} differing pixels={diff_px:.4%}")
    assert rel <= 1e-2
    assert sum(R.stats()["nn_queries"] for R in rs) > 0


def test_proxy_stage_errors_without_proxy_mode(gpu_required):
    cfg = dprt.make_config(16, 16, scene_size=1, proxy_mode=0)
    R = dprt.Renderer(cfg)
    with pytest.raises(dprt.DprtError):
        R.secondary_trace()
    with pytest.raises(dprt.DprtError):
        R.bucket_queries(0, False)
