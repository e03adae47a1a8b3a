"""GPU parity of the real-scene front end (-m gpu): instanced + indexed meshes uploaded through
dprt_upload_instanced_chunk, albedo / opacity maps with the alpha cut-out inside every trace (kernel.cu:311-359 and its copies
in the other pipelines), texture-mapped base colour in MainRay's closest-hit program (kernel.cu:251-281) and the lat-long
environment map (kernel.cu:28-48), through the C ABI against the oracle on the same procedural garden. Bit-exact like the
rest of the loop: both sides implement the binary32 texture specification of DESIGN.md section 4."""
import numpy as np
import pytest

from helpers import D, assert_bits_equal, assert_records_equal, build_garden_pair, dprt, random_rays
from test_gpu_parity import _stagewise_bounce

pytestmark = pytest.mark.gpu


def _garden_rays(n, seed):
    rays = random_rays(n, seed, lo=0.0, hi=1.0)
    rays["origin"][:, 2] = 0.9
    rays["direction"][:, 2] = -np.abs(rays["direction"][:, 2]) - 0.3
    rays["direction"] /= np.linalg.norm(rays["direction"], axis=1, keepdims=True)
    return rays


def test_garden_geometry_only_closest_hit(gpu_required, oracle):
    """Instancing + flatten + upload alone: no textures bound, so every leaf card is opaque."""
    rs, world, _ = build_garden_pair(oracle, 1, 64, 36, textures=False, env_map=False)
    rays = np.concatenate([_garden_rays(40000, 7), dprt.scene.camera_rays(dprt.scene.default_camera(320, 180))])
    hg, ho = rs[0].trace_closest(rays), world.trace_closest(0, rays)
    assert_bits_equal(hg["primID"], ho["primID"], "closest-hit primitive ids, instanced geometry")
    assert_bits_equal(hg["t"], ho["t"], "closest-hit t, instanced geometry")
    assert (hg["primID"] >= 0).mean() > 0.5


def test_garden_closest_hit_with_cutouts(gpu_required, oracle):
    rs, world, g = build_garden_pair(oracle, 1, 64, 36)
    R = rs[0]
    rays = np.concatenate([_garden_rays(60000, 11), dprt.scene.camera_rays(dprt.scene.default_camera(320, 180))])
    off = np.full(5, -1, np.int32)
    # material -> texture table all -1: texture coordinates are present, nothing is cut out
    R.set_material_textures(off); world.set_material_textures(off)
    hg2, ho2 = R.trace_closest(rays), world.trace_closest(0, rays)
    assert_bits_equal(hg2["primID"], ho2["primID"], "closest-hit primitive ids, textures off")
    # textures on: candidates on transparent texels are dropped inside the traversal
    R.set_material_textures(g["material_textures"]); world.set_material_textures(g["material_textures"])
    hg, ho = R.trace_closest(rays), world.trace_closest(0, rays)
    assert_bits_equal(hg["primID"], ho["primID"], "closest-hit primitive ids (alpha cut-outs inside the traversal)")
    assert_bits_equal(hg["t"], ho["t"], "closest-hit t")
    assert (hg2["primID"] != hg["primID"]).sum() > 500       # the cut-outs are exercised
    # removing the texture itself (slot emptied, material still points at it) is the same as no texture
    R.set_texture(dprt.real_scene.TEX_LEAF, None)
    hg3 = R.trace_closest(rays)
    assert_bits_equal(hg3["primID"], ho2["primID"], "closest-hit primitive ids, leaf texture removed")


def test_garden_single_rank_stagewise_bit_exact(gpu_required, oracle):
    w, h = 192, 108
    rs, world, _ = build_garden_pair(oracle, 1, w, h, bounces=2)
    R = rs[0]
    N, spc = w * h, R.cfg.shadowPathCount
    R.enable_hit_prim(True); world.enable_hit_prim(True)
    R.reset_frame(); world.reset_frame()
    R.begin_sample(0); world.begin_sample(0)
    R.path_gen(); world.path_gen(0)
    for _ in range(3):
        _stagewise_bounce(R, world, 0, N, spc)
    assert (world.download(0, D.BUF_ENV, 3 * N) > 0).any()


@pytest.mark.parametrize("retrace", [0, 1])
def test_garden_image_bit_exact(gpu_required, oracle, retrace):
    rs, world, _ = build_garden_pair(oracle, 1, 160, 90, spp=3, bounces=3, main_ray_retrace=retrace)
    img_g, img_o = rs[0].launch(), world.launch()
    assert np.isfinite(img_g).all() and img_g.max() > 0
    assert_bits_equal(img_g, img_o, f"garden image, mainRayRetrace={retrace}")
    sg, so = rs[0].stats(), world.stats(0)
    for k in ("rays_traverse", "rays_shade", "rays_shadow"):
        assert sg[k] == so[k], k


def test_garden_two_ranks_group_bit_exact(gpu_required, oracle):
    """Two instanced, textured scene objects on two ranks (one GPU, in-process group): migration + cut-outs + env map."""
    rs, world, _ = build_garden_pair(oracle, 2, 128, 72, spp=2, bounces=3)
    G = dprt.RankGroup(rs)
    img_g, img_o = G.launch(), world.launch()
    assert sum(R.stats()["paths_sent_offrank"] for R in rs) > 0
    for r, R in enumerate(rs):
        n = R.path_size
        assert n == world.path_size(r)
        assert_records_equal(R.download(D.BUF_PATHS, n), world.download(r, D.BUF_PATHS, n), f"rank {r} paths")
        assert_bits_equal(R.download(D.BUF_ENV), world.download(r, D.BUF_ENV, 3 * 128 * 72), f"rank {r} envLightingBuffer")
        assert_bits_equal(R.download(D.BUF_DIRECT, 3 * 128 * 72), world.download(r, D.BUF_DIRECT, 3 * 128 * 72), f"rank {r} directLightingBuffer")
    assert np.allclose(img_g, img_o, rtol=1e-6, atol=1e-7)      # the reduce sums ranks in another order than the oracle's loop


def test_garden_training_rays_see_cutouts(gpu_required, oracle):
    """The Vis / Precom generators run the same any-hit program (vis_ray_kernel.cu:33-81, precom_ray_kernel.cu:33-81)."""
    rs, world, _ = build_garden_pair(oracle, 1, 64, 36)
    rays = _garden_rays(30000, 23)
    rays["tMin"] = 1e-5
    fg, lg = rs[0].gen_train_data(0, rays)
    fo, lo = world.gen_train_data(0, rays)
    assert_bits_equal(fg, fo, "Vis features"); assert_bits_equal(lg, lo, "Vis labels")
    rays["tMin"] = 1e-2
    fg, lg, vg = rs[0].gen_precom_data(0, rays)
    fo, lo, vo = world.gen_precom_data(0, rays)
    assert_bits_equal(fg, fo, "Precom features"); assert_bits_equal(lg, lo, "Precom labels"); assert_bits_equal(vg, vo, "Precom valid")


def test_garden_samples_in_flight_share_textures(gpu_required, oracle):
    """dprt_adopt_scene hands the texels, the material -> texture table and the environment map to the other contexts."""
    rs, world, _ = build_garden_pair(oracle, 1, 128, 72, spp=4, bounces=2)
    S = dprt.SamplesInFlight(rs[0], k=2)
    try:
        img_g = S.launch()
    finally:
        S.close()
    img_o = world.launch()
    assert np.allclose(img_g, img_o, rtol=1e-6, atol=1e-7)


def test_garden_full_size_one_million_instanced_triangles_1080p(gpu_required, oracle):
    """Benchmark-size case of the front end: ~1 M instanced triangles (30 K leaf instances, three levels), 1920x1080, textures,
    cut-outs and environment map on -- the image and every path record bit-exact against the oracle."""
    w, h = 1920, 1080
    rs, world, g = build_garden_pair(oracle, 1, w, h, spp=1, bounces=2, clusters=9000, ground=(300, 300), cluster_scale=0.004)
    assert g["objects"][0].ntris > 900000
    R = rs[0]
    img_g, img_o = R.launch(), world.launch()
    assert_bits_equal(img_g, img_o, "garden image, 1 M instanced triangles at 1080p")
    n = R.path_size
    assert n == world.path_size(0) and n > 100000
    tot = n * (1 + R.cfg.shadowPathCount)
    assert_records_equal(R.download(D.BUF_PATHS, tot), world.download(0, D.BUF_PATHS, tot), "paths + shadow paths after the last bounce")
