"""The C++ host (csrc/dprt_render.cpp, the Renderer::launch equivalent) against the oracle: the binary reads a scene
file, drives libdprt.so through the C ABI only, writes a PFM; the image must carry the oracle's bits."""
import os
import subprocess

import numpy as np
import pytest

from helpers import assert_bits_equal, dprt

pytestmark = pytest.mark.gpu
BIN = os.path.join(os.path.dirname(dprt.host.LIB_PATH), "dprt_render")


def _scene_and_oracle(oracle, tmp_path, W, tris, w, h, spp, bounces, path_gen_mode):
    chunks, mats, lights = dprt.scene.make_scene(W, tris, water_frac=0.02)
    cam = dprt.scene.default_camera(w, h)
    path = str(tmp_path / f"scene_w{W}.dprt")
    dprt.scene.save_scene(path, chunks, mats, lights, cam)
    cfg = dprt.make_config(w, h, spp=spp, bounces=bounces, scene_size=W, proxy_mode=0, path_gen_mode=path_gen_mode)
    world = oracle.World(cfg, W)
    for c in chunks:
        world.add_object(c.index, c.desc(False), c.verts, c.normals, c.mats)
    world.set_materials(mats); world.set_lights(lights); world.set_camera(cam)
    return path, world.launch()


@pytest.mark.parametrize("W", [1, 3])
def test_dprt_render_binary_matches_oracle(gpu_required, oracle, tmp_path, W):
    assert os.path.exists(BIN), "build() did not produce dprt_render"
    scene, img_o = _scene_and_oracle(oracle, tmp_path, W, 8000, 160, 90, 2, 3, 1 if W > 1 else 0)
    out = str(tmp_path / "img.pfm")
    p = subprocess.run([BIN, "--scene", scene, "--out", out, "--spp", "2", "--bounces", "3", "--world", str(W)],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    assert '"rays_walked"' in p.stdout
    img = dprt.scene.load_pfm(out)
    assert img.shape == (90, 160, 3) and img.max() > 0
    assert_bits_equal(img.reshape(-1), np.asarray(img_o, np.float32).reshape(-1), f"dprt_render image, W={W}")


def test_dprt_render_one_process_per_gpu(gpu_required, oracle, tmp_path):
    """Two dprt_render processes, one per GPU, NCCL unique id through a file: the reference's deployment."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    scene, img_o = _scene_and_oracle(oracle, tmp_path, 2, 8000, 160, 90, 2, 3, 1)
    out, idf = str(tmp_path / "img2.pfm"), str(tmp_path / "nccl.id")
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([BIN, "--scene", scene, "--out", out, "--spp", "2", "--bounces", "3", "--nccl-id-file", idf],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    img = dprt.scene.load_pfm(out)
    err = np.abs(img.reshape(-1) - np.asarray(img_o, np.float32).reshape(-1)).max() / max(1e-30, float(np.abs(img_o).max()))
    assert err <= 1e-6, f"ncclReduce image differs: {err}"      # the reduce sums ranks in an order NCCL chooses


def test_dprt_render_frame_loop_animation_and_exr(gpu_required, oracle, tmp_path):
    """The frame loop of launch() (renderer.cpp:1938-2059): per frame the first two lights move (LIGHT_MOVE), the camera
    origin moves (CAMERA_MOVE), the frame is reset, rendered and saved as <frame><name>.exr -- every frame bit-exact
    against the oracle rendering with the same moved camera and lights."""
    W, tris, w, h, spp, bounces, frames = 1, 8000, 160, 90, 2, 2, 3
    chunks, mats, lights = dprt.scene.make_scene(W, tris)
    cam = dprt.scene.default_camera(w, h)
    scene = str(tmp_path / "anim.dprt")
    dprt.scene.save_scene(scene, chunks, mats, lights, cam)
    out = str(tmp_path / "anim.exr")
    cmove, lmove, lstart = np.float32([0.01, 0.02, -0.005]), np.float32([0.05, 0.0, 0.01]), np.float32([0.2, 0.0, 0.0])
    fmt = lambda v: ",".join(repr(float(x)) for x in v)
    p = subprocess.run([BIN, "--scene", scene, "--out", out, "--spp", str(spp), "--bounces", str(bounces), "--frames", str(frames),
                        "--camera-move", fmt(cmove), "--light-move", fmt(lmove), "--light-start", fmt(lstart), "--inflight", "2"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout.count('"frame"') == frames
    cfg = dprt.make_config(w, h, spp=spp, bounces=bounces, scene_size=W)
    world = oracle.World(cfg, W)
    c = chunks[0]
    world.add_object(c.index, c.desc(False), c.verts, c.normals, c.mats)
    world.set_materials(mats)
    L = lights.copy()
    origin = np.float32(list(cam.origin))
    for f in range(frames):
        for i in range(min(2, L.size)):
            for k in ("p0", "p1", "p2"):
                if f == 0:
                    L[k][i] = L[k][i] + lstart
                L[k][i] = L[k][i] - lmove
        origin = origin + cmove
        cam.origin[:] = origin.tolist()
        world.set_lights(L); world.set_camera(cam)
        img_o = world.launch()
        img = dprt.scene.load_exr(str(tmp_path / f"{f}anim.exr"))
        err = np.abs(img - img_o).max() / max(1e-30, float(np.abs(img_o).max()))
        assert err <= 1e-6, (f, err)                       # two samples in flight: the per-pixel sum over samples in another order
    assert not np.array_equal(dprt.scene.load_exr(str(tmp_path / "0anim.exr")), dprt.scene.load_exr(str(tmp_path / "2anim.exr")))


@pytest.mark.parametrize("W", [1, 2])
def test_dprt_render_real_scene_file_matches_oracle(gpu_required, oracle, tmp_path, W):
    """The real-scene variant of the scene file (DPRTSCN2: flattened instanced geometry with texture coordinates, albedo /
    opacity maps, texture slot per material, environment map) through the C++ host: the oracle's bits."""
    from helpers import build_garden_pair
    w, h, spp, bounces = 160, 90, 2, 3
    _, world, g = build_garden_pair(oracle, W, w, h, spp=spp, bounces=bounces, gpu=False)
    if W > 1:      # dprt_render stripes path generation over the ranks of a group
        cfg = dprt.make_config(w, h, spp=spp, bounces=bounces, scene_size=W, path_gen_mode=1)
        world2 = oracle.World(cfg, W)
        for ob in g["objects"]:
            world2.add_instanced_object(ob.index, ob.desc(False), ob.meshes, ob.instances)
        world2.set_materials(g["materials"]); world2.set_lights(g["lights"]); world2.set_camera(dprt.scene.default_camera(w, h))
        for slot, t in g["textures"].items():
            world2.set_texture(slot, t)
        world2.set_material_textures(g["material_textures"]); world2.set_env_map(g["env_map"], g["env_rotation"])
        world = world2
    img_o = world.launch()
    scene, out = str(tmp_path / "garden.dprt"), str(tmp_path / "garden.pfm")
    dprt.real_scene.save_scene_v2(scene, g, dprt.scene.default_camera(w, h), dprt.flatten_instances)
    p = subprocess.run([BIN, "--scene", scene, "--out", out, "--spp", str(spp), "--bounces", str(bounces), "--world", str(W)],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    img = dprt.scene.load_pfm(out)
    if W == 1:
        assert_bits_equal(img.reshape(-1), np.asarray(img_o, np.float32).reshape(-1), "dprt_render image of the garden")
    else:
        err = np.abs(img - img_o).max() / max(1e-30, float(np.abs(img_o).max()))
        assert err <= 1e-6, err
