"""Synthetic scene / light / camera generators for the BASELINE.json configs.

The reference's scene loader is not in the tree; the benchmark scenes are jittered
height-field sheets, one per spatial cell of the unit cube, each cell owned by one rank (= one scene chunk,
``AccelerationStructure.nodeID`` of renderer.cpp:1834-1839). Vertex jitter and material assignment come from
the reference RNG (``tea<4>`` / ``rnd`` of optix/random.hpp:31-67, vectorised here in uint32 numpy).
Everything produced here is plain input data for both libdprt and the oracle.
"""
import numpy as np

from . import ctypes_defs as D


def tea4_np(val0, val1):
    v0 = np.asarray(val0, np.uint32).copy()
    v1 = np.broadcast_to(np.asarray(val1, np.uint32), v0.shape).copy()
    s0 = np.uint32(0)
    with np.errstate(over="ignore"):
        for _ in range(4):
            s0 = np.uint32((int(s0) + 0x9E3779B9) & 0xFFFFFFFF)
            v0 += ((v1 << np.uint32(4)) + np.uint32(0xA341316C)) ^ (v1 + s0) ^ ((v1 >> np.uint32(5)) + np.uint32(0xC8013EA4))
            v1 += ((v0 << np.uint32(4)) + np.uint32(0xAD90777D)) ^ (v0 + s0) ^ ((v0 >> np.uint32(5)) + np.uint32(0x7E95761E))
    return v0


def rnd_np(seed):
    """One LCG step: returns (float32 in [0,1), new seed)."""
    with np.errstate(over="ignore"):
        seed = seed * np.uint32(1664525) + np.uint32(1013904223)
    return ((seed & np.uint32(0x00FFFFFF)).astype(np.float32) / np.float32(0x01000000)), seed


def cell_layout(W):
    """Spatial cells of the unit cube, one per rank: (min, max) per cell."""
    dims = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(W)
    if dims is None:
        dims = (W, 1, 1)
    cells = []
    for k in range(W):
        ix, iy, iz = k % dims[0], (k // dims[0]) % dims[1], k // (dims[0] * dims[1])
        mn = np.array([ix / dims[0], iy / dims[1], iz / dims[2]], np.float64)
        mx = np.array([(ix + 1) / dims[0], (iy + 1) / dims[1], (iz + 1) / dims[2]], np.float64)
        cells.append((mn, mx, iz, dims[2]))
    return cells


def terrain_height(x, y, ts):
    """The height function of make_heightfield_chunk (before vertex jitter), as a fraction of the cell's z range."""
    return np.clip(0.45 + 0.22 * np.sin(7.0 * x + 0.6 * ts) * np.cos(6.0 * y - 0.3 * ts) + 0.08 * np.sin(23.0 * x * y + ts), 0.02, 0.98)


def slabs_from_cuts(cuts):
    return [(np.array([cuts[k], 0.0, 0.0]), np.array([cuts[k + 1], 1.0, 1.0]), 0, 1) for k in range(len(cuts) - 1)]


def balanced_slab_layout(W, cam, terrain_seed=0):
    """Load-balanced k-d partition of the unit cube along x: W slabs whose boundaries are the k/W quantiles of where the
    camera's primary rays land on the landscape (ray-marched against the analytic height function; the usual way a
    data-parallel renderer places its cuts: by expected load, not by volume). Shadow rays start where primary rays land,
    so this balances most of the work; every slab still gets the same number of triangles."""
    nx, ny = 384, 216
    a = (np.arange(nx) + 0.5) / nx * 2.0 - 1.0
    b = 1.0 - (np.arange(ny) + 0.5) / ny * 2.0
    A, B = np.meshgrid(a, b, indexing="xy")
    U, V, Wv, O = (np.array(list(v), np.float64) for v in (cam.U, cam.V, cam.W, cam.origin))
    d = A[..., None] * U + B[..., None] * V + Wv
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    hx = np.full(A.shape, np.nan)
    alive = np.ones(A.shape, bool)
    for t in np.arange(0.0, 4.0, 0.004):
        p = O + t * d
        x, y, z = p[..., 0], p[..., 1], p[..., 2]
        inside = (x >= 0) & (x <= 1) & (y >= 0) & (y <= 1)
        hit = alive & inside & (z <= terrain_height(x, y, terrain_seed))
        hx[hit] = x[hit]
        alive &= ~hit
    xs = np.sort(hx[~np.isnan(hx)])
    cuts = [0.0] + [float(xs[int(len(xs) * k / W)]) for k in range(1, W)] + [1.0]
    return [(np.array([cuts[k], 0.0, 0.0]), np.array([cuts[k + 1], 1.0, 1.0]), 0, 1) for k in range(W)]


def make_heightfield_chunk(cell_min, cell_max, nx, ny, seed, hole_frac=0.0, n_materials=16, water_frac=0.0, terrain_seed=None):
    """nx*ny quads -> 2*nx*ny triangles (minus holes). Returns verts9, normals9, mat_ids (float32/int32).
    terrain_seed: phase of the height function (default: the chunk's own seed, i.e. every chunk its own landscape; the same
    value for all chunks gives ONE landscape cut into chunks, continuous across the cuts up to the vertex jitter)."""
    ts = seed if terrain_seed is None else terrain_seed
    gx, gy = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="ij")
    x = cell_min[0] + (cell_max[0] - cell_min[0]) * gx / nx
    y = cell_min[1] + (cell_max[1] - cell_min[1]) * gy / ny
    vid = (gx * (ny + 1) + gy).astype(np.uint32)
    s = tea4_np(vid, np.uint32(0xC0FFEE + seed))
    j, s = rnd_np(s)
    zr = cell_max[2] - cell_min[2]
    h = 0.45 + 0.22 * np.sin(7.0 * x + 0.6 * ts) * np.cos(6.0 * y - 0.3 * ts) + 0.08 * np.sin(23.0 * x * y + ts)   # == terrain_height before the clip
    amp = 0.25 * min((cell_max[0] - cell_min[0]) / nx, (cell_max[1] - cell_min[1]) / ny) / max(zr, 1e-9)
    z = cell_min[2] + zr * np.clip(h + amp * (j.astype(np.float64) - 0.5), 0.02, 0.98)
    P = np.stack([x, y, z], -1)                                   # [nx+1, ny+1, 3]
    # per-vertex normals from central differences
    dxv = np.gradient(P, axis=0)
    dyv = np.gradient(P, axis=1)
    nrm = np.cross(dxv, dyv)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    p00, p10, p01, p11 = P[:-1, :-1], P[1:, :-1], P[:-1, 1:], P[1:, 1:]
    n00, n10, n01, n11 = nrm[:-1, :-1], nrm[1:, :-1], nrm[:-1, 1:], nrm[1:, 1:]
    t0 = np.concatenate([p00, p10, p11], -1).reshape(-1, 9)
    t1 = np.concatenate([p00, p11, p01], -1).reshape(-1, 9)
    m0 = np.concatenate([n00, n10, n11], -1).reshape(-1, 9)
    m1 = np.concatenate([n00, n11, n01], -1).reshape(-1, 9)
    verts = np.stack([t0, t1], 1).reshape(-1, 9)
    norms = np.stack([m0, m1], 1).reshape(-1, 9)
    qid = np.arange(nx * ny, dtype=np.uint32)
    ms = tea4_np(qid, np.uint32(7 + seed))
    mat = (ms % np.uint32(max(1, n_materials - 1))).astype(np.int32)
    if water_frac > 0.0:
        wsel, _ = rnd_np(tea4_np(qid, np.uint32(0x5EA + seed)))
        mat = np.where(wsel < water_frac, n_materials - 1, mat).astype(np.int32)
    mats = np.repeat(mat, 2)
    if hole_frac > 0.0:
        qx, qy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
        blob = 0.5 + 0.5 * np.sin(9.0 * qx / nx * np.pi + seed) * np.sin(7.0 * qy / ny * np.pi + 2 * seed)
        keep = np.repeat((blob.reshape(-1) >= hole_frac), 2)
        verts, norms, mats = verts[keep], norms[keep], mats[keep]
    return verts.astype(np.float32), norms.astype(np.float32), mats.astype(np.int32)


def make_materials(n=16, water_last=False):
    m = np.zeros(n, D.MATERIAL_DTYPE)
    s = tea4_np(np.arange(n, dtype=np.uint32), np.uint32(0xA1BED0))
    for c in range(3):
        v, s = rnd_np(s)
        m["baseColor"][:, c] = 0.25 + 0.7 * v
    if water_last:
        m["bsdfType"][n - 1] = 1
        m["baseColor"][n - 1] = 1.0
    return m


def make_lights(scale=1.0):
    """Two area-light triangles above the scene, facing down; Le is the reference default (renderer.cpp:1784)."""
    Le = np.array([891.443777, 505.928150, 154.625939], np.float32) * np.float32(scale)
    q = np.array([[0.25, 0.25, 2.0], [0.75, 0.25, 2.0], [0.75, 0.75, 2.0], [0.25, 0.75, 2.0]], np.float32)
    L = np.zeros(2, D.LIGHT_DTYPE)
    L["p0"][0], L["p1"][0], L["p2"][0] = q[0], q[2], q[1]
    L["p0"][1], L["p1"][1], L["p2"][1] = q[0], q[3], q[2]
    L["Le"][:] = Le
    return L


def reference_lights(scene="default"):
    """The per-scene area-light tables of the reference's launch() (renderer.cpp:1725-1796: one #if branch per scene), as
    dprt_light_tri records: "default" (the #else branch: Moana island), "san_miguel", "air_drome", "bistro". The coordinates
    are scene data in the reference's world units; the triangle vertex order is the reference's (it fixes the light normal)."""
    base = np.array([891.443777, 505.928150, 154.625939], np.float32)
    grey = np.float32(505.928150)
    tables = {
        "default": ([(101346.539, 202660.438, 189948.188), (106779.617, 187339.562, 201599.453), (83220.3828, 202660.438, 198400.547),
                     (101346.539, 202660.438, 189948.188), (88653.4609, 187339.562, 210051.812), (83220.3828, 202660.438, 198400.547)],
                    [((0, 1, 2), base), ((3, 4, 5), base)]),
        "san_miguel": ([(14779.412109375, 153398.8125, -4307.61865234375), (-2001.59326171875, 146156.09375, -12428.0302734375),
                        (4770.7685546875, 157817.109375, 12434.7119140625), (-2001.59326171875, 146156.09375, -12428.0302734375)],
                       [((0, 2, 3), base), ((3, 1, 0), base)]),
        "air_drome": ([(-2114.32861328125, 81.61964416503906, 140.63539123535156), (-2114.179443359375, 81.80250549316406, -59.34486389160156),
                       (-1942.9859619140625, 442.6954345703125, 140.6354217529297), (-1942.8370361328125, 442.87811279296875, -59.34484100341797)],
                      [((0, 3, 2), base), ((0, 1, 3), base)]),
        "bistro": ([(0.12184541672468185, 8.149029731750488, 0.31903180480003357), (0.13482901453971863, 8.29036808013916, 0.45987018942832947),
                    (0.2922881841659546, 8.166015625, 0.3377407491207123), (0.3052719831466675, 8.307354927062988, 0.4785784184932709),
                    (-4.677332878112793, 9.999991416931152, -6.165306091308594), (3.3226675987243652, 9.999991416931152, -6.165306091308594),
                    (-4.677332878112793, 9.999991416931152, 13.834686279296875), (3.3226675987243652, 9.999991416931152, 13.834686279296875),
                    (6.292665004730225, 4.862171173095703, 1.3878426551818848), (6.260426998138428, 4.918705940246582, 1.4343327283859253),
                    (6.340075969696045, 4.868965148925781, 1.4374041557312012), (6.307837963104248, 4.925500869750977, 1.4838939905166626)],
                   [((0, 2, 3), 10.0 * grey), ((3, 1, 0), 10.0 * grey), ((7, 6, 4), 0.001 * grey), ((4, 5, 7), 0.001 * grey),
                    ((8, 10, 11), 30.0 * grey), ((11, 9, 8), 30.0 * grey)]),
    }
    pts, tris = tables[scene]
    P = np.asarray(pts, np.float32)
    L = np.zeros(len(tris), D.LIGHT_DTYPE)
    for k, ((a, b, c), le) in enumerate(tris):
        L["p0"][k], L["p1"][k], L["p2"][k] = P[a], P[b], P[c]
        L["Le"][k] = np.broadcast_to(np.float32(le), 3)
    return L


# x-cuts of the benchmark landscape (seed 0, default_camera) that equalise the rays each chunk owner walks over ALL bounces
# of the per-sample loop, measured with the oracle on a coarse mesh: profiles/calibrate_slabs.py (which prints this table).
# Empty entry: fall back to the primary-ray quantiles of balanced_slab_layout.
CALIBRATED_SLAB_CUTS = {2: [0.0, 0.49372, 1.0], 4: [0.0, 0.31297, 0.49534, 0.68698, 1.0],
                        8: [0.0, 0.21956, 0.31019, 0.38708, 0.49715, 0.59792, 0.68832, 0.78837, 1.0]}


def _is_default_camera(cam):
    ref = default_camera(cam.width, cam.height)
    return all(abs(a - b) < 1e-6 for v in ("origin", "U", "V", "W") for a, b in zip(getattr(cam, v), getattr(ref, v)))


def default_camera(width, height):
    # SURVEY.md 8(d) proposed (0.5,-1.5,0.8)->(0.5,0.5,0.3) at 40 deg; from there only 12 % of the 16:9 frame
    # hits the unit-square sheet. This oblique, narrower view keeps ~90 % of the primary rays on geometry.
    return D.make_camera((0.5, -0.4, 1.0), (0.5, 0.5, 0.45), (0.0, 0.0, 1.0), 28.0, width, height)


class Chunk:
    def __init__(self, index, node_id, verts, normals, mats, aabb=None):
        """aabb: (min, max) known without the vertices (make_scene(only=...): every rank must describe a chunk by the same
        box whether or not it holds the geometry); default: the vertices' own bounds."""
        self.index, self.node_id = index, node_id
        self.verts, self.normals, self.mats = verts, normals, mats
        if aabb is None:
            mn = verts.reshape(-1, 3).min(0).astype(np.float32)
            mx = verts.reshape(-1, 3).max(0).astype(np.float32)
        else:
            mn, mx = np.asarray(aabb[0], np.float32), np.asarray(aabb[1], np.float32)
        eps = np.float32(1e-4)
        self.aabb_min, self.aabb_max = mn - eps, mx + eps

    def desc(self, is_proxy):
        return D.make_object_desc(self.node_id, self.aabb_min, self.aabb_max, is_proxy=int(is_proxy))

    @property
    def ntris(self):
        return self.verts.shape[0]


def make_scene(W, tris_per_chunk, water_frac=0.0, seed=0, layout="cells", camera=None, continuous=None, cuts=None, only=None):
    """W chunks (one per rank); ~tris_per_chunk triangles each. Returns (chunks, materials, lights).
    layout "cells": 2x1x1 / 2x2x1 / 2x2x2 spatial cells (upper cells perforated), every cell its own landscape;
    "slabs": ONE continuous landscape over the unit square cut into W x-slabs -- at `cuts` (W + 1 increasing x values) when
    given, else at the calibrated all-bounce cuts for the benchmark camera (CALIBRATED_SLAB_CUTS), else where the camera's
    primary rays put equal load (balanced_slab_layout).
    only: iterable of chunk indices whose geometry is wanted (one process per GPU with 12.5 M-triangle chunks: a rank builds
    its own chunk only). Every chunk is then described by its analytic box (cell footprint x the height function's clip range)
    on every rank, and the chunks outside `only` come without vertices."""
    slabs = layout == "slabs" and W > 1
    if continuous is None:
        continuous = slabs
    if slabs:
        if cuts is None and seed == 0 and camera is not None and _is_default_camera(camera):
            cuts = CALIBRATED_SLAB_CUTS.get(W)
        cells = slabs_from_cuts(cuts) if cuts is not None else balanced_slab_layout(W, camera, seed)
    else:
        cells = cell_layout(W)
    chunks = []
    for k, (mn, mx, iz, nz) in enumerate(cells):
        hole = 0.35 if (nz > 1 and iz == nz - 1) else 0.0           # upper sheets let rays through to the lower cells
        quads = tris_per_chunk / 2 / (1.0 - 0.45 * (hole > 0))
        ax, ay = mx[0] - mn[0], mx[1] - mn[1]
        nx = max(2, int(round(np.sqrt(quads * ax / ay))))
        ny = max(2, int(round(quads / nx)))
        box = None
        if only is not None:
            zr = mx[2] - mn[2]
            box = (np.array([mn[0], mn[1], mn[2] + 0.02 * zr]), np.array([mx[0], mx[1], mn[2] + 0.98 * zr]))
            if k not in set(only):
                chunks.append(Chunk(k, k, None, None, None, aabb=box))
                continue
        v, n, m = make_heightfield_chunk(mn, mx, nx, ny, seed + k, hole_frac=hole, water_frac=water_frac,
                                         terrain_seed=seed if continuous else None)
        chunks.append(Chunk(k, k, v, n, m, aabb=box))
    return chunks, make_materials(16, water_last=water_frac > 0), make_lights()


def camera_rays(cam, sample=0):
    """The PathGen rays as dprt_ray records (float64 math, for harness use; the parity path is the kernel)."""
    w, h = cam.width, cam.height
    pix = np.arange(w * h, dtype=np.uint32)
    s = tea4_np(pix, np.uint32(sample))
    x1, s = rnd_np(s)
    x2, s = rnd_np(s)
    col, row = (pix % w).astype(np.float64), (pix // w).astype(np.float64)
    a = 2.0 * (col + x1) / w - 1.0
    b = 1.0 - 2.0 * (row + x2) / h
    U, V, Wv = (np.array(list(v), np.float64) for v in (cam.U, cam.V, cam.W))
    d = a[:, None] * U + b[:, None] * V + Wv
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros(w * h, D.RAY_DTYPE)
    rays["origin"] = np.array(list(cam.origin), np.float32)
    rays["direction"] = d.astype(np.float32)
    rays["tMin"] = D.DPRT_EPSILON
    rays["tMax"] = np.finfo(np.float32).max
    return rays


def save_scene(path, chunks, materials, lights, cam, models=None):
    """Scene file of the C++ host csrc/dprt_render.cpp (layout documented there). models: {scene_index: (vis_blob, depth_blob)}."""
    import ctypes as C
    with open(path, "wb") as f:
        f.write(b"DPRTSCN1")
        f.write(np.array([len(chunks), len(materials), len(lights)], np.int32).tobytes())
        f.write(bytes(C.string_at(C.addressof(cam), C.sizeof(cam))))
        f.write(np.ascontiguousarray(materials).tobytes())
        f.write(np.ascontiguousarray(lights).tobytes())
        for c in chunks:
            d = c.desc(False)
            f.write(bytes(C.string_at(C.addressof(d), C.sizeof(d))))
            f.write(np.array([c.ntris], np.int64).tobytes())
            f.write(np.ascontiguousarray(c.verts, np.float32).tobytes())
            f.write(np.ascontiguousarray(c.normals, np.float32).tobytes())
            f.write(np.ascontiguousarray(c.mats, np.int32).tobytes())
            vb, db = (models or {}).get(c.index, (None, None))
            for blob in (vb, db):
                b = bytes(blob) if blob is not None else b""
                f.write(np.array([len(b)], np.int64).tobytes())
                f.write(b)


def load_exr(path):
    """The OpenEXR subset dprt_render writes (single part, uncompressed scanlines, FLOAT channels B/G/R, increasing Y)
    -> [h, w, 3] RGB float32. Enough of a reader to check the writer without an EXR library."""
    import struct
    b = open(path, "rb").read()
    magic, version = struct.unpack_from("<II", b, 0)
    assert magic == 20000630 and (version & 0xff) == 2 and not (version & 0x1e00), "not a single-part scanline OpenEXR 2 file"
    o, attrs = 8, {}
    while b[o] != 0:
        e = b.index(b"\0", o); name = b[o:e].decode(); o = e + 1
        e = b.index(b"\0", o); typ = b[o:e].decode(); o = e + 1
        (size,) = struct.unpack_from("<i", b, o); o += 4
        attrs[name] = (typ, b[o:o + size]); o += size
    o += 1
    assert attrs["compression"][1] == b"\0" and attrs["lineOrder"][1] == b"\0"
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    chans, c = [], attrs["channels"][1]
    p = 0
    while c[p] != 0:
        e = c.index(b"\0", p); nm = c[p:e].decode(); p = e + 1
        (ptype,) = struct.unpack_from("<i", c, p); p += 16
        assert ptype == 2, "FLOAT channels only"
        chans.append(nm)
    assert chans == sorted(chans) and set(chans) == {"R", "G", "B"}
    offs = struct.unpack_from(f"<{h}Q", b, o)
    img = np.zeros((h, w, 3), np.float32)
    for k, off in enumerate(offs):
        y, size = struct.unpack_from("<ii", b, off)
        assert size == w * 4 * len(chans)
        row = np.frombuffer(b, "<f4", w * len(chans), off + 8).reshape(len(chans), w)
        for ci, nm in enumerate(chans):
            img[y - y0, :, "RGB".index(nm)] = row[ci]
    return img


def save_pfm(path, img):
    a = np.ascontiguousarray(img, np.float32)
    with open(path, "wb") as f:
        f.write(f"PF\n{a.shape[1]} {a.shape[0]}\n-1.0\n".encode())
        f.write(a[::-1].tobytes())


def load_pfm(path):
    """RGB float32 PFM as written by dprt_render (little endian, bottom-up) -> [h, w, 3] top-down."""
    with open(path, "rb") as f:
        assert f.readline().strip() == b"PF"
        w, h = (int(x) for x in f.readline().split())
        scale = float(f.readline())
        assert scale < 0
        img = np.frombuffer(f.read(), "<f4").reshape(h, w, 3)
    return img[::-1].copy()
