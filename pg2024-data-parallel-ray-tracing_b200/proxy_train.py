"""Training side of the neural visibility / depth proxies.

Mirrors the reference's ``trainingcode/`` pipeline around the operator ``dprt_gen_train_data`` (the Vis pipeline,
optix/vis_ray_kernel.cu:98-161):

* :func:`sample_training_rays` -- query rays of one proxy AABB. The reference fills ``rayBuffer`` in host code that
  is not in its tree; what the renderer later asks a proxy is always "from a point on the AABB surface, looking
  inward" (shadow_ray_kernel.cu:198-350: the entry point for rays that start outside, the *exit* point with the
  reversed direction for rays that start inside), so that is what is sampled here: area-uniform surface points,
  directions uniform over the inward hemisphere, tMin = 1e-5, tMax = FLT_MAX (vis_ray_kernel.cu:121-134).
* :func:`vis_dataset` / :func:`depth_dataset` -- trainingcode/datasets.py:149-227: label 1.0 means "missed";
  the vis set keeps every hit and 1.5 x as many misses (``radio = 1.5``) with target 1 = occluded, 0 = free; the depth
  set keeps the hits with target t / maxLength.
* :func:`train_proxy` -- trainingcode/main.py:14-171: seed 19990201, 80/20 split, batches of 12 800, Adam (5e-4 for
  the 256-wide trunk), MSE (vis) / L1 (depth), ReduceLROnPlateau(factor 0.1, patience 10), reshuffle every epoch.
* the trained module goes to the renderer through :func:`proxy.pack_module` (utils/exportHalfModule.py equivalent).

PyTorch lives on this side only.
"""
import numpy as np

from . import ctypes_defs as D
from . import proxy

SEED = 19990201          # trainingcode/main.py:76


def sample_training_rays(aabb_min, aabb_max, n, seed=0):
    """n rays from area-uniform points on the AABB surface into the box (uniform inward hemisphere)."""
    rng = np.random.default_rng(seed)
    mn, mx = np.asarray(aabb_min, np.float64), np.asarray(aabb_max, np.float64)
    ext = mx - mn
    area = np.array([ext[1] * ext[2], ext[1] * ext[2], ext[0] * ext[2], ext[0] * ext[2], ext[0] * ext[1], ext[0] * ext[1]])
    face = rng.choice(6, size=n, p=area / area.sum())
    axis, side = face // 2, face % 2                       # side 0 = min face (inward normal +axis), 1 = max face
    o = mn + rng.random((n, 3)) * ext
    o[np.arange(n), axis] = np.where(side == 0, mn[axis], mx[axis])
    # uniform hemisphere around the inward normal
    z = rng.random(n)
    ph = 2.0 * np.pi * rng.random(n)
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    loc = np.stack([r * np.cos(ph), r * np.sin(ph), z], 1)
    d = np.zeros((n, 3))
    a1, a2 = (axis + 1) % 3, (axis + 2) % 3
    d[np.arange(n), a1] = loc[:, 0]
    d[np.arange(n), a2] = loc[:, 1]
    d[np.arange(n), axis] = np.where(side == 0, loc[:, 2], -loc[:, 2])
    rays = np.zeros(n, D.RAY_DTYPE)
    rays["origin"] = o.astype(np.float32)
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["tMin"] = 1e-5
    rays["tMax"] = np.finfo(np.float32).max
    return rays


def vis_dataset(features, labels, ratio=1.5, seed=SEED):
    """loadNormalizedDatasetsBalanceVIS (datasets.py:149-193): all hits + ratio x hits misses; target 1 = hit."""
    rng = np.random.default_rng(seed)
    hit = np.nonzero(labels != 1.0)[0]
    miss = np.nonzero(labels == 1.0)[0]
    miss = rng.permutation(miss)[: int(hit.size * ratio)]
    idx = np.concatenate([miss, hit])
    return features[idx], (labels[idx] != 1.0).astype(np.float32)


def depth_dataset(features, labels):
    """loadNormalizedDatasetsDepth (datasets.py:195-227): hits only, target = t / maxLength."""
    hit = labels != 1.0
    return features[hit], labels[hit].astype(np.float32)


def train_proxy(data, target, kind="vis", width=256, nres=4, epochs=20, lr=None, batch=12800, seed=SEED, device=None, log=None):
    """trainingcode/main.py:14-171. Returns (module on CPU in eval mode, [test loss per epoch])."""
    import torch
    torch.manual_seed(seed)
    np.random.seed(seed)
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    lr = lr if lr is not None else (1e-4 if width >= 512 else 5e-4)
    X = torch.as_tensor(np.ascontiguousarray(data, np.float32))
    y = torch.as_tensor(np.ascontiguousarray(target, np.float32))
    perm = torch.as_tensor(np.random.permutation(X.shape[0]))
    X, y = X[perm], y[perm]
    ntrain = int(X.shape[0] * 0.8)                                   # getDatasets, datasets.py:270-285
    Xtr, ytr, Xte, yte = X[:ntrain].to(device), y[:ntrain].to(device), X[ntrain:].to(device), y[ntrain:].to(device)
    model = proxy.make_proxy(width, nres).to(device)
    loss_fn = torch.nn.MSELoss() if kind == "vis" else torch.nn.L1Loss()
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=0.1, patience=10)
    hist = []
    for ep in range(epochs):
        model.train()
        for i in range(0, ntrain, batch):
            pred = torch.squeeze(model(Xtr[i:i + batch]), 1)
            loss = loss_fn(pred, ytr[i:i + batch])
            opt.zero_grad()
            loss.backward()
            opt.step()
        model.eval()
        with torch.no_grad():
            tot, nb = 0.0, 0
            for i in range(0, Xte.shape[0], batch):
                tot += loss_fn(torch.squeeze(model(Xte[i:i + batch]), 1), yte[i:i + batch]).item()
                nb += 1
        hist.append(tot / max(1, nb))
        sched.step(hist[-1])
        if log:
            log(f"epoch {ep + 1}: test loss {hist[-1]:.6f} lr {opt.param_groups[0]['lr']:.2e}")
        p2 = torch.randperm(ntrain, device=device)                    # shuffleDatasets, datasets.py:287-292
        Xtr, ytr = Xtr[p2], ytr[p2]
    return model.cpu().eval(), hist


def train_chunk_proxies(gen, aabb_min, aabb_max, n_rays=400000, epochs=20, width=256, nres=4, seed=0, device=None, log=None):
    """gen(rays) -> (features, labels): ``Renderer.gen_train_data`` bound to a local object (or the oracle's).
    Returns (vis_blob, depth_blob, info)."""
    rays = sample_training_rays(aabb_min, aabb_max, n_rays, seed)
    feat, lab = gen(rays)
    xv, yv = vis_dataset(feat, lab)
    xd, yd = depth_dataset(feat, lab)
    vis, hv = train_proxy(xv, yv, "vis", width, nres, epochs, device=device, log=log)
    dep, hd = train_proxy(xd, yd, "depth", width, nres, epochs, device=device, log=log)
    info = {"rays": int(n_rays), "hit_fraction": float((lab != 1.0).mean()), "vis_samples": int(yv.size), "depth_samples": int(yd.size),
            "vis_test_loss": hv, "depth_test_loss": hd}
    return proxy.pack_module(vis), proxy.pack_module(dep), info
