"""Neural visibility/depth proxy: PyTorch definition (training side) and the weight blob libdprt consumes.

Interface of the reference's ``trainingcode/module.py`` proxy networks: ``forward(x[N,5]) -> [N,1]`` with
``x[:,0:3]`` the AABB-normalised entry point and ``x[:,3:5]`` = (phi/2pi, theta/pi).
:class:`ResidualProxy` is parameterised by trunk width and number of residual blocks and uses the *same
parameter names* as the reference classes, so a reference ``state_dict`` loads unchanged:

* ``ResidualProxy(256, 4)`` == ``NeuralVisNetworkWith4Res256SingleOutput`` (module.py:755-794, the class the
  export scripts instantiate: utils/exportHalfModule.py:16-23)
* ``ResidualProxy(256, 6)`` == ``NeuralVisNetworkWith6Res256SingleOutput`` (module.py:796-837)

PyTorch is used on this (training / export) side only; inference inside the renderer is mlp.cu.

* ``ResidualProxy(128, 4)`` == ``NeuralVisNetworkWith4Res128SingleOutput`` (module.py:839-878); ``sigmoid=True`` gives the
  ``...SingleOutputSigmoid`` heads (module.py:880-958). The 512-wide trunk (module.py:701-753) can be trained and packed,
  but libdprt only runs 256- and 128-wide trunks.

Blob layout ("proxy weight blob", little-endian): u32 magic 'LMRP' (0x50524D4C), u32 width, u32 nres, u32 flags (bit 0: Sigmoid head),
then fp32 row-major tensors in this order: enc3.L0 W[32,3] b[32], enc3.L1 W[width/2,32] b[width/2],
enc2.L0 W[32,2] b[32], enc2.L1 W[width/2,32] b[width/2], nres x (W[width,width] b[width]),
post.L0 W[64,width] b[64], post.L1 W[1,64] b[1].
"""
import struct

import numpy as np

MAGIC = 0x50524D4C


def _torch():
    import torch
    return torch


def make_proxy(width=256, nres=4, sigmoid=False):
    """Build the torch module (deferred import so that the renderer side never needs torch). sigmoid: the
    ``...SingleOutputSigmoid`` heads (module.py:880-958) end in nn.Sigmoid instead of nn.LeakyReLU."""
    torch = _torch()
    nn = torch.nn
    F = torch.nn.functional

    class _Res(nn.Module):
        def __init__(self, w):
            super().__init__()
            self.block = nn.Sequential(nn.Linear(w, w))

        def forward(self, x):
            return F.leaky_relu(x + self.block(x))

    class ResidualProxy(nn.Module):
        def __init__(self, w, n, sig):
            super().__init__()
            self.width, self.nres, self.sigmoid = w, n, bool(sig)
            self.encoding3to64 = nn.Sequential(nn.Linear(3, 32), nn.LeakyReLU(), nn.Linear(32, w // 2), nn.LeakyReLU())
            self.encoding2to64 = nn.Sequential(nn.Linear(2, 32), nn.LeakyReLU(), nn.Linear(32, w // 2), nn.LeakyReLU())
            self.res_block = nn.Sequential(*[_Res(w) for _ in range(n)])
            self.post_block = nn.Sequential(nn.Linear(w, 64), nn.LeakyReLU(), nn.Linear(64, 1), nn.Sigmoid() if sig else nn.LeakyReLU())

        def forward(self, x):
            out1 = torch.cat([self.encoding3to64(x[:, 0:3]), self.encoding2to64(x[:, 3:5])], dim=1)
            return self.post_block(out1 + self.res_block(out1))

    return ResidualProxy(width, nres, sigmoid)


def pack_state_dict(sd, width=256, nres=4, sigmoid=False):
    """state_dict (reference key names) -> blob bytes."""
    def t(name):
        v = sd[name]
        v = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
        return np.ascontiguousarray(v, np.float32).reshape(-1)

    parts = [t("encoding3to64.0.weight"), t("encoding3to64.0.bias"), t("encoding3to64.2.weight"), t("encoding3to64.2.bias"),
             t("encoding2to64.0.weight"), t("encoding2to64.0.bias"), t("encoding2to64.2.weight"), t("encoding2to64.2.bias")]
    for i in range(nres):
        parts += [t(f"res_block.{i}.block.0.weight"), t(f"res_block.{i}.block.0.bias")]
    parts += [t("post_block.0.weight"), t("post_block.0.bias"), t("post_block.2.weight"), t("post_block.2.bias")]
    body = np.concatenate(parts)
    half = width // 2
    expect = (32 * 3 + 32 + half * 32 + half) + (32 * 2 + 32 + half * 32 + half) + nres * (width * width + width) + (64 * width + 64) + 65
    assert body.size == expect, (body.size, expect)
    return struct.pack("<IIII", MAGIC, width, nres, 1 if sigmoid else 0) + body.tobytes()


def pack_module(module):
    return pack_state_dict(module.state_dict(), module.width, module.nres, getattr(module, "sigmoid", False))


def unpack_blob(blob):
    """blob -> dict of numpy arrays (for numpy-side reference math)."""
    magic, width, nres, flags = struct.unpack_from("<IIII", blob, 0)
    assert magic == MAGIC
    a = np.frombuffer(blob, np.float32, offset=16)
    half = width // 2
    o = [0]

    def take(*shape):
        n = int(np.prod(shape))
        v = a[o[0]:o[0] + n].reshape(shape)
        o[0] += n
        return v

    d = {"width": width, "nres": nres, "sigmoid": bool(flags & 1)}
    d["e3w0"], d["e3b0"], d["e3w1"], d["e3b1"] = take(32, 3), take(32), take(half, 32), take(half)
    d["e2w0"], d["e2b0"], d["e2w1"], d["e2b1"] = take(32, 2), take(32), take(half, 32), take(half)
    d["rw"], d["rb"] = [], []
    for _ in range(nres):
        d["rw"].append(take(width, width))
        d["rb"].append(take(width))
    d["pw0"], d["pb0"], d["pw1"], d["pb1"] = take(64, width), take(64), take(1, 64), take(1)
    assert o[0] == a.size
    return d


def spread_output_(module, gain=3.0, seed=0):
    """SURVEY.md section 7 hard part 3: random-init proxies give near-constant outputs (all below the 0.5
    decision threshold). For decision tests re-scale the last layers in place so outputs straddle 0.5."""
    torch = _torch()
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        x = torch.rand(4096, 5, generator=g)
        last = module.post_block[2]
        z = module.post_block[1](module.post_block[0](
            torch.cat([module.encoding3to64(x[:, 0:3]), module.encoding2to64(x[:, 3:5])], 1)
            + module.res_block(torch.cat([module.encoding3to64(x[:, 0:3]), module.encoding2to64(x[:, 3:5])], 1))))
        y = last(z)
        mu, sd = y.mean().item(), y.std().item() + 1e-8
        last.weight.mul_(gain * 0.25 / sd)
        last.bias.copy_((last.bias - mu) * (gain * 0.25 / sd) + 0.5)
    return module
