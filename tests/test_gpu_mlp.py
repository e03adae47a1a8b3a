"""GPU parity of the fused tcgen05 proxy MLP (-m gpu), called through dprt_mlp_infer of the C ABI.

Tolerance (BASELINE.json north_star): |gpu - PyTorch-CPU fp32 reference module| <= 1e-3 abs on the random-init
network of BASELINE config 1, with fp16 operands (the reference's deployed NN_Float) and fp32 accumulation.
bf16 operands are checked at 4e-3 (8-bit mantissa; measured 6.6e-4 .. 1.2e-3 in numpy emulation).
"""
import os

import numpy as np
import pytest

from helpers import D, dprt

pytestmark = pytest.mark.gpu


def _renderer_with_proxy(blob, mlp_dtype):
    cfg = dprt.make_config(16, 16, scene_size=2, proxy_mode=1, mlp_dtype=mlp_dtype)
    R = dprt.Renderer(cfg, rank=0, world=2)
    R.upload_proxy(1, dprt.make_object_desc(1, [0, 0, 0], [1, 1, 1], is_proxy=1), blob, blob)
    return R


def _model(nres, spread=False):
    import torch
    torch.manual_seed(19990201)
    m = dprt.proxy.make_proxy(256, nres).eval()
    if spread:
        dprt.proxy.spread_output_(m, gain=3.0, seed=1)
    return m


# north_star fixes proxy parity at 1e-3 abs: fp16 operands (the reference's NN_Float, and the default mlpDtype) meet it.
# bf16 operands (mlpDtype = 0, opt-in) do NOT: 8 mantissa bits give ~3e-3 on this chain. That row is an informational bound
# on the opt-in mode, not a parity claim -- every parity statement in DESIGN.md refers to fp16.
@pytest.mark.parametrize("nres,mlp_dtype,tol", [(4, 1, 1e-3), (6, 1, 1e-3), pytest.param(4, 0, 4e-3, id="4-bf16-out-of-tolerance-mode-4e-3")])
def test_mlp_matches_reference_module_golden(gpu_required, oracle, golden_dir, nres, mlp_dtype, tol):
    g = np.load(os.path.join(golden_dir, "mlp_golden.npz"))
    m = _model(nres)
    blob = dprt.proxy.pack_module(m)
    R = _renderer_with_proxy(blob, mlp_dtype)
    x16 = g["x_f16"]
    y = R.mlp_infer(1, 0, x16).view(np.float16).astype(np.float32)
    ref = g[f"y_{nres}res256"]
    err = np.abs(y - ref).max()
    print(f"nres={nres} dtype={'bf16' if mlp_dtype == 0 else 'fp16'} max|gpu-ref|={err:.3e}")
    assert err <= tol
    yo, _ = oracle.mlp_forward(blob, x16)
    assert np.abs(y - yo).max() <= tol


@pytest.mark.parametrize("width,nres,sig,spread", [(128, 4, False, False), (128, 4, False, True), (128, 4, True, False), (256, 4, True, False)])
def test_mlp_other_reference_variants(gpu_required, oracle, golden_dir, width, nres, sig, spread):
    """The 128-wide trunk (embedded in the 256-wide kernel with zero weights) and the Sigmoid heads against the reference's
    own module.py outputs (tests/golden/mlp_golden.npz), fp16 operands, 1e-3 abs."""
    import torch
    g = np.load(os.path.join(golden_dir, "mlp_golden.npz"))
    torch.manual_seed(19990201)
    m = dprt.proxy.make_proxy(width, nres, sigmoid=sig).eval()
    if spread:
        dprt.proxy.spread_output_(m, gain=3.0, seed=1)
    blob = dprt.proxy.pack_module(m)
    R = _renderer_with_proxy(blob, 1)
    x16 = g["x_f16"]
    y = R.mlp_infer(1, 0, x16).view(np.float16).astype(np.float32)
    ref = g[f"y_{nres}res{width}" + ("_sigmoid" if sig else "") + ("_spread" if spread else "")]
    err = np.abs(y - ref).max()
    print(f"width={width} nres={nres} sigmoid={sig} spread={spread} max|gpu-ref|={err:.3e}")
    if spread:
        # spread_output_ rescales the last layer so that outputs straddle 0.5; the random-init 128-wide network is nearly
        # constant (output range 8e-4), so that gain is ~5 000 and multiplies the fp16 rounding of the hidden layers with it:
        # this row checks the embedded trunk's wiring on well-separated outputs (range 2.3), not the 1e-3 tolerance
        assert err <= 1e-2 and ((y > 0.5) != (ref > 0.5)).mean() <= 0.02
    else:
        assert err <= 1e-3
    R.close()


def test_mlp_rejects_the_512_wide_trunk(gpu_required):
    import torch
    torch.manual_seed(1)
    blob = dprt.proxy.pack_module(dprt.proxy.make_proxy(512, 4).eval())
    cfg = dprt.make_config(16, 16, scene_size=2, proxy_mode=1)
    R = dprt.Renderer(cfg, rank=0, world=2)
    with pytest.raises(dprt.DprtError) as e:
        R.upload_proxy(1, dprt.make_object_desc(1, [0, 0, 0], [1, 1, 1], is_proxy=1), blob, blob)
    assert "512" in str(e.value)
    R.close()


@pytest.mark.parametrize("n", [1, 127, 128, 129, 300, 20000, 148 * 128 * 3 + 77])
def test_mlp_ragged_batch_sizes(gpu_required, oracle, n):
    m = _model(4)
    blob = dprt.proxy.pack_module(m)
    R = _renderer_with_proxy(blob, 1)
    x16 = np.random.default_rng(n).random((n, 5)).astype(np.float16).view(np.uint16)
    y = R.mlp_infer(1, 1, x16).view(np.float16).astype(np.float32)
    yo, _ = oracle.mlp_forward(blob, x16)
    assert y.shape == (n,)
    assert np.abs(y - yo).max() <= 1e-3
    assert R.mlp_infer(1, 0, x16[:0]).size == 0


def test_mlp_decisions_on_spread_network(gpu_required, oracle, golden_dir):
    """Thresholded outputs (pred > 0.5, frame_buffer_update.cu:53): decisions may only flip inside the error band."""
    g = np.load(os.path.join(golden_dir, "mlp_golden.npz"))
    m = _model(4, spread=True)
    blob = dprt.proxy.pack_module(m)
    R = _renderer_with_proxy(blob, 1)
    x16 = np.random.default_rng(5).random((50000, 5)).astype(np.float16).view(np.uint16)
    y = R.mlp_infer(1, 0, x16).view(np.float16).astype(np.float32)
    yo, _ = oracle.mlp_forward(blob, x16)
    err = np.abs(y - yo)
    assert err.max() < 2e-2          # outputs are ~40x the raw network's scale
    flips = (y > 0.5) != (yo > 0.5)
    assert np.all(np.abs(yo[flips] - 0.5) < 2e-2)
    assert flips.mean() < 0.01
    yg = R.mlp_infer(1, 0, g["x_f16"]).view(np.float16).astype(np.float32)
    assert np.abs(yg - g["y_4res256_spread"]).max() < 2e-2
