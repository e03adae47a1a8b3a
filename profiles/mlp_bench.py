#!/usr/bin/env python
"""Time the fused proxy-MLP kernel alone (2^20 and 2^22 device-resident queries, CUDA events on the launching stream).
usage: python profiles/mlp_bench.py"""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dprt = importlib.import_module("pg2024-data-parallel-ray-tracing_b200")
import torch
torch.manual_seed(19990201)
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
for nres in (4, 6):
    blob = dprt.proxy.pack_module(dprt.proxy.make_proxy(256, nres).eval())
    flop = 2 * (3 * 32 + 2 * 32 + 2 * 32 * 128 + nres * 256 * 256 + 64 * 256 + 64)
    for name, dt in (("bf16", 0), ("fp16", 1)):
        cfg = dprt.make_config(16, 16, scene_size=2, proxy_mode=1, mlp_dtype=dt)
        P = dprt.Renderer(cfg, rank=0, world=2)
        P.upload_proxy(1, dprt.make_object_desc(1, [0, 0, 0], [1, 1, 1], is_proxy=1), blob, blob)
        for n in (1 << 20, 1 << 22):
            x = np.random.default_rng(0).random((n, 5)).astype(np.float16).view(np.uint16)
            dx, dy = P.device_alloc(x.nbytes), P.device_alloc(n * 2)
            P.h2d(dx, x)
            for _ in range(3):
                P.mlp_infer_device(1, 0, dx, n, dy)
            P.synchronize()
            P.timer_start()
            k = 20
            for _ in range(k):
                P.mlp_infer_device(1, 0, dx, n, dy)
            ms = P.timer_stop() / k
            tf = flop * n / (ms * 1e-3) / 1e12
            print(f"nres={nres} {name} n={n}: {ms:.4f} ms  {n / ms / 1e3:.1f} Mq/s  {tf:.1f} TFLOP/s  frac_of_measured_burst={tf / pk['bf16_tflops']:.3f}", flush=True)
            P.device_free(dx); P.device_free(dy)
        P.close()
