#!/usr/bin/env python
"""Sweep the persistent trace kernel's knobs (env overrides) with bench.py; prints one line per setting.
usage: python profiles/tune_trace.py [refill,trivote,blocksPerSM ...]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
grid = [(8, 1, 6), (8, 8, 6), (8, 12, 6), (8, 16, 6), (8, 20, 6), (4, 12, 6), (12, 12, 6), (16, 16, 6)]
if len(sys.argv) > 1:
    grid = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for r, v, b in grid:
    env = dict(os.environ, DPRT_TRACE_REFILL=str(r), DPRT_TRACE_TRIVOTE=str(v), DPRT_TRACE_BLOCKS_PER_SM=str(b))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "4", "--skip-cpu"], env=env,
                         capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        st = d["stages"]
        print(f"refill={r:2d} trivote={v:2d} blocks/SM={b} value={d['value']:8.1f} Mrays/s  ms/step={d['ms_per_step']:.3f}  " +
              "  ".join(f"{k}={st[k]['ms'] / st[k]['launches']:.3f}ms" for k in ("traverse", "shade", "shadow_trace")) +
              f"  primary={d['primary_closest_hit']['value']:.0f} Mrays/s", flush=True)
    except Exception as e:
        print(r, v, b, "failed", e, out.stderr[-300:], flush=True)
