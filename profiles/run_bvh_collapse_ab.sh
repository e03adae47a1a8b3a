#!/bin/bash
# One gpurun call: A/B of the two BVH8 collapses (DPRT_BVH_COLLAPSE=greedy|optimal) on the HBM-resident 12.5 M-triangle chunk and
# on the 1 M-triangle benchmark chunk, after a parity subset with the optimal collapse. Results -> gpurun_out/r3d_*.
mkdir -p gpurun_out
DPRT_BVH_COLLAPSE=optimal timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "image_bit_exact or stagewise or hit_cache" > gpurun_out/r3d_parity_optimal.log 2>&1
echo "parity rc=$?" >> gpurun_out/r3d_parity_optimal.log
for mode in optimal greedy; do
  DPRT_BVH_COLLAPSE=$mode timeout 75 python bench.py --tris 12500000 --steps 4 --skip-cpu --skip-extras --skip-oracle-counts > gpurun_out/r3d_12p5M_$mode.json 2> gpurun_out/r3d_12p5M_$mode.err
done
for mode in optimal greedy; do
  DPRT_BVH_COLLAPSE=$mode timeout 40 python bench.py --steps 8 --skip-cpu --skip-extras --skip-oracle-counts > gpurun_out/r3d_1M_$mode.json 2> gpurun_out/r3d_1M_$mode.err
done
tail -3 gpurun_out/r3d_parity_optimal.log
python - <<'PY'
import json
for f in ("12p5M_optimal", "12p5M_greedy", "1M_optimal", "1M_greedy"):
    try:
        d = json.loads(open(f"gpurun_out/r3d_{f}.json").read().strip().splitlines()[-1])
        st = d["stages"]
        print(f, round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 3), "ms", {k: round(st[k]["ms"] / d["steps"], 3) for k in ("traverse", "shade", "shadow_trace")})
    except Exception as e:
        print(f, "no result", e)
PY
