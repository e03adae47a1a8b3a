#!/bin/bash
# quick A/B of trace-kernel knobs: bash profiles/quick_bench.sh "VAR=val VAR2=val" ...   (one bench run per argument;
# BENCH_ARGS="--serial 1" inside an argument adds bench.py flags to that run)
for envs in "$@"; do
  env $envs bash -c 'timeout 200 python bench.py --steps 4 --skip-cpu --skip-extras $BENCH_ARGS 2>/dev/null' | tail -1 > /tmp/qb.json
  python - "$envs" <<'PY'
import json, sys
j = json.loads(open("/tmp/qb.json").read())
st = j["stages"]
print(f"{sys.argv[1]:40s} value={j['value']:8.1f} Mrays/s ms/step={j['ms_per_step']:.3f} " + " ".join(f"{k}={st[k]['ms'] / j['steps']:.3f}" for k in ("traverse", "shade", "shadow_trace")), flush=True)
PY
done
