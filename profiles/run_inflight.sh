#!/bin/bash
# A/B of samples in flight at N GPUs. usage: gpurun --gpus N -- bash profiles/run_inflight.sh <tag> <N> <steps> "<k list>"
set -u
TAG=${1:-rX}; N=${2:-8}; STEPS=${3:-6}; KS=${4:-"1 2 3"}
O=gpurun_out
for K in $KS; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530 + K))"
  timeout 600 $TR bench.py --gpus $N --steps $STEPS --warmup 3 --inflight $K --skip-oracle-counts > $O/${TAG}_n${N}_k${K}.json 2> $O/${TAG}_n${N}_k${K}.err; echo "N=$N K=$K rc=$?"; tail -2 $O/${TAG}_n${N}_k${K}.err | cut -c1-300
done
python - <<PY
import json
for K in "$KS".split():
    f = f"$O/${TAG}_n${N}_k{K}.json"
    try:
        l = [json.loads(x) for x in open(f) if x.startswith("{")][-1]
        print("N=$N K=" + K, "value", round(l["value"]), "ms/step", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"]), "parity", (l.get("parity") or {}).get("ok"), "in flight", l.get("samples_in_flight"))
    except Exception as e:
        print(K, "ERR", e)
PY
