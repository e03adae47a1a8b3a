#!/bin/bash
# One gpurun --gpus 8 call: default bench at N = 8 and N = 4 (peer-memory exchange). usage: gpurun --gpus 8 -- bash profiles/run_scale.sh <tag> [steps] [extra bench flags]
set -u
TAG=${1:-rX}; STEPS=${2:-4}; shift 2 || true
O=gpurun_out
for N in 8 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N))"
  timeout 900 $TR bench.py --gpus $N --steps $STEPS --warmup 3 "$@" > $O/${TAG}_bench_n${N}.json 2> $O/${TAG}_bench_n${N}.err; echo "bench N=$N rc=$?"; tail -2 $O/${TAG}_bench_n${N}.err
done
python - <<PY
import json
for N in (8, 4):
    f = f"${TAG}_bench_n{N}"
    try:
        l = [json.loads(x) for x in open(f"$O/{f}.json") if x.startswith("{")][-1]
        print(f, "value", round(l["value"]), "ms/step", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"]), "parity", (l.get("parity") or {}).get("ok"),
              "plane", l.get("exchange_data_plane"), "iters", l["alltoall"]["exchange_iters_per_step"], "wait", round(l["alltoall"]["exchange_wait_ms_per_step_max"], 2),
              "GBps", round(l["alltoall"]["GBps_per_gpu_max"], 1), "reduce_ms", round(l["image_reduce"]["ms"], 3), "lb", round(l["load_balance"]["rays_walked_max_over_mean"], 3))
        print("   ranks", [(r["rank"], round(r["rays_walked_per_step"] / 1e6, 2), round(r["busy_ms_per_step"], 2), round(r["exchange_ms_per_step"], 2)) for r in l["ranks"]])
        print("   stages", {k: (round(v["ms"] / l["steps"], 3), v["launches"] // l["steps"]) for k, v in l["stages"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
