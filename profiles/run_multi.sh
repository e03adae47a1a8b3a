#!/bin/bash
# One gpurun --gpus N call: multi-process parity tests, then bench lines for peer-memory vs NCCL exchange and proxies on/off.
# usage (from the repo root): gpurun --gpus N -- bash profiles/run_multi.sh <tag> <N> [steps]
set -u
TAG=${1:-rX}; N=${2:-2}; STEPS=${3:-4}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cpp_host.py -x -q -m gpu > $O/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 $O/${TAG}_pytest_multi.log
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 3 > $O/${TAG}_bench_n${N}.json 2> $O/${TAG}_bench_n${N}.err; echo "bench p2p rc=$?"; tail -2 $O/${TAG}_bench_n${N}.err
DPRT_P2P=0 timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps $STEPS --warmup 3 --skip-oracle-counts > $O/${TAG}_bench_n${N}_nccl.json 2> $O/${TAG}_bench_n${N}_nccl.err; echo "bench nccl rc=$?"; tail -2 $O/${TAG}_bench_n${N}_nccl.err
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps $STEPS --warmup 3 --proxy 1 --skip-oracle-counts > $O/${TAG}_bench_n${N}_proxy.json 2> $O/${TAG}_bench_n${N}_proxy.err; echo "bench proxy rc=$?"; tail -2 $O/${TAG}_bench_n${N}_proxy.err
python - <<PY
import json
for f in ("${TAG}_bench_n${N}", "${TAG}_bench_n${N}_nccl", "${TAG}_bench_n${N}_proxy"):
    try:
        l = [json.loads(x) for x in open(f"$O/{f}.json") if x.startswith("{")][-1]
        print(f, "value", round(l["value"]), "ms/step", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"]), "parity", (l.get("parity") or {}).get("ok"),
              "plane", l.get("exchange_data_plane"), "iters", l["alltoall"]["exchange_iters_per_step"], "wait", round(l["alltoall"]["exchange_wait_ms_per_step_max"], 2),
              "GBps", round(l["alltoall"]["GBps_per_gpu_max"], 1), "reduce_ms", round(l["image_reduce"]["ms"], 3), "lb", round(l["load_balance"]["rays_walked_max_over_mean"], 3))
    except Exception as e:
        print(f, "ERR", e)
PY
