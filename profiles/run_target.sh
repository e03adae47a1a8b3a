#!/bin/bash
# One gpurun --gpus 8 call: BASELINE configs[3] (8 x 12.5 M triangles, 16 spp, trained proxies) + the default bench at N = 8 with proxies off / on.
set -u
TAG=${1:-rX}; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29551 profiles/run_config4.py --out $O/${TAG}_config4_n8.json > $O/${TAG}_config4_n8.log 2>&1; echo "config4 rc=$?"; grep -v "^$" $O/${TAG}_config4_n8.log | tail -2 | cut -c1-600
timeout 600 $TR --master-port 29552 bench.py --gpus 8 --steps 6 > $O/${TAG}_bench_n8.json 2> $O/${TAG}_bench_n8.err; echo "bench N=8 rc=$?"; tail -2 $O/${TAG}_bench_n8.err | cut -c1-300
timeout 600 $TR --master-port 29553 bench.py --gpus 8 --steps 6 --proxy 1 --skip-oracle-counts > $O/${TAG}_bench_n8_proxy.json 2> $O/${TAG}_bench_n8_proxy.err; echo "bench N=8 proxy rc=$?"; tail -2 $O/${TAG}_bench_n8_proxy.err | cut -c1-300
python - <<PY
import json
for f in ("${TAG}_bench_n8", "${TAG}_bench_n8_proxy"):
    try:
        l = [json.loads(x) for x in open(f"$O/{f}.json") if x.startswith("{")][-1]
        print(f, "value", round(l["value"]), "ms/step", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"]), "parity", (l.get("parity") or {}).get("ok"), "K", l["samples_in_flight"],
              "iters", l["alltoall"]["exchange_iters_per_step"], "lb", round(l["load_balance"]["rays_walked_max_over_mean"], 3), "reduce_ms", round(l["image_reduce"]["ms"], 3))
        print("   ranks", [(r["rank"], round(r["rays_walked_per_step"] / 1e6, 2), round(r["busy_ms_per_step"], 2), round(r["exchange_ms_per_step"], 2)) for r in l["ranks"]])
        print("   stages", {k: (round(v["ms"] / l["steps"], 3), v["launches"] // l["steps"]) for k, v in l["stages"].items()})
    except Exception as e:
        print(f, "ERR", e)
try:
    c = json.load(open("$O/${TAG}_config4_n8.json"))
    print(c["config"], "| scene", round(c["scene_seconds"], 1), "s upload", round(c["upload_seconds_incl_bvh8_build"], 1), "s train", round(c["proxy_training"]["seconds"], 1), "s")
    for r in c["runs"]:
        print("  ", r["label"], "| Mrays/s", round(r["Mrays_per_s"]), "samples/s", round(r["samples_per_s"] / 1e6, 1), "M ms/sample", round(r["ms_per_sample"], 2), "iters", r["alltoall"]["exchange_iters_per_sample"], "mlp", r["mlp"])
        print("     stages", {k: round(v["ms"], 2) for k, v in r["stages_profiled_sample_rank0"].items()})
    print("  image on vs off", c.get("image_proxy_on_vs_off"))
except Exception as e:
    print("config4 ERR", e)
PY
