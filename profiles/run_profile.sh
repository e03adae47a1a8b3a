#!/bin/bash
# One gpurun call: plain runs first (must exit 0), then the ncu launch list of the default bench command and the
# --set full captures of the dominant kernels. usage (from the repo root): gpurun -- bash profiles/run_profile.sh <tag>
set -u
TAG=${1:-rX}
O=gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-extras"
WORK="python profiles/prof_workload.py"
$BENCH > $O/${TAG}_bench_plain.json 2> $O/${TAG}_bench_plain.err || { echo "plain bench failed"; tail -5 $O/${TAG}_bench_plain.err; exit 1; }
$WORK > $O/${TAG}_work_plain.log 2>&1 || { echo "plain workload failed"; tail -5 $O/${TAG}_work_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv $BENCH > $O/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:trace_kernel -c 200 --csv --log-file $O/${TAG}_trace_dram.csv $BENCH > $O/${TAG}_ncu_dram.log 2>&1
echo "trace dram rc=$?"
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 0 -c 3 -o $O/${TAG}_trace -f $BENCH > $O/${TAG}_ncu_trace.log 2>&1
echo "trace rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"partition_kernel|mlp_kernel" -s 4 -c 8 -o $O/${TAG}_part_mlp -f $WORK > $O/${TAG}_ncu_part.log 2>&1
echo "partition/mlp rc=$?"
ls -la $O | tail -12
