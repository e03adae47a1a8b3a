#!/bin/bash
# compute-sanitizer over small invocations of every kernel family (memcheck: out-of-bounds / misaligned accesses in global and shared
# memory; racecheck: shared-memory hazards in the warp-queue / partition / MLP staging code). usage: gpurun -- bash profiles/sanitize.sh <tag>
set -u
TAG=${1:-rX}; O=gpurun_out
export DPRT_P2P_TIMEOUT_MS=120000
SEL='single_rank_image_bit_exact or test_multi_rank_group_peer_memory_exchange_bit_exact and 2-1-0 or test_partition_random_keys or test_proxy_image_random_init_bit_exact or test_samples_in_flight_same_image and 2'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --target-processes all python -m pytest tests/test_gpu_parity.py tests/test_gpu_proxy.py -x -q -m gpu -k "$SEL" > $O/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $O/${TAG}_memcheck.log | tail -6
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --target-processes all python -m pytest tests/test_gpu_parity.py tests/test_gpu_mlp.py -x -q -m gpu -k "single_rank_image_bit_exact or test_partition_random_keys or test_mlp_ragged_batch_sizes and 300" > $O/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard" $O/${TAG}_racecheck.log | tail -6
