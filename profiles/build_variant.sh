#!/bin/bash
# A/B builds of the trace kernels: bash profiles/build_variant.sh <name> "<extra nvcc flags for kernels.cu>"
# -> pg2024-data-parallel-ray-tracing_b200/libdprt_<name>.so (same ABI; select it with DPRT_LIB=<path>). Experiments only.
set -e
NAME=$1; EXTRA=${2:-}
cd "$(dirname "$0")/../pg2024-data-parallel-ray-tracing_b200/csrc"
make -s -j8
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc -O3 -std=c++17 -lineinfo $ARCH -I../../include -I. -Xcompiler -fPIC,-fopenmp,-Wall --fmad=false -Xptxas -v $EXTRA -c kernels.cu -o /tmp/kernels_$NAME.o 2> /tmp/kernels_$NAME.log
grep -A1 "trace_kernelILi0ELb0" /tmp/kernels_$NAME.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo
nvcc -shared $ARCH -o ../libdprt_$NAME.so /tmp/kernels_$NAME.o partition.o epilogue.o mlp.o p2p_exchange.o dprt_api.o bvh_build.o scene_flatten.o -ldl -Xcompiler -fopenmp -lgomp
echo "built libdprt_$NAME.so"
