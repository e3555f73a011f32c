#!/bin/bash
# Build the timeline variant of the library (-DCUNET_TRACE: in-kernel clock64() marks, common.cuh) next to the product:
#   bash tools/build_trace.sh && CUNET_LIB=$PWD/cu-net_b200/libcunet_b200_trace.so python tools/trace_kernels.py
set -e
cd "$(dirname "$0")/.."
out=cu-net_b200/csrc/build/trace
mkdir -p $out
for f in cu-net_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DCUNET_TRACE -c $f -o $out/$(basename $f .cu).o &
done
wait
nvcc -shared -o cu-net_b200/libcunet_b200_trace.so $out/*.o -gencode arch=compute_100a,code=sm_100a
echo built cu-net_b200/libcunet_b200_trace.so
