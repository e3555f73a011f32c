#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv_bwd.py -q -x > gpurun_out/t_bwd.log 2>&1; echo "bwd rc=$?" | tee gpurun_out/rc_dev.log
tail -n 3 gpurun_out/t_bwd.log
python tools/trace_dgrad.py > gpurun_out/trace_dgrad_inplace.log 2>&1; head -8 gpurun_out/trace_dgrad_inplace.log
bash tools/gpu_ab.sh prev:libcunet_b200_prev.so new:libcunet_b200.so
