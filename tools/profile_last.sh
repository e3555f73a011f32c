#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests -m gpu -x -q > gpurun_out/t_all_final.log 2>&1; echo "tests rc=$?" > gpurun_out/rc_final.log
timeout 40 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cunet8_final.json 2> gpurun_out/bench_cunet8_final.err; echo "bench8 rc=$?" >> gpurun_out/rc_final.log
timeout 40 ncu --set full --clock-control none --import-source on -k regex:conv_dgrad_v2 -c 1 -f -o gpurun_out/ncu_r1_dgrad_inplace python tools/time_dgrad.py > gpurun_out/ncu_dgrad_final.log 2>&1
timeout 20 python tools/time_bwd3x3.py > gpurun_out/time_ops_final.log 2>&1
CASE=3x3 timeout 20 python tools/time_fwd.py >> gpurun_out/time_ops_final.log 2>&1
cat gpurun_out/rc_final.log; tail -n 3 gpurun_out/t_all_final.log; cat gpurun_out/time_ops_final.log
