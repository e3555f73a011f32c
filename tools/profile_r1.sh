#!/bin/bash
# One gpurun call: bench lines, eager launch list of one CU-Net-8 step, ncu --set full of the three conv kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cunet8.json 2> gpurun_out/bench_cunet8.err
timeout 300 python bench.py --config cunet2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cunet2_fp32.json 2> gpurun_out/bench_cunet2.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cunet --launch-skip ${SKIP:-2400} --launch-count ${COUNT:-800} --csv \
   --log-file gpurun_out/launches_cunet8_eager.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/launches_run.log 2>&1
for k in fwd dgrad wgrad; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_${k} -c 1 -f -o gpurun_out/ncu_${k}_320to128 python tools/time_${k}.py > gpurun_out/ncu_${k}.log 2>&1
done
python tools/time_fwd.py > gpurun_out/time_fwd.log 2>&1
python tools/time_small.py > gpurun_out/time_small.log 2>&1
ls -la gpurun_out
