#!/bin/bash
# eager launch list of one CU-Net-8 training step (our kernels only; three warm-up steps skipped)
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(conv_|stem_|mse_|decode_|pack_|rmsprop|bn_)' \
   --launch-skip ${SKIP:-1200} --launch-count ${COUNT:-1300} --csv \
   --log-file gpurun_out/launches_cunet8_eager.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/launches_run.log 2>&1
wc -l gpurun_out/launches_cunet8_eager.csv
