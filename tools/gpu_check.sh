#!/bin/bash
# usage: bash tools/gpu_check.sh "<pytest -k expr or empty>"   -- targeted tests, then the full GPU suite, then the bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv_bwd.py -x -q -k "fused" > gpurun_out/t_fused.log 2>&1; echo "fused rc=$?" | tee gpurun_out/rc.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/rc.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cunet8.json 2> gpurun_out/bench_cunet8.err; echo "bench rc=$?" | tee -a gpurun_out/rc.log
tail -3 gpurun_out/t_fused.log gpurun_out/t_all.log
python -c "
import json;d=json.load(open('gpurun_out/bench_cunet8.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'])"
