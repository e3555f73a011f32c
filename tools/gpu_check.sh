#!/bin/bash
# targeted tests, the full GPU suite, bench variants
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv_fwd.py tests/test_gpu_conv_bwd.py -q -k "bf16 or fused" > gpurun_out/t_conv.log 2>&1; echo "conv rc=$?" | tee gpurun_out/rc.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/rc.log
for v in "base:" "f2big:CUNET_FWD_V2_MIN_TILES=148" "f2off:CUNET_FWD_V2_MIN_TILES=-1"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name rc=$?" | tee -a gpurun_out/rc.log
  python -c "
import json;d=json.load(open('gpurun_out/bench_$name.json'));print('$name',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline_all']['fwd']['us'])" | tee -a gpurun_out/rc.log
done
tail -n 15 gpurun_out/t_conv.log; tail -n 5 gpurun_out/t_all.log
