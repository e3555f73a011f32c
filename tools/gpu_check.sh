#!/bin/bash
# targeted tests, the full GPU suite (PDL off / on), bench (PDL off / on, fwd3x3 off)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv_fwd.py -q -k "3x3 and bf16" > gpurun_out/t_fwd3.log 2>&1; echo "fwd3 rc=$?" | tee gpurun_out/rc.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "all rc=$?" | tee -a gpurun_out/rc.log
CUNET_PDL=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_all_pdl.log 2>&1; echo "all_pdl rc=$?" | tee -a gpurun_out/rc.log
for v in "base:" "pdl:CUNET_PDL=1" "nof3:CUNET_FWD3X3_OFF=1" "pdl_eager:CUNET_PDL=1 EAGER=1"; do
  name=${v%%:*}; envs=${v#*:}
  extra=""; if [[ "$envs" == *EAGER* ]]; then extra="--no-graph"; fi
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $extra > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name rc=$?" | tee -a gpurun_out/rc.log
  python -c "
import json;d=json.load(open('gpurun_out/bench_$name.json'));print('$name',d['value'],d['ms_per_step'],d['e2e']['value'])" | tee -a gpurun_out/rc.log
done
tail -n 5 gpurun_out/t_fwd3.log gpurun_out/t_all.log gpurun_out/t_all_pdl.log
