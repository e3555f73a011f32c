#!/bin/bash
# A/B of library variants: bash tools/gpu_ab.sh name:libfile ...
mkdir -p gpurun_out
for v in "$@"; do
  name=${v%%:*}; lib=${v#*:}
  CUNET_LIB=$PWD/cu-net_b200/$lib timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name rc=$?" | tee -a gpurun_out/rc_ab.log
  python -c "
import json;d=json.load(open('gpurun_out/bench_$name.json'));r=d['roofline_all'];print('$name',d['value'],d['ms_per_step'],r['fwd']['us'],r['dgrad']['us'],r['wgrad']['us'])" | tee -a gpurun_out/rc_ab.log
done
