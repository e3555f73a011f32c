mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_conv_bwd.py -q -rf -k "quantized or bwd1x1" 2>&1 | tail -4) > gpurun_out/r2n_pytest.log 2>&1
timeout 300 python tools/time_bwd1x1.py 320up64 288up64 192_64 256_64 head68_64 320up32 320up16 320up64_b3 > gpurun_out/r2n_time_bwd.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-loss-check > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
cat gpurun_out/r2n_pytest.log gpurun_out/r2n_time_bwd.log; python -c "
import json; d=json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
