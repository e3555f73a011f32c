mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_conv_bwd.py -q -rf -k "3x3" 2>&1 | tail -4) > gpurun_out/r2p_pytest.log 2>&1
for r in 4 8 16 32 64; do RES=$r timeout 100 python tools/time_bwd3x3.py 2>&1 | tail -1; done > gpurun_out/r2p_bwd3x3.log
RES=16 CUNET_PDL=0 CUNET_LIB=$PWD/cu-net_b200/libcunet_b200_trace.so timeout 100 python tools/trace_kernels.py bwd3x3 2>&1 | tail -3 >> gpurun_out/r2p_bwd3x3.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-loss-check > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
cat gpurun_out/r2p_pytest.log gpurun_out/r2p_bwd3x3.log; python -c "
import json; d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
