#!/bin/bash
# Round-2 evidence in one gpurun call: GPU tests, smoke, bench lines of all four BASELINE configs (+ the reference arm),
# eager launch list of one CU-Net-8 step on the final kernels, ncu --set full of the two new 1x1 kernels on the bench op,
# per-op timings and in-kernel timelines.  Everything lands in gpurun_out/ (copy what is to be judged into profiles/).
mkdir -p gpurun_out
O=gpurun_out/r2final
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
(time timeout 900 python -m pytest tests -m gpu -q -rf 2>&1 | tail -15) > $O/pytest_gpu.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.log
for c in cunet8 cunet2 cunet8bin cunet16; do
  timeout 500 python bench.py --config $c --steps 20 --warmup 5 > $O/bench_$c.json 2> $O/bench_$c.err; echo "bench $c rc=$?" >> $O/rc.log
done
timeout 400 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_cunet8_reference.json 2> $O/bench_ref.err; echo "ref rc=$?" >> $O/rc.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(conv_|stem_|mse_|decode_|pack_|rmsprop|bn_|quant_)' \
   --launch-skip ${SKIP:-1200} --launch-count ${COUNT:-1300} --csv --log-file $O/launches_cunet8_eager.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-loss-check > $O/launches_run.log 2>&1
python tools/launch_summary.py $O/launches_cunet8_eager.csv $O/launches_cunet8_eager_onestep.csv > $O/launches_summary.txt 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_bwd1x1 --launch-skip 4 --launch-count 1 -f \
   -o $O/ncu_r2_bwd1x1_320up64 python tools/time_bwd1x1.py 320up64 > $O/ncu_bwd1x1.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_fwd_v3 --launch-skip 4 --launch-count 1 -f \
   -o $O/ncu_r2_fwd_v3_320up64 python tools/time_fwd_v3.py 320up64 > $O/ncu_fwd_v3.log 2>&1
timeout 300 python tools/time_bwd1x1.py > $O/time_bwd1x1.log 2>&1
timeout 300 python tools/time_fwd_v3.py > $O/time_fwd_v3.log 2>&1
timeout 200 python tools/time_bwd3x3.py > $O/time_3x3.log 2>&1
CASE=3x3 timeout 200 python tools/time_fwd.py >> $O/time_3x3.log 2>&1
CUNET_LIB=$PWD/cu-net_b200/libcunet_b200_trace.so timeout 200 python tools/time_bwd1x1.py trace 320up64 > $O/trace_bwd1x1_320up64.log 2>&1
CUNET_LIB=$PWD/cu-net_b200/libcunet_b200_trace.so timeout 200 python tools/time_fwd_v3.py trace > $O/trace_fwd_v3_320up64.log 2>&1
cat $O/rc.log; tail -4 $O/pytest_gpu.log; cat $O/time_bwd1x1.log | head -4
