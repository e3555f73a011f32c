#!/bin/bash
# Round-end evidence in one gpurun call: GPU tests, smoke, bench lines (both arms), eager launch list of one CU-Net-8
# step, ncu --set full of the five conv kernels on the bench shapes.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "tests rc=$?" | tee gpurun_out/rc_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/rc_final.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_cunet8.json 2> gpurun_out/bench_cunet8.err; echo "bench8 rc=$?" | tee -a gpurun_out/rc_final.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_cunet8_reference.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" | tee -a gpurun_out/rc_final.log
timeout 300 python bench.py --config cunet2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cunet2_fp32.json 2> gpurun_out/bench_cunet2.err; echo "bench2 rc=$?" | tee -a gpurun_out/rc_final.log
timeout 300 python bench.py --config cunet16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cunet16.json 2> gpurun_out/bench_cunet16.err; echo "bench16 rc=$?" | tee -a gpurun_out/rc_final.log
bash tools/launch_list.sh
for spec in "dgrad:time_dgrad.py:conv_dgrad_v2" "wgrad:time_wgrad.py:conv_wgrad_v2" "fwd:time_fwd.py:conv_fwd_kernel" "bwd3x3:time_bwd3x3.py:conv_bwd3x3" "fwd3x3:time_fwd.py:conv_fwd3x3"; do
  k=${spec%%:*}; rest=${spec#*:}; script=${rest%%:*}; kern=${rest#*:}
  c=up; if [ "$k" = "fwd3x3" ]; then c=3x3; fi
  CASE=$c timeout 200 ncu --set full --clock-control none --import-source on -k regex:$kern -c 1 -f -o gpurun_out/ncu_r1_${k} python tools/$script > gpurun_out/ncu_${k}.log 2>&1
done
for s in time_fwd.py time_dgrad.py time_bwd3x3.py; do python tools/$s 2>&1 | grep -E " us|us " ; done > gpurun_out/time_ops.log
CASE=3x3 python tools/time_fwd.py >> gpurun_out/time_ops.log 2>&1
python tools/time_small.py >> gpurun_out/time_ops.log 2>&1
cat gpurun_out/rc_final.log; tail -n 3 gpurun_out/t_all.log; cat gpurun_out/time_ops.log
