mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_bwd1x1 --launch-skip 4 --launch-count 1 -f -o gpurun_out/ncu_r2_bwd1x1_320up64 python tools/time_bwd1x1.py 320up64 > gpurun_out/r2g_ncu.log 2>&1
timeout 300 python tools/time_bwd1x1.py 320up64 192_64 > gpurun_out/r2g_time.log 2>&1
cat gpurun_out/r2g_time.log; tail -3 gpurun_out/r2g_ncu.log
