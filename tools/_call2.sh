#!/bin/bash
mkdir -p gpurun_out/mg2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/mg2/cunet8_n2.json 2> gpurun_out/mg2/cunet8_n2.err
python -c "import json;d=json.load(open('gpurun_out/mg2/cunet8_n2.json'));print('N=2 weak',d['value'],d['ms_per_step'],'strong',d['strong_scaling'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py > gpurun_out/mg2/dp_check_n2.log 2>&1
tail -4 gpurun_out/mg2/dp_check_n2.log
