mkdir -p gpurun_out/mg
run() { # N config extra...
  N=$1; C=$2; shift 2
  if [ "$N" = "1" ]; then timeout 400 python bench.py --gpus 1 --config $C --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/mg/${C}_n${N}.json 2> gpurun_out/mg/${C}_n${N}.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --config $C --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/mg/${C}_n${N}.json 2> gpurun_out/mg/${C}_n${N}.err; fi
  echo "$C N=$N rc=$?" >> gpurun_out/mg/rc.log
}
nvidia-smi -L > gpurun_out/mg/gpus.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 tools/dp_check.py > gpurun_out/mg/dp_check.log 2>&1
run 8 cunet8
for N in 4 2 1; do run $N cunet8 --no-loss-check; done
for N in 8 4 2 1; do run $N cunet2 --no-loss-check; done
run 8 cunet8bin --no-loss-check
run 8 cunet16 --no-loss-check
cat gpurun_out/mg/rc.log; tail -3 gpurun_out/mg/dp_check.log
