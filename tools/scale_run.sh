#!/bin/bash
# usage: tools/scale_run.sh N [c5]   -- the driver's scaling command at N GPUs (and config C5 at N = 8)
mkdir -p gpurun_out/r2g
N=$1
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 "${@:3}"; }
tr $N 29601 --steps 20 --warmup 5 > gpurun_out/r2g/scale_n$N.json 2> gpurun_out/r2g/scale_n$N.err
if [ "$2" = "c5" ]; then
  tr $N 29604 --steps 20 --warmup 5 --rows 100000000 --k 100 > gpurun_out/r2g/c5_n$N.json 2> gpurun_out/r2g/c5_n$N.err
fi
if [ "$2" = "tests" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2g/pytest_multi_n$N.log
fi
