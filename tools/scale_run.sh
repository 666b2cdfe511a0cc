#!/bin/bash
# One 8-GPU box: the driver's scaling commands (N = 8, 4, 2) and config C5 (100 M rows, top-100).
mkdir -p gpurun_out/r2
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 "${@:3}"; }
tr 8 29601 --steps 20 --warmup 5 > gpurun_out/r2/scale_n8.json 2> gpurun_out/r2/scale_n8.err
tr 4 29602 --steps 20 --warmup 5 > gpurun_out/r2/scale_n4.json 2> gpurun_out/r2/scale_n4.err
tr 2 29603 --steps 20 --warmup 5 > gpurun_out/r2/scale_n2.json 2> gpurun_out/r2/scale_n2.err
tr 8 29604 --steps 20 --warmup 5 --rows 100000000 --k 100 > gpurun_out/r2/c5_n8.json 2> gpurun_out/r2/c5_n8.err
tr 8 29605 --steps 20 --warmup 5 --in-flight 3 > gpurun_out/r2/scale_n8_f3.json 2> gpurun_out/r2/scale_n8_f3.err
nvidia-smi topo -m > gpurun_out/r2/topo.txt 2>&1
