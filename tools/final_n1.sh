#!/bin/bash
# Round-end evidence on one GPU: GPU tests, the default bench line, the ncu launch list of the same
# command and one --set full capture of the two scans.
mkdir -p gpurun_out/r2g
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2g/pytest_gpu.log
python bench.py > gpurun_out/r2g/bench_n1.json 2> gpurun_out/r2g/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g/bench_ref.json 2> gpurun_out/r2g/bench_ref.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stages --parity-queries 0 --min-time 0"
$B > gpurun_out/r2g/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dense_|bm25|merge|fuse|rescore|pack_|finalize" -s 720 -c 28 --csv --log-file gpurun_out/r2g/r2_launches_bench.csv $B > gpurun_out/r2g/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"dense_scan|bm25_scan" -s 200 -c 2 -o gpurun_out/r2g/prof_r2_scan $B > gpurun_out/r2g/ncu_f.log 2>&1
