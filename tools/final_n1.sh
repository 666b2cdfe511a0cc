#!/bin/bash
# Round-end evidence on one GPU: GPU tests, smoke, the default bench line, the reference arm.
# (ncu launch list + --set full capture of the same chain: tools/r2b_profile.sh)
mkdir -p gpurun_out/r2g
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2g/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/r2g/bench_n1.json 2> gpurun_out/r2g/bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g/bench_ref.json 2> gpurun_out/r2g/bench_ref.err; echo "ref rc=$?"
python bench.py --prefilter off --no-stages --no-cpu-baseline > gpurun_out/r2g/bench_n1_f16.json 2>> gpurun_out/r2g/bench_n1.err; echo "f16 rc=$?"
tail -3 gpurun_out/r2g/pytest_gpu.log; tail -2 gpurun_out/r2g/smoke.log
