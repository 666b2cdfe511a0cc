mkdir -p gpurun_out/r2b
timeout 600 python -m pytest tests/test_gpu_prefilter.py -x -q > gpurun_out/r2b/t_prefilter.log 2>&1; echo "prefilter rc=$?" 
timeout 300 python tools/scan_perf.py 10000000 1250000 > gpurun_out/r2b/scan_perf.txt 2>&1; echo "scan_perf rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2b/t_parity.log 2>&1; echo "parity rc=$?"
tail -5 gpurun_out/r2b/t_prefilter.log; cat gpurun_out/r2b/scan_perf.txt | tail -3; tail -3 gpurun_out/r2b/t_parity.log
