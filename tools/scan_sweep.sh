#!/bin/bash
# One GPU call: rebuild dense.cu per variant (nvcc -D flags) and time the scan kernel.
# usage: bash tools/scan_sweep.sh "-DLRX_SCAN_STAGES=2" "-DLRX_SCAN_STAGES=3" ...
for v in "$@"; do
  LRX_ONLY=dense.cu LRX_EXTRA_NVCC="$v" python legal-rag-engine_b200/build.py > /dev/null 2>&1 || echo "build failed: $v"
  echo "$v :: $(python tools/scan_perf.py 10000000 1250000 2>&1 | tail -1)"
done
LRX_ONLY=dense.cu python legal-rag-engine_b200/build.py > /dev/null 2>&1
