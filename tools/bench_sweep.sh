#!/bin/bash
# One GPU call: rebuild dense.cu per variant (nvcc -D flags) and run the headline bench at 10 M rows.
for v in "$@"; do
  LRX_ONLY=dense.cu LRX_EXTRA_NVCC="$v" python legal-rag-engine_b200/build.py > /dev/null 2>&1 || echo "build failed: $v"
  python bench.py --no-cpu-baseline --no-stages --steps 100 2>/dev/null | tail -1 > /tmp/b.json
  echo "$v :: $(python -c "import json; d=json.load(open('/tmp/b.json')); print(round(d['value'],1), round(d['ms_per_step'],4), 'scan', round(d['roofline']['ms_per_launch'],4), 'bm25', round(d['bm25_kernel']['ms_per_launch'],4))")"
done
LRX_ONLY=dense.cu python legal-rag-engine_b200/build.py > /dev/null 2>&1
