#!/bin/bash
mkdir -p gpurun_out/r2
for rep in 1 2; do
for lib in base new d4; do
  if [ $lib = new ]; then unset LRX_LIB; else export LRX_LIB=tools/ab/liblrx_$lib.so; fi
  for cfg in "10000000 20" "1250000 20"; do set -- $cfg
    echo -n "$lib rows=$1 K=$2 rep$rep "; python tools/bm25_perf.py --rows $1 --K $2 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['scan_ms'],4))"
  done
done
done
for lib in new d4 new d4; do
  if [ $lib = new ]; then unset LRX_LIB; else export LRX_LIB=tools/ab/liblrx_$lib.so; fi
  for rows in 10000000 1250000; do
  python bench.py --rows $rows --steps 100 --warmup 10 --no-cpu-baseline --no-stages --parity-queries 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; b=d['bm25_kernel']
print('$lib', $rows, 'q/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'dense in/alone', round(r['ms_per_launch'],4), round(r['alone']['ms_per_launch'],4), 'bm25 in/alone', round(b['in_step']['ms_per_launch'],4), round(b['alone']['ms_per_launch'],4), 'parity', d['parity']['mismatches'])"
  done
done
