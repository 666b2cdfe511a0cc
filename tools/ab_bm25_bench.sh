#!/bin/bash
# same-box A/B of two library builds through bench.py: 10 M rows (N = 1) and the 8-GPU shard size
mkdir -p gpurun_out/r2
for rep in 1 2; do
for lib in base new; do
  if [ $lib = base ]; then export LRX_LIB=tools/ab/liblrx_base.so; else unset LRX_LIB; fi
  for rows in 10000000 1250000; do
    python bench.py --rows $rows --steps 100 --warmup 10 --no-cpu-baseline --no-stages --parity-queries 2 > gpurun_out/r2/ab_${lib}_${rows}_$rep.json 2> gpurun_out/r2/ab_${lib}_${rows}_$rep.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2/ab_${lib}_${rows}_$rep.json').read().strip().splitlines()[-1])
    r=d['roofline']; b=d['bm25_kernel']
    print('$lib', $rows, 'rep$rep', 'q/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'dense in/alone', round(r['ms_per_launch'],4), round(r['alone']['ms_per_launch'],4), 'bm25 in/alone', round(b['in_step']['ms_per_launch'],4), round(b['alone']['ms_per_launch'],4), 'parity', d['parity']['mismatches'])
except Exception as e:
    print('$lib', $rows, 'FAILED', e)
PY
  done
done
done
