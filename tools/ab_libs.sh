#!/bin/bash
# same-box A/B of library builds through bench.py: tools/ab_libs.sh name1 name2 ... (tools/ab/liblrx_<name>.so;
# "cur" = the in-tree build), 10 M rows (N = 1) and the 8-GPU shard size
mkdir -p gpurun_out/r2b
for rep in 1 2; do
for lib in "$@"; do
  if [ $lib = cur ]; then unset LRX_LIB; else export LRX_LIB=tools/ab/liblrx_$lib.so; fi
  for rows in 10000000 1250000; do
    f=gpurun_out/r2b/ab_${lib}_${rows}_$rep
    python bench.py --rows $rows --steps 100 --warmup 10 --no-cpu-baseline --no-stages --parity-queries 2 > $f.json 2> $f.err
    python - <<PY
import json
try:
    d=json.loads(open('$f.json').read().strip().splitlines()[-1])
    r=d['roofline']; b=d['bm25_kernel']
    print('$lib', $rows, 'rep$rep', 'q/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'two', round(d['two_users_per_step']['value'],1), 'dense in/alone', round(r['ms_per_launch'],4), round(r['alone']['ms_per_launch'],4), 'bm25 in/alone', round(b['in_step']['ms_per_launch'],4), round(b['alone']['ms_per_launch'],4), 'parity', d['parity']['mismatches'])
except Exception as e:
    print('$lib', $rows, 'FAILED', e)
PY
  done
done
done
