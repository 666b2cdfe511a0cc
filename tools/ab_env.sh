#!/bin/bash
# same-box A/B of an environment switch through bench.py: tools/ab_env.sh VAR v1 v2 ...
mkdir -p gpurun_out/r2b
VAR=$1; shift
for rep in 1 2; do for v in "$@"; do for rows in 10000000 1250000; do
  f=gpurun_out/r2b/env_${VAR}_${v}_${rows}_$rep
  env $VAR=$v python bench.py --rows $rows --steps 100 --warmup 10 --no-cpu-baseline --no-stages --parity-queries 2 > $f.json 2> $f.err
  python - <<PY
import json
try:
    d=json.loads(open('$f.json').read().strip().splitlines()[-1])
    r=d['roofline']; b=d['bm25_kernel']
    print('$VAR=$v', $rows, 'rep$rep', 'q/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'two', round(d['two_users_per_step']['value'],1), 'dense in/alone', round(r['ms_per_launch'],4), round(r['alone']['ms_per_launch'],4), 'bm25 in/alone', round(b['in_step']['ms_per_launch'],4), round(b['alone']['ms_per_launch'],4), 'pass', round(r['pass_ms_per_step'],4), 'parity', d['parity']['mismatches'])
except Exception as e:
    print('$VAR=$v', $rows, 'FAILED', e)
PY
done; done; done
