#!/bin/bash
# A/B of the launch chain at the 8-GPU shard size on one GPU (rows 1.25 M, two batches in flight)
mkdir -p gpurun_out/r2
run() { # name, env..., -- args
  name=$1; shift
  env "$@" python bench.py --rows 1250000 --steps 400 --warmup 40 --in-flight 2 --no-cpu-baseline --no-stages --parity-queries 0 $EXTRA > gpurun_out/r2/ab_$name.json 2> gpurun_out/r2/ab_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2/ab_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'scan', round(d['roofline']['ms_per_launch'],4), 'bm25', round(d['bm25_kernel']['in_step']['ms_per_launch'],4), 'bm25 alone', round(d['bm25_kernel']['alone']['ms_per_launch'],4), 'host', round(d['host_enqueue_ms_per_step'],4))
except Exception as e:
    print('$name', 'FAILED', e)
PY
}
run f32_graph X=1
run f32_eager LRX_NO_GRAPH=1
EXTRA="--kernel-events off" run f32_graph_noev X=1
run f64_graph LRX_LIB=tools/ab/liblrx_f64.so
run f64_eager LRX_LIB=tools/ab/liblrx_f64.so LRX_NO_GRAPH=1
EXTRA="--kernel-events off" run f64_graph_noev LRX_LIB=tools/ab/liblrx_f64.so
EXTRA="--kernel-events off --in-flight 3" run f64_graph_noev_f3 LRX_LIB=tools/ab/liblrx_f64.so
