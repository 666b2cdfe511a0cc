#!/bin/bash
# A/B of the BM25 scan variants built by tools/bm25_variants.sh
mkdir -p gpurun_out/r2
for v in b0h0 b0h1 b1h0 b1h1 b2h0 b2h1; do
  for cfg in "1250000 20" "10000000 20" "12500000 200" "1250000 200"; do
    set -- $cfg
    echo -n "$v rows=$1 K=$2 " 
    LRX_LIB=tools/ab/liblrx_$v.so python tools/bm25_perf.py --rows $1 --K $2 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['scan_ms'],4), round(d['call_ms'],4))"
  done
done
