#!/bin/bash
# builds variants of the BM25 scan (bootstrap / histogram knobs) into tools/ab/ for A/B runs
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/ab
for v in "0 0" "0 1" "1 0" "1 1" "2 0" "2 1"; do
  set -- $v
  LRX_EXTRA_NVCC="-DLRX_BM_BOOT=$1 -DLRX_BM_HIST=$2" LRX_ONLY=bm25.cu python legal-rag-engine_b200/build.py > /dev/null
  cp legal-rag-engine_b200/csrc/liblrx.so tools/ab/liblrx_b$1h$2.so
done
LRX_ONLY=bm25.cu python legal-rag-engine_b200/build.py > /dev/null
