#!/usr/bin/env python
"""Text path (encoder in the call) with several batches in flight: lrx_search_text_host_begin / _end on
n handles over one index.  python tools/text_pipeline_perf.py [--rows 10000000] [--in-flight 3]"""
import argparse, json, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from legal_rag_engine_b200 import synth
from legal_rag_engine_b200.device_index import DeviceIndex
from legal_rag_engine_b200.encoder import SentenceEncoder
ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--in-flight", type=int, default=3)
ap.add_argument("--steps", type=int, default=200)
a = ap.parse_args()
dev = DeviceIndex(0)
x = synth.device_vectors(a.rows, dev.device, seed=1234)
dev.set_corpus(x, 0)
bm = synth.device_bm25(a.rows, dev.device, seed=777)
dev.set_postings(bm["term_ptr"], bm["postings"], bm["doc_len"], bm["idf"], bm["avgdl"])
devs = [dev] + [dev.clone_view() for _ in range(a.in_flight - 1)]
sd = synth.bert_state_dict(42, 0.02)
encs = [SentenceEncoder(d, state_dict=sd) for d in devs]
streams = [torch.cuda.Stream() for _ in devs]
for d, s in zip(devs, streams):
    with torch.cuda.stream(s):
        d.use_current_stream()
POOL, NSUB, NT, S = 16, 4, 8, 32
terms, _ = synth.host_query_terms(POOL * NSUB, NT, seed=999)
lists = [[terms[(p * NSUB + b) * NT:(p * NSUB + b + 1) * NT].tolist() for b in range(NSUB)] for p in range(POOL)]
toks = [synth.token_batch(NSUB, S, seed=100 + p) for p in range(POOL)]
W = [0.5, 0.6, 0.5, 0.6]
def run(n, pipelined):
    res = None
    for i in range(n):
        d = devs[i % len(devs)] if pipelined else devs[0]
        ids, lens = toks[i % POOL]
        if pipelined:
            if i >= len(devs):
                res = d.search_host_end()
            d.search_text_host_begin(ids, lens, lists[i % POOL], 10, W, "rrf")
        else:
            res = d.search_text_host(ids, lens, lists[i % POOL], 10, W, "rrf")
    if pipelined:
        for i in range(max(0, n - len(devs)), n):
            res = devs[i % len(devs)].search_host_end()
    return res
out = {"rows": a.rows, "in_flight": a.in_flight}
for name, pipe in (("blocking", False), ("pipelined", True)):
    run(12, pipe); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(a.steps, pipe); torch.cuda.synchronize()
    out[name + "_qps"] = round(a.steps / (time.perf_counter() - t0), 1)
print(json.dumps(out))
