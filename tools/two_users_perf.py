#!/usr/bin/env python
"""Hybrid search with TWO user queries (8 sub-queries) per launch chain against one (4 sub-queries):
user queries per second, host buffers in and out, three batches in flight.
python tools/two_users_perf.py [--rows 10000000]"""
import argparse, json, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from legal_rag_engine_b200 import synth
from legal_rag_engine_b200.device_index import DeviceIndex
ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--steps", type=int, default=150)
a = ap.parse_args()
dev = DeviceIndex(0)
x = synth.device_vectors(a.rows, dev.device, seed=1234)
dev.set_corpus(x, 0)
bm = synth.device_bm25(a.rows, dev.device, seed=777)
dev.set_postings(bm["term_ptr"], bm["postings"], bm["doc_len"], bm["idf"], bm["avgdl"])
devs = [dev, dev.clone_view(), dev.clone_view()]
streams = [torch.cuda.Stream() for _ in devs]
for d, s in zip(devs, streams):
    with torch.cuda.stream(s):
        d.use_current_stream()
POOL, NT = 16, 8
out = {"rows": a.rows}
for users in (1, 2):
    B = 4 * users
    terms, _ = synth.host_query_terms(POOL * B, NT, seed=999)
    lists = [[terms[(p * B + b) * NT:(p * B + b + 1) * NT].tolist() for b in range(B)] for p in range(POOL)]
    qh = [synth.host_queries(B, seed=4321 + p) for p in range(POOL)]
    W = [0.5, 0.6, 0.5, 0.6] * users
    def run(n):
        for i in range(n):
            d = devs[i % 3]
            if i >= 3:
                d.search_host_end()
            d.search_host_begin(qh[i % POOL], lists[i % POOL], 10, W, "rrf")
        for i in range(max(0, n - 3), n):
            devs[i % 3].search_host_end()
    run(9); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(a.steps); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[f"users_per_chain_{users}"] = {"user_queries_per_s": round(a.steps * users / dt, 1), "ms_per_chain": round(dt / a.steps * 1e3, 4)}
print(json.dumps(out))
