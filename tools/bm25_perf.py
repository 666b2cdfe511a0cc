#!/usr/bin/env python
"""BM25 stage timing at the C4 shape: python tools/bm25_perf.py [--rows 10000000]"""
import argparse, json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from legal_rag_engine_b200 import synth
from legal_rag_engine_b200.device_index import DeviceIndex
ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--K", type=int, default=20)
a = ap.parse_args()
dev = DeviceIndex(0)
bm = synth.device_bm25(a.rows, dev.device, seed=777)
x = torch.zeros((8, 384), dtype=torch.float16, device=dev.device)
dev.set_corpus(x, 0); dev.n_local = a.rows
dev._ck(dev.lib.lrx_set_corpus(dev.h, x.data_ptr(), a.rows, 0, 384))   # only n_local matters for K3
dev.set_postings(bm["term_ptr"], bm["postings"], bm["doc_len"], bm["idf"], bm["avgdl"])
df = bm["df"].cpu().numpy()
POOL, NSUB, NT = 16, 4, 8
terms, _ = synth.host_query_terms(POOL * NSUB, NT, seed=999)
terms = terms.reshape(POOL, NSUB * NT)
ptr = torch.from_numpy((np.arange(NSUB + 1) * NT).astype(np.int32)).cuda()
t_dev = torch.from_numpy(terms).cuda()
cand = torch.randint(0, a.rows, (NSUB, 20), device="cuda")
for i in range(3):
    dev.bm25(t_dev[i % POOL], ptr, cand, a.K)
torch.cuda.synchronize()
dev.profile(True); dev.profile_read(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.iters):
    dev.bm25(t_dev[i % POOL], ptr, cand, a.K)
e1.record(); torch.cuda.synchronize()
ms, n = dev.profile_read(1)
by = float(np.mean([df[terms[i % POOL]].sum() * 8 for i in range(a.iters)]))
print(json.dumps({"rows": a.rows, "call_ms": e0.elapsed_time(e1) / a.iters, "scan_ms": ms / n,
                  "bytes": by, "GBps": by / (ms / n) / 1e6}))
dev.close()
