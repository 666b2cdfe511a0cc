#!/usr/bin/env python
"""K2a (streaming scan, 4 queries per matrix pass) against K2b (tensor-core batched) per call, for the
batch sizes in between: python tools/k2_crossover.py  -> one JSON line per (rows, B)."""
import json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from legal_rag_engine_b200 import synth
from legal_rag_engine_b200.device_index import DeviceIndex
dev = DeviceIndex(0)
x = synth.device_vectors(10_000_000, dev.device, seed=1234)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timed(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rows in (1_000_000, 10_000_000):
    dev.set_corpus(x[:rows], 0)
    for B in (4, 5, 8, 12, 16, 32, 64):
        q = torch.from_numpy(synth.host_queries(B, seed=B)).to(dev.device)
        a = timed(lambda: dev.dense_topk(q, 20))
        b = timed(lambda: dev.dense_topk_batched(q, 20))
        ra, rb = dev.dense_topk(q, 20), dev.dense_topk_batched(q, 20)
        same = bool((ra[2] == rb[2]).all().item() and (ra[0] == rb[0]).all().item())
        print(json.dumps({"rows": rows, "B": B, "k2a_ms": round(a, 4), "k2b_ms": round(b, 4), "same": same,
                          "k2b_flags": int(rb[3].sum().item())}))
dev.close()
