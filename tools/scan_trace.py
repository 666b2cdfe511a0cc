#!/usr/bin/env python
"""Timeline of CTA 0 of dense_scan_kernel (clock64 stamps via lrx_debug_set_trace)."""
import sys, ctypes as C
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from legal_rag_engine_b200 import synth
from legal_rag_engine_b200.device_index import DeviceIndex
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
dev = DeviceIndex(0)
x = synth.device_vectors(rows, dev.device, seed=1234)
dev.set_corpus(x, 0)
q = torch.from_numpy(synth.host_queries(4, seed=4321)).cuda()
tr = torch.zeros(128, dtype=torch.int64, device="cuda")
for _ in range(3):
    dev.dense_topk(q, 20)
torch.cuda.synchronize()
dev._ck(dev.lib.lrx_debug_set_trace(dev.h, C.c_void_p(tr.data_ptr())))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); dev.dense_topk(q, 20); e1.record(); torch.cuda.synchronize()
dev._ck(dev.lib.lrx_debug_set_trace(dev.h, C.c_void_p(0)))
t = tr.cpu().numpy(); t0 = t[0]
us = lambda v: round((v - t0) / 1965.0, 2) if v else None
print("call (scan+merge+rescore) us:", round(e0.elapsed_time(e1) * 1e3, 1))
print("queries loaded", us(t[1]), "loop end", us(t[2]), "lists written", us(t[3]))
print("tile starts 0..19:", [us(v) for v in t[8:28]])
print("every 16th tile from 32:", [us(v) for v in t[30:40] if v])
dev.close()
