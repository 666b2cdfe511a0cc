#!/usr/bin/env python
"""Dense scan kernel timing: python tools/scan_perf.py [rows ...]  (events inside the library)"""
import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from legal_rag_engine_b200 import synth
from legal_rag_engine_b200.device_index import DeviceIndex
rows_list = [int(a) for a in sys.argv[1:]] or [10_000_000, 1_250_000]
dev = DeviceIndex(0)
out = {}
for rows in rows_list:
    x = synth.device_vectors(rows, dev.device, seed=1234)
    q = torch.from_numpy(synth.host_queries(4, seed=4321)).cuda()
    for pre in (True, False):                      # int8 shadow scan / plain fp16 scan
        dev.set_corpus(x, 0, prefilter=pre)
        for _ in range(5):
            res = dev.dense_topk(q, 20)
        torch.cuda.synchronize()
        dev.profile(True); dev.profile_read(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            dev.dense_topk(q, 20)
        e1.record(); torch.cuda.synchronize()
        ms, n = dev.profile_read(0); dev.profile(False)
        row_bytes = 388 if pre else 768
        out[f"{rows}_{'q8' if pre else 'f16'}"] = {
            "scan_ms": round(ms / n, 4), "TBps": round(rows * row_bytes / (ms / n) / 1e9, 3),
            "call_ms": round(e0.elapsed_time(e1) / 30, 4), "flags": int(res[3].sum().item()),
            "bounds": dev.prefilter_bounds}
    del x
print(json.dumps(out))
dev.close()
