#!/usr/bin/env python
"""K2b stage timing (config C3): python tools/dense_batched_perf.py [--rows 1000000] [--B 1024]"""
import argparse, json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from legal_rag_engine_b200 import synth
from legal_rag_engine_b200.device_index import DeviceIndex
ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--B", type=int, nargs="+", default=[64, 256, 1024, 4096])
ap.add_argument("--K", type=int, default=20)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
dev = DeviceIndex(0)
x = synth.device_vectors(a.rows, dev.device, seed=1234)
dev.set_corpus(x, 0)
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
peak = peaks.get("bf16_tflops", 1590.0)
for B in a.B:
    q = torch.from_numpy(synth.host_queries(B, seed=4321)).cuda()
    for _ in range(3):
        out = dev.dense_topk_batched(q, a.K)
    torch.cuda.synchronize()
    assert int(out[3].sum().item()) == 0 or __import__("os").environ.get("LRX_LIB")
    dev.profile(True); dev.profile_read(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        dev.dense_topk_batched(q, a.K)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    kms, kn = dev.profile_read(0)
    dev.profile(False)
    Bp = (B + 127) // 128 * 128
    tiles = (a.rows + 255) // 256
    stride = max(1, min(8, tiles * 8 // (16 * a.K)))
    if (tiles + stride - 1) // stride * 8 > 16384:
        stride = (tiles * 8 + 16383) // 16384
    flop = 2.0 * Bp * 384 * a.rows * (1 + 1.0 / stride)
    print(json.dumps({"rows": a.rows, "B": B, "K": a.K, "call_ms": round(ms, 4), "queries_per_s": round(B / ms * 1e3),
                      "gemm_kernels_ms": round(kms / a.iters, 4), "tflops_gemm": round(flop / (kms / a.iters) / 1e9, 1),
                      "frac_of_measured_bf16_peak": round(flop / (kms / a.iters) / 1e9 / peak, 4),
                      "useful_tflops_call": round(2.0 * B * 384 * a.rows / ms / 1e9, 1)}), flush=True)
dev.close()
