#!/usr/bin/env python
"""Encoder stage timing (config C2): sequences/s and tensor-pipe roofline fraction.
    python tools/enc_perf.py [--B 1024] [--S 128] [--iters 5]"""
import argparse, json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from legal_rag_engine_b200 import synth
from legal_rag_engine_b200.device_index import DeviceIndex
from legal_rag_engine_b200.encoder import SentenceEncoder

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, nargs="+", default=[1, 64, 1024, 4096])
ap.add_argument("--S", type=int, nargs="+", default=[128, 256])
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--half-pad", action="store_true")
a = ap.parse_args()
dev = DeviceIndex(0)
enc = SentenceEncoder(dev, state_dict=synth.bert_state_dict(42, 0.02))
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
peak = peaks.get("bf16_tflops", 1590.0)
for S in a.S:
    for B in a.B:
        ids, lens = synth.token_batch(B, S, seed=1, full=not a.half_pad)
        if a.half_pad:
            lens[:] = S // 2; ids[:, S // 2:] = 0; ids[:, S // 2 - 1] = 102
        d_ids, d_lens = torch.from_numpy(ids).cuda(), torch.from_numpy(lens).cuda()
        for _ in range(3):
            enc.encode_ids_device(d_ids, d_lens)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            enc.encode_ids_device(d_ids, d_lens)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        flop = B * S * (6 * (2 * 384 * 1152 + 2 * 384 * 384 + 2 * 2 * 384 * 1536) + 6 * 4 * S * 384)
        print(json.dumps({"B": B, "S": S, "ms": round(ms, 4), "seq_per_s": round(B / ms * 1e3, 1),
                          "tflops": round(flop / ms / 1e9, 1), "frac_of_measured_bf16_peak": round(flop / ms / 1e9 / peak, 4)}), flush=True)
dev.close()
