#!/usr/bin/env python
"""Latency of the encoder at query size: the one-launch cluster kernel against the GEMM chain."""
import json, os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from legal_rag_engine_b200 import synth
from legal_rag_engine_b200.device_index import DeviceIndex
from legal_rag_engine_b200.encoder import SentenceEncoder
dev = DeviceIndex(0)
enc = SentenceEncoder(dev, state_dict=synth.bert_state_dict(42, 0.02))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for B, S in ((1, 16), (4, 16), (4, 32), (2, 64), (8, 16), (16, 32), (32, 32), (64, 32), (32, 64)):
    ids, lens = synth.token_batch(B, S, seed=1, full=True)
    d_ids, d_lens = torch.from_numpy(ids).cuda(), torch.from_numpy(lens).cuda()
    out = {"B": B, "S": S}
    for name, env in (("small_ms", None), ("chain_ms", "1")):
        if env: os.environ["LRX_NO_SMALL_ENCODER"] = env
        else: os.environ.pop("LRX_NO_SMALL_ENCODER", None)
        for _ in range(5): enc.encode_ids_device(d_ids, d_lens)
        torch.cuda.synchronize(); e0.record()
        for _ in range(50): enc.encode_ids_device(d_ids, d_lens)
        e1.record(); torch.cuda.synchronize()
        out[name] = round(e0.elapsed_time(e1) / 50, 4)
    print(json.dumps(out), flush=True)
dev.close()
