#!/usr/bin/env python
"""One pass over the tensor-core kernels for ncu: encoder (two full chunks at S=128) and the
batched dense scorer (B=1024 over 1 M rows).  `--warm N` untimed passes first."""
import argparse, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from legal_rag_engine_b200 import synth
from legal_rag_engine_b200.device_index import DeviceIndex
from legal_rag_engine_b200.encoder import SentenceEncoder
ap = argparse.ArgumentParser()
ap.add_argument("--warm", type=int, default=1)
a = ap.parse_args()
dev = DeviceIndex(0)
enc = SentenceEncoder(dev, state_dict=synth.bert_state_dict(42, 0.02))
ids, lens = synth.token_batch(148, 128, seed=1, full=True)
d_ids, d_lens = torch.from_numpy(ids).cuda(), torch.from_numpy(lens).cuda()
x = synth.device_vectors(1_000_000, dev.device, seed=1234)
dev.set_corpus(x, 0)
q = torch.from_numpy(synth.host_queries(1024, seed=4321)).cuda()
ids_q, lens_q = synth.token_batch(4, 32, seed=2, full=True)           # a fan-out query batch: the one-launch cluster kernel
dq_ids, dq_lens = torch.from_numpy(ids_q).cuda(), torch.from_numpy(lens_q).cuda()
for _ in range(a.warm + 1):
    enc.encode_ids_device(d_ids, d_lens)
    dev.dense_topk_batched(q, 20)
    enc.encode_ids_device(dq_ids, dq_lens)
torch.cuda.synchronize()
print("ok")
dev.close()
