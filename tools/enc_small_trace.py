#!/usr/bin/env python
"""Timeline of CTA 0 of encoder_small_kernel (clock64 stamps via lrx_debug_set_trace):
per layer 11 stamps = after QKV | barrier | attention | barrier | out-proj | barrier | LN1+up | barrier |
down | barrier | LN2."""
import ctypes as C, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from legal_rag_engine_b200 import synth
from legal_rag_engine_b200.device_index import DeviceIndex
from legal_rag_engine_b200.encoder import SentenceEncoder
dev = DeviceIndex(0)
enc = SentenceEncoder(dev, state_dict=synth.bert_state_dict(42, 0.02))
B, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4, 32)
ids, lens = synth.token_batch(B, S, seed=1, full=True)
d_ids, d_lens = torch.from_numpy(ids).cuda(), torch.from_numpy(lens).cuda()
for _ in range(3): enc.encode_ids_device(d_ids, d_lens)
tr = torch.zeros(128, dtype=torch.int64, device="cuda")
dev._ck(dev.lib.lrx_debug_set_trace(dev.h, C.c_void_p(tr.data_ptr())))
enc.encode_ids_device(d_ids, d_lens)
torch.cuda.synchronize()
dev._ck(dev.lib.lrx_debug_set_trace(dev.h, C.c_void_p(0)))
t = tr.cpu().numpy()
t = t[t > 0]
names = ["qkv", "bar1", "attn", "bar2", "out", "bar3", "ln1+up", "bar4", "down", "bar5", "ln2"]
print("embed", int(t[1] - t[0]), "cycles; total", int(t[-1] - t[0]))
for l in range(6):
    base = 1 + 11 * l
    d = [int(t[base + i + 1] - t[base + i]) for i in range(11)]
    print("layer", l, " ".join(f"{n}={v}" for n, v in zip(names, d)), "sum", sum(d))
dev.close()
