#!/usr/bin/env python
"""Timeline of CTA 0 of one tensor-core GEMM launch (clock64 stamps via lrx_debug_set_trace)."""
import sys, ctypes as C
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from legal_rag_engine_b200.device_index import DeviceIndex
dev = DeviceIndex(0)
M = 148 * 128
tr = torch.zeros(128, dtype=torch.int64, device="cuda")
for (N, K, epi, name) in ((1152, 384, 0, "QKV"), (1536, 384, 1, "FFN1"), (384, 384, 2, "Wo+LN"), (384, 1536, 2, "FFN2+LN")):
    a = torch.randn(M, K, device="cuda").half(); w = (torch.randn(N, K, device="cuda") * 0.05).half()
    bias = torch.randn(N, device="cuda"); res = torch.randn(M, 384, device="cuda").half()
    g = torch.ones(384, device="cuda"); b = torch.zeros(384, device="cuda")
    kw = dict(bias=bias, residual=res, gamma=g, beta=b) if epi == 2 else dict(bias=bias)
    for _ in range(3):
        dev.gemm_f16(a, w, epi=epi, **kw)
    torch.cuda.synchronize()
    dev._ck(dev.lib.lrx_debug_set_trace(dev.h, C.c_void_p(tr.data_ptr())))
    tr.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dev.gemm_f16(a, w, epi=epi, **kw); e1.record(); torch.cuda.synchronize()
    dev._ck(dev.lib.lrx_debug_set_trace(dev.h, C.c_void_p(0)))
    t = tr.cpu().numpy(); t0 = t[0]
    us = lambda x: (x - t0) / 1965.0 if x else -1
    print(f"== {name}: call {e0.elapsed_time(e1)*1e3:.1f} us (incl. tensor-map build); setup done {us(t[1]):.2f}, loops done {us(t[2]):.2f}, exit {us(t[3]):.2f}")
    for u in range(8):
        m = t[16 + 4 * u: 20 + 4 * u]; e = t[64 + 4 * u: 68 + 4 * u]
        if m[3] == 0: break
        print(f"  tile {u}: MMA wait_tempty {us(m[0]):6.2f}->{us(m[1]):6.2f} first_full {us(m[2]):6.2f} commit {us(m[3]):6.2f} | EPI wait {us(e[0]):6.2f}->{us(e[1]):6.2f} done {us(e[2]):6.2f}")
dev.close()
