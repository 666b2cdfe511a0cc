#!/usr/bin/env python
"""Per-SASS-instruction view of an ncu report: python tools/ncu_src.py rep.ncu-rep [top]
Prints totals, the opcode mix, stall reasons, and the instructions with the most stall samples."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout.splitlines()
r = csv.reader(out); next(r); hdr = next(r)
ix = {h: i for i, h in enumerate(hdr)}
rows = [x for x in r if len(x) == len(hdr)]
I = lambda x, k: int(x[ix[k]] or 0)
tot = sum(I(x, "Instructions Executed") for x in rows); samp = sum(I(x, "# Samples") for x in rows)
print("warp instructions", tot, "samples", samp, "sass lines", len(rows))
ops = collections.Counter(); ss = collections.Counter()
for x in rows:
    op = x[ix["Source"]].split(); o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
    ops[o] += I(x, "Instructions Executed"); ss[o] += I(x, "# Samples")
for o, c in ops.most_common(22):
    print(f"  {o:10s} {100*c/tot:5.1f}% of instr  {100*ss[o]/samp:5.1f}% of samples")
st = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
print("stalls:", {k[6:]: sum(I(x, k) for x in rows) for k in st if sum(I(x, k) for x in rows) > samp // 200})
print("--- hottest instructions (samples, executed, source, top stalls)")
for n, x in sorted(((I(x, "# Samples"), x) for x in rows), key=lambda t: -t[0])[:top]:
    why = sorted(((I(x, k), k[6:]) for k in st), reverse=True)[:2]
    print(f"{n:6d} {I(x,'Instructions Executed'):9d}  {x[ix['Source']].strip():60s} {why}")
