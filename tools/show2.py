#!/usr/bin/env python
"""Headline fields of bench JSON lines: python tools/show2.py file.json ..."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    r, b = d.get("roofline", {}), d.get("bm25_kernel", {})
    print(f"{f}: value {d.get('value'):.1f} e2e {d['e2e']['value']:.1f} ms/step {d['ms_per_step']:.4f} "
          f"two_users {d.get('two_users_per_step', {}).get('value')} parity {d.get('parity', {}).get('mismatches')} "
          f"clocks {d.get('clocks', {}).get('sm_mhz')} {d.get('clocks', {}).get('reasons')}")
    print(f"   roofline {r.get('kernel')} {r.get('ms_per_launch'):.4f} ms in step = {r.get('achieved'):.0f} GB/s frac {r.get('frac'):.3f}; "
          f"alone {r['alone']['ms_per_launch']:.4f} ms frac {r['alone']['frac']:.3f}; step_frac {r['concurrent']['step_frac']:.3f}; "
          f"pass_ms {r.get('pass_ms_per_step')}")
    print(f"   bm25 in step {b['in_step']['ms_per_launch']:.4f} ms alone {b['alone']['ms_per_launch']:.4f} ms frac alone {b['alone']['frac']:.3f}")
    st = d.get("stages")
    if st:
        for k, v in st.items():
            if isinstance(v, dict):
                print("   stage", k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("queries_per_s", "scan_ms", "frac", "call_ms", "kernel", "ms", "seq_per_s", "frac_call")})
    if "cpu_baseline" in d:
        print("   cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
