import json, sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if 'error' in d: print(f, json.dumps(d)[:800]); continue
        d2=d.get("two_users_per_step") or {}; print(f.split('/')[-1], round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'timed', d['timed_steps'], 'e2e', round(d['e2e']['value'],1), 'parity', d['parity']['mismatches'] if d.get('parity') else None, 'scan', round(d['roofline']['ms_per_launch'],4), round(d['roofline']['frac'],3), 'bm25', round(d['bm25_kernel']['in_step']['ms_per_launch'],4), 'host', round(d["host_enqueue_ms_per_step"],4), "two-users", round(d2.get("value",0),1))
    except Exception as e:
        print(f, 'FAILED', repr(e))
        try: print(open(f.replace('.json','.err')).read()[-1500:])
        except Exception: pass
