"""Prints the configs block of one or two bench lines side by side: python scripts/show_bench.py ours.json [reference.json]"""
import json, sys
a = json.loads(open(sys.argv[1]).read())
b = json.loads(open(sys.argv[2]).read()) if len(sys.argv) > 2 else None
for k in ("value", "ms_per_step"):
    print(k, a[k], b[k] if b else "")
for k in ("e2e", "e2e_dropin", "strong", "cpu_baseline"):
    print(k, {x: (round(y, 3) if isinstance(y, float) else y) for x, y in (a.get(k) or {}).items() if x in ("value", "ms_per_step", "cores", "kind")},
          {x: (round(y, 3) if isinstance(y, float) else y) for x, y in ((b or {}).get(k) or {}).items() if x in ("value", "ms_per_step")} if b else "")
rb = {c["name"]: c for c in (b.get("configs") or [])} if b else {}
for c in a.get("configs") or []:
    r = c["roofline"]
    o = rb.get(c["name"])
    print(f"{c['name']:46s} {c['ms']:9.4f} ms {c['matrices_per_s']/1e6:10.2f} M/s {c['gflops']:8.0f} GF {r['bound']:6s} {r['frac']:.3f}"
          + (f"   ref {o['ms']:10.4f} ms  x{o['ms']/c['ms']:.2f}" if o else ""))
