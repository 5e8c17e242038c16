"""Prints a bench_ops.py JSON as a table: python scripts/show_ops.py file.json [filter]"""
import json, sys
d = json.load(open(sys.argv[1]))
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for r in d:
    line = (f"{r['op']:8s} {str(r['shape']):16s} {r['dtype']} k={r['batch']:9d} {r['ms_median']:8.3f} ms {r['Mmat_per_s']:10.2f} Mmat/s "
            f"hbm {r['frac_hbm_measured']*100:5.1f}% {r['GFLOPs']:8.0f} GF ref {r.get('ref_ms', 0):8.3f} ms x{r.get('speedup_vs_ref', 0):.2f}")
    if flt in line: print(line)
