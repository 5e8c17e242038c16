"""Top stall sites of a kernel in an .ncu-rep (source page, SASS view): python scripts/ncu_top.py rep kernel-regex [N]
Prints the N SASS instructions with the most warp-stall samples, their dominant stall reasons, and phase totals."""
import csv, subprocess, sys, io, collections
rep, pat = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
# several kernels may be concatenated: split on the "Kernel Name" rows
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}; blocks.append(cur); continue
    if cur is not None: cur["rows"].append(row)
b = blocks[which]
hdr, rows = b["rows"][0], b["rows"][1:]
print(b["name"][:150], "| kernels in report:", len(blocks))
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ci["# Samples"]]) for r in rows)
agg = collections.Counter()
for r in rows:
    for s in stall_cols: agg[s] += int(r[ci[s]] or 0)
print("total samples", tot, "instructions", len(rows), "executed warp-inst", sum(int(r[ci["Instructions Executed"]]) for r in rows))
print("stall totals:", ", ".join(f"{k[6:]}={v*100/tot:.1f}%" for k, v in agg.most_common(8)))
order = sorted(range(len(rows)), key=lambda i: -int(rows[i][ci["# Samples"]]))[:N]
for i in sorted(order):
    r = rows[i]
    st = sorted(((int(r[ci[s]] or 0), s[6:]) for s in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {int(r[ci['# Samples']])*100/tot:5.1f}%  exec={r[ci['Instructions Executed']]:>9s}  {r[ci['Source']].strip()[:70]:70s} " + " ".join(f"{n}:{c}" for c, n in st if c))
if len(sys.argv) > 5:
    # cumulative sample share per block of K instructions
    K = int(sys.argv[5]); acc = collections.Counter(); st = collections.defaultdict(collections.Counter)
    for i, r in enumerate(rows):
        acc[i // K] += int(r[ci["# Samples"]])
        for s in stall_cols: st[i // K][s[6:]] += int(r[ci[s]] or 0)
    for b_ in sorted(acc):
        print(f"[{b_*K:5d}-{b_*K+K-1:5d}] {acc[b_]*100/tot:5.1f}%  " + " ".join(f"{n}:{c*100//max(acc[b_],1)}%" for n, c in st[b_].most_common(4)))
