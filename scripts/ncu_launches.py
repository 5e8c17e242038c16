import csv, sys, collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
seq=[(r[ki], float(r[vi].replace(",",""))) for r in rows[1:]]
tail=int(sys.argv[2]) if len(sys.argv)>2 else 40
for k,v in seq[-tail:]:
    import re
    k=re.sub(r"\(.*","",k)[:70]
    print(f"{v/1e3:10.1f} us  {k}")
