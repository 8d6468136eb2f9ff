"""Aggregate an `ncu --page source --csv --print-source=cuda,sass` export by source line.
usage: python tools/ncu_lines.py export.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
files = {}
cur = None
hdr = None
agg = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        # two "Source" columns: first is cuda, second sass
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0] != "":  # a source line row (aggregated over its SASS)
        try:
            samples = float(r[hdr["# Samples"]] or 0)
            inst = float(r[hdr["Instructions Executed"]] or 0)
            thr = float(r[hdr["Thread Instructions Executed"]] or 0)
        except ValueError:
            continue
        agg.append((cur, int(r[0]), r[1].strip()[:90], samples, inst, thr))
ts = sum(a[3] for a in agg) or 1
ti = sum(a[4] for a in agg) or 1
print(f"total samples {ts:.0f} instructions {ti:.3g}")
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    print(f"{a[3]/ts*100:5.1f}% t {a[4]/ti*100:5.1f}% i  thr/inst {a[5]/max(a[4],1):5.1f}  {a[0]}:{a[1]}  {a[2]}")
