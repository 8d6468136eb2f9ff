"""Print 'metric = value unit' lines from an `ncu --csv` (details page) log, optionally scaled per unit of work.
usage: python tools/ncu_csv.py log.csv [divisor]"""
import csv, sys
div = float(sys.argv[2]) if len(sys.argv) > 2 else None
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
i_name, i_unit, i_val, i_k = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("Kernel Name")
for r in rows[1:]:
    try:
        v = float(r[i_val].replace(",", ""))
    except ValueError:
        continue
    extra = f"   ({v / div:.3f} per unit)" if div else ""
    print(f"{r[i_k][:28]:28s} {r[i_name]:75s} {v:.6g} {r[i_unit]}{extra}")
