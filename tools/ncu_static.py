"""Static code size by source line (SASS instructions per line) and where the 'no instruction' (instruction
cache miss) stall samples fall, from an `ncu --page source --csv --print-source=cuda,sass` export.
usage: python tools/ncu_static.py export.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur = None; hdr = None; line = None
size = {}; noinst = {}; text = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if r[0] == "Function Name" or hdr is None: continue
    if r[0] != "":
        line = (cur, int(r[0])); text[line] = r[1].strip()[:80]; continue
    if line is None: continue
    size[line] = size.get(line, 0) + 1
    try: noinst[line] = noinst.get(line, 0) + float(r[hdr["stall_no_inst"]] or 0)
    except (ValueError, KeyError): pass
tot = sum(size.values()); tn = sum(noinst.values()) or 1
print(f"SASS instructions {tot} ({tot * 16 / 1024:.1f} KB); no_inst samples {tn:.0f}")
# by function region (coarse): group lines
for k, v in sorted(size.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{v:5d} instr {v * 16:6d} B  no_inst {noinst.get(k, 0) / tn * 100:4.1f}%  {k[0]}:{k[1]}  {text.get(k, '')}")
