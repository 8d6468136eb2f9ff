#!/bin/bash
# C2 (1024 envs): wall time per dependent step through the separate / fused entry points, launch list of a step
out=gpurun_out
tag=${1:-rX}
export HBN_QUERY_CACHE=/tmp/hbn_queries
echo "== C2 step"; for m in separate fused; do timeout 120 python tools/c2_step.py 1024 50 $m 2>&1 | tail -1; done | tee $out/${tag}_c2.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_c2_launches.csv python tools/c2_step.py 1024 4 fused > /dev/null 2>&1
python - $tag <<'PY' | tee -a gpurun_out/${tag}_c2.log
import csv, collections, sys
rows=[r for r in csv.reader(open(f"gpurun_out/{sys.argv[1]}_c2_launches.csv")) if len(r)>10]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
rows=rows[1:]
agg=collections.OrderedDict()
for r in rows[-36:]:
    k=r[ik].split("(")[0][:44]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[iv].replace(",",""))
for k,(n,t) in agg.items(): print(f"  {k:46s} {n/4:.1f} per step  {t/4/1000:.1f} us per step")
print("  total per step (serialised, cold):", sum(t for n,t in agg.values())/4/1000, "us")
PY
echo "== bench c2 / c4"; for c in c2 c4; do timeout 600 python bench.py --config $c --no-cpu-baseline 2>/dev/null | python -c 'import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j["metric"], "value %.4g e2e %.4g ms/step %.3f parity %s" % (j["value"], j["e2e"]["value"], j["ms_per_step"], j.get("parity_sample")))'; done
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
