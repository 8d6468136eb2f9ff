#!/bin/bash
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
echo "== variants (200k, bit-exactness)"; VARIANT_CFGS=0,38,45,46,47,48 timeout 240 python tools/variant_check.py 2>&1 | grep "^C4" | tee $out/r2m_variants.log
run() { r=$(env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --queries 1000000 --no-cpu-baseline --no-strong 2>/dev/null | tail -1)
  echo "$*: $(echo "$r" | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("value %.0f q/s path_ms %.2f snap_ms %.2f" % (j["value"], j["roofline"]["kernel_ms_per_step"], j["roofline"]["snap_ms_per_step"]))')"; }
{ run HBN_LANE_CFG=38; run HBN_LANE_CFG=45; run HBN_LANE_CFG=46; run HBN_LANE_CFG=47; run HBN_LANE_CFG=48; } 2>&1 | tee $out/r2m_sweep.log
echo "== C2 step"; for m in separate fused graph; do timeout 120 python tools/c2_step.py 1024 50 $m 2>&1 | tail -1; done | tee $out/r2m_c2.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/r2m_c2_launches.csv python tools/c2_step.py 1024 4 fused > /dev/null 2>&1
python - <<'PY' | tee -a gpurun_out/r2m_c2.log
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2m_c2_launches.csv")) if len(r)>10]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
rows=rows[1:]
# the last 4 steps x kernels: take the last 36 launches (9 per step)
agg=collections.OrderedDict()
for r in rows[-36:]:
    k=r[ik].split("(")[0][:40]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[iv].replace(",",""))
for k,(n,t) in agg.items(): print(f"  {k:42s} {n/4:.1f} per step  {t/4/1000:.1f} us per step")
print("  total per step (serialised, cold):", sum(t for n,t in agg.values())/4/1000, "us")
PY
