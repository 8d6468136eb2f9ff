#!/bin/bash
# snap pipeline: GPU tests, bench lines of the snap-bound configs, launch list of the snap kernels
tag=${1:-rX}
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/${tag}_pytest.log
for c in c4snap c5wall c4; do echo "== bench --config $c"; timeout 900 python bench.py --config $c > $out/${tag}_bench_$c.json 2> $out/${tag}_bench_$c.err; cut -c1-200 $out/${tag}_bench_$c.json; tail -2 $out/${tag}_bench_$c.err; done
echo "== launch list (c4snap)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_snap_launches.csv \
  python bench.py --config c4snap --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - $tag <<'PY'
import csv, collections, sys
rows=[r for r in csv.reader(open(f"gpurun_out/{sys.argv[1]}_snap_launches.csv")) if len(r)>10]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ik].split("(")[0][:40]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[iv].replace(",",""))
for k,(n,t) in agg.items(): print(f"  {k:42s} {n:4d} launches  {t/1e6:.3f} ms total  {t/n/1000:.1f} us each")
PY
echo "== ncu full k_snap_walk"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_snap_walk -s 8 -c 1 -o $out/${tag}_snapwalk -f python bench.py --config c4snap --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu -i $out/${tag}_snapwalk.ncu-rep --page raw --csv > $out/${tag}_snap_walk_raw.csv 2>/dev/null
ncu -i $out/${tag}_snapwalk.ncu-rep --page source --csv --print-source=cuda,sass > $out/${tag}_snap_walk_src.csv 2>/dev/null
rm -f $out/${tag}_snapwalk.ncu-rep
python tools/ncu_summary.py $out/${tag}_snap_walk_raw.csv | grep -E "^==|time_duration|thread_inst_executed_per|issue_active|inst_executed.sum|warps_active|long_scoreboard|lg_throttle|registers"
python tools/ncu_lines.py $out/${tag}_snap_walk_src.csv 14
