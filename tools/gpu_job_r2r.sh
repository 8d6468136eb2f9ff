#!/bin/bash
# code-size variants (instruction cache): rolled replay / hoisted policies / rolled visit, on the table kernel
# (60, 61, 65) and on the node-directory kernel (62, 63, 64, 66, 67)
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
echo "== variants (200k, bit-exactness)"; VARIANT_CFGS=0,60,61,65,62,63,66,67,66g40 timeout 400 python tools/variant_check.py 2>&1 | grep "^C4\|rror" | tee $out/r2r_variants.log
run() { r=$(env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --queries 1000000 --no-cpu-baseline --no-strong 2>$out/r2r_err.log | tail -1)
  echo "$*: $(echo "$r" | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("value %.0f q/s path_ms %.2f snap_ms %.2f" % (j["value"], j["roofline"]["kernel_ms_per_step"], j["roofline"]["snap_ms_per_step"]))')"; }
{ for c in 0 60 61 65 62 63 64 66 67; do run HBN_LANE_CFG=$c; done; } 2>&1 | tee $out/r2r_sweep.log
for c in 65 66; do
echo "== metrics cfg $c"
HBN_LANE_CFG=$c timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_astar_lane -s 6 -c 1 --csv --log-file $out/r2r_ncu_cfg$c.csv \
  python bench.py --steps 1 --warmup 3 --queries 1000000 --no-cpu-baseline --no-strong > /dev/null 2>&1
python - $c <<'PY'
import csv,sys
c=sys.argv[1]
for r in csv.reader(open(f"gpurun_out/r2r_ncu_cfg{c}.csv")):
    if len(r)>10 and r[0]!="ID": print(" ", r[4][:40], r[-3], r[-2], r[-1])
PY
done 2>&1 | tee $out/r2r_metrics.log
