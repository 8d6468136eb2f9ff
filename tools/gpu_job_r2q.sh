#!/bin/bash
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
run() { r=$(env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --queries 1000000 --no-cpu-baseline --no-strong 2>$out/r2q_err.log | tail -1)
  echo "$*: $(echo "$r" | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("value %.0f q/s path_ms %.2f snap_ms %.2f" % (j["value"], j["roofline"]["kernel_ms_per_step"], j["roofline"]["snap_ms_per_step"]))')"; }
{ run HBN_LANE_CFG=51; } 2>&1 | tee $out/r2q_sweep.log
echo "== ncu full cfg 50"
HBN_LANE_CFG=50 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_astar_lane -s 6 -c 1 -o $out/r2q_astar_lane50 -f \
  python bench.py --steps 1 --warmup 3 --queries 1000000 --no-cpu-baseline --no-strong > $out/r2q_ncu.log 2>&1
ncu -i $out/r2q_astar_lane50.ncu-rep --page raw --csv > $out/r2q_astar_lane50_raw.csv 2>/dev/null
ncu -i $out/r2q_astar_lane50.ncu-rep --page source --csv --print-source=cuda,sass > $out/r2q_astar_lane50_src.csv 2>/dev/null
rm -f $out/r2q_astar_lane50.ncu-rep
python tools/ncu_lines.py $out/r2q_astar_lane50_src.csv 40
