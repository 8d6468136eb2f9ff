#!/bin/bash
# node-directory variant of k_astar_lane (cfg 50) against the shipped one: bit-exactness at 200 k (also with the
# overflow path forced), time at 1 M, DRAM bytes per launch, L2 set-aside experiment
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
echo "== variants (200k, bit-exactness)"; VARIANT_CFGS=0,50,50g40,50g4 timeout 300 python tools/variant_check.py 2>&1 | grep "^C4\|rror" | tee $out/r2p_variants.log
run() { r=$(env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --queries 1000000 --no-cpu-baseline --no-strong 2>$out/r2p_err.log | tail -1)
  echo "$*: $(echo "$r" | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("value %.0f q/s path_ms %.2f snap_ms %.2f parity %s" % (j["value"], j["roofline"]["kernel_ms_per_step"], j["roofline"]["snap_ms_per_step"], j.get("parity_sample")))')"; grep hbn $out/r2p_err.log | head -2; }
{ run HBN_LANE_CFG=0; run HBN_LANE_CFG=50; run HBN_LANE_CFG=50 HBN_L2_PERSIST_MB=64; run HBN_LANE_CFG=0 HBN_L2_PERSIST_MB=64; } 2>&1 | tee $out/r2p_sweep.log
for c in 50; do
echo "== dram cfg $c"
HBN_LANE_CFG=$c timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_red.sum,lts__t_sector_hit_rate.pct,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:k_astar_lane -s 6 -c 2 --csv --log-file $out/r2p_ncu_cfg$c.csv \
  python bench.py --steps 1 --warmup 3 --queries 1000000 --no-cpu-baseline --no-strong > /dev/null 2>&1
python - $c <<'PY'
import csv,sys
c=sys.argv[1]
for r in csv.reader(open(f"gpurun_out/r2p_ncu_cfg{c}.csv")):
    if len(r)>10 and r[0]!="ID": print(" ", r[4][:40], r[-3], r[-2], r[-1])
PY
done 2>&1 | tee $out/r2p_dram.log
