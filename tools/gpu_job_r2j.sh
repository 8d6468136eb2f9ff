#!/bin/bash
# Round 2: bitmap variant (cfg 37), record-store-normal (cfg 36) vs shipped V10; then the full GPU test suite.
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
echo "== variants (200k, bit-exactness)"; VARIANT_CFGS=0,1,36,37 timeout 240 python tools/variant_check.py 2>&1 | grep "^C4" | tee $out/r2j_variants.log
for cfg in 0 36 37; do
  r=$(HBN_LANE_CFG=$cfg timeout 300 python bench.py --steps 2 --warmup 3 --queries 1000000 --no-cpu-baseline --no-strong 2>/dev/null | tail -1)
  echo "cfg $cfg: $(echo "$r" | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("value %.0f q/s path_ms %.2f snap_ms %.2f" % (j["value"], j["roofline"]["kernel_ms_per_step"], j["roofline"]["snap_ms_per_step"]))')"
done 2>&1 | tee $out/r2j_sweep.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum
for cfg in 37; do
  HBN_LANE_CFG=$cfg timeout 600 ncu --metrics $M --clock-control none -k regex:k_astar_lane -s 3 -c 1 --csv --log-file $out/r2j_ncu_cfg$cfg.csv python bench.py --steps 1 --warmup 3 --queries 1000000 --no-cpu-baseline --no-strong > /dev/null 2>&1
  echo "== cfg $cfg (per expansion)"; python tools/ncu_csv.py $out/r2j_ncu_cfg$cfg.csv 863971677 | cut -c30-
done 2>&1 | tee $out/r2j_dram.log
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $out/r2j_pytest.log
