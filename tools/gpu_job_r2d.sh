#!/bin/bash
# Round 2: is k_astar_lane bound by the number of searches in flight (L2 footprint) or by latency?
# Sweep resident one-warp blocks per SM for the shipped kernel (cfg 0, 63 shared heap entries) and cfg 23 (95 entries).
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
for cfg in 0 23; do
  for b in 16 12 10 8 6 4 3; do
    if [ $cfg = 23 ] && [ $b -gt 11 ]; then continue; fi
    r=$(HBN_LANE_CFG=$cfg HBN_FP_BLOCKS_PER_SM=$b timeout 300 python bench.py --steps 2 --warmup 3 --queries 1000000 --no-cpu-baseline 2>/dev/null | tail -1)
    echo "cfg $cfg blocks/SM $b: $(echo "$r" | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("value %.0f q/s path_ms %.2f snap_ms %.2f" % (j["value"], j["roofline"]["kernel_ms_per_step"], j["roofline"]["snap_ms_per_step"]))')"
  done
done 2>&1 | tee $out/r2d_blocks_sweep.log
