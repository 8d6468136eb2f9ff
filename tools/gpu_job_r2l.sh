#!/bin/bash
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
echo "== variants (200k, bit-exactness)"; VARIANT_CFGS=0,38,39 timeout 240 python tools/variant_check.py 2>&1 | grep "^C4" | tee $out/r2l_variants.log
run() { r=$(env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --queries 1000000 --no-cpu-baseline --no-strong 2>/dev/null | tail -1)
  echo "$*: $(echo "$r" | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("value %.0f q/s path_ms %.2f snap_ms %.2f" % (j["value"], j["roofline"]["kernel_ms_per_step"], j["roofline"]["snap_ms_per_step"]))')"; }
{ run HBN_LANE_CFG=0; run HBN_LANE_CFG=38; run HBN_LANE_CFG=39; run HBN_LANE_CFG=0 HBN_FP_BLOCKS_PER_SM=14; run HBN_LANE_CFG=0 HBN_FP_BLOCKS_PER_SM=12; } 2>&1 | tee $out/r2l_sweep.log
echo "== new gpu tests"; timeout 900 python -m pytest tests/test_greedy_follower.py tests/test_gpu_boundary.py -m gpu -x -q -k "follower or dispatcher or Follower" 2>&1 | tail -8 | tee $out/r2l_pytest.log
echo "== c5wall"; timeout 600 python bench.py --config c5wall --steps 3 2>$out/r2l_c5wall.err | tail -1 | cut -c1-400
