#!/bin/bash
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
echo "== variants (200k)"; VARIANT_CFGS=0,61,68,70 timeout 400 python tools/variant_check.py 2>&1 | grep "^C4\|rror" | tee $out/r2s_variants.log
run() { r=$(env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --queries 1000000 --no-cpu-baseline --no-strong 2>$out/r2s_err.log | tail -1)
  echo "$*: $(echo "$r" | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("value %.0f q/s path_ms %.2f snap_ms %.2f" % (j["value"], j["roofline"]["kernel_ms_per_step"], j["roofline"]["snap_ms_per_step"]))')"; }
{ for c in 61 68 70 71 72; do run HBN_LANE_CFG=$c; done; } 2>&1 | tee $out/r2s_sweep.log
