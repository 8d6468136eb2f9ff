#!/bin/bash
# Round 2: L2 fetch granularity experiment on the new default kernel (V = 10), then the full GPU test suite.
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
for fetch in "" 32 128; do
  r=$(HBN_L2_FETCH=$fetch timeout 300 python bench.py --steps 2 --warmup 3 --queries 1000000 --no-cpu-baseline 2>$out/r2g_err.log | tail -1)
  echo "L2 fetch '$fetch': $(echo "$r" | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("value %.0f q/s e2e %.0f path_ms %.2f snap_ms %.2f" % (j["value"], j["e2e"]["value"], j["roofline"]["kernel_ms_per_step"], j["roofline"]["snap_ms_per_step"]))')"; grep "L2 fetch" $out/r2g_err.log | head -1
done 2>&1 | tee $out/r2g_fetch.log
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/r2g_pytest.log
