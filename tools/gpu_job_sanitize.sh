#!/bin/bash
# compute-sanitizer over a small run of every batched entry point (memcheck, racecheck) + the new GPU test
out=gpurun_out
tag=${1:-rX}
echo "== new test"; timeout 600 python -m pytest tests -m gpu -x -q -k "tiers or snap or obstacle" 2>&1 | tail -3
echo "== plain"; timeout 300 python tools/small_all.py c4_building 8192 2>&1 | tail -6
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck python tools/small_all.py c4_building 8192 > $out/${tag}_memcheck.log 2>&1; tail -8 $out/${tag}_memcheck.log
echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck python tools/small_all.py c4_building 8192 > $out/${tag}_racecheck_all.log 2>&1; tail -8 $out/${tag}_racecheck_all.log
