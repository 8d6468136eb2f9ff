#!/bin/bash
# Round 2, first GPU call: re-profile the shipped k_astar_lane at HEAD, time the queued variants, racecheck.
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== variants (200k)"; VARIANT_CFGS=0,0k0,24,17,22,20,21,18,19 timeout 240 python tools/variant_check.py 2>&1 | tee $out/${tag}_variants.log | tail -20
echo "== 1M sweep"; timeout 700 python tools/sweep_fp.py 1000000 2>&1 | tee $out/${tag}_sweep.log | tail -12
echo "== ncu full k_astar_lane"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_astar_lane -s 3 -c 1 -o $out/${tag}_astar_lane -f \
  python bench.py --steps 1 --warmup 3 --queries 1000000 --no-cpu-baseline > $out/${tag}_ncu.log 2>&1
tail -3 $out/${tag}_ncu.log | cut -c1-300
ncu -i $out/${tag}_astar_lane.ncu-rep --page raw --csv > $out/${tag}_astar_lane_raw.csv 2>/dev/null
ncu -i $out/${tag}_astar_lane.ncu-rep --page source --csv > $out/${tag}_astar_lane_src.csv 2>/dev/null
echo "== racecheck"
timeout 600 compute-sanitizer --tool racecheck python tools/small_fp.py c4_building 20000 > $out/${tag}_racecheck.log 2>&1
tail -8 $out/${tag}_racecheck.log
