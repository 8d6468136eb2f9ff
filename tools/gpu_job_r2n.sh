#!/bin/bash
# Round 2, profile of the SHIPPED configuration: GPU tests, ncu --set full of k_astar_lane and of the other
# query kernels, launch list of the bench command, racecheck, bench lines.
tag=r2n
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $out/${tag}_pytest.log
echo "== bench"; timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; cut -c1-300 $out/${tag}_bench.json
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; cut -c1-300 $out/${tag}_bench_ref.json
echo "== ncu full k_astar_lane"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_astar_lane -s 3 -c 1 -o $out/${tag}_astar_lane -f \
  python bench.py --steps 1 --warmup 3 --queries 1000000 --no-cpu-baseline --no-strong > $out/${tag}_ncu.log 2>&1
ncu -i $out/${tag}_astar_lane.ncu-rep --page raw --csv > $out/${tag}_astar_lane_raw.csv 2>/dev/null
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 3 --queries 1000000 --no-cpu-baseline --no-strong > $out/${tag}_launch.log 2>&1
echo "== ncu full: wall, trystep, random, snap pipeline"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_wall -s 2 -c 1 -o $out/${tag}_wall -f python bench.py --config c5wall --queries 2000000 --steps 1 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_random -s 2 -c 1 -o $out/${tag}_random -f python bench.py --config c5rand --queries 2000000 --steps 1 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_trystep_a -s 2 -c 1 -o $out/${tag}_trystep -f python tools/c2_step.py 262144 2 separate > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_snap_ -s 20 -c 5 -o $out/${tag}_snap -f python bench.py --config c4snap --steps 1 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fp_funnel -s 3 -c 1 -o $out/${tag}_funnel -f python bench.py --steps 1 --warmup 3 --queries 1000000 --no-cpu-baseline --no-strong > /dev/null 2>&1
for k in wall random trystep snap funnel; do ncu -i $out/${tag}_$k.ncu-rep --page raw --csv > $out/${tag}_${k}_raw.csv 2>/dev/null; done
echo "== racecheck"
timeout 600 compute-sanitizer --tool racecheck python tools/small_fp.py c4_building 20000 > $out/${tag}_racecheck.log 2>&1
tail -4 $out/${tag}_racecheck.log
ls -la $out/${tag}_*
