#!/bin/bash
# One gpurun call that re-measures what a round's profile summary quotes for the SHIPPED configuration (1 GPU, ~10 min):
#   gpurun --timeout 3400 -- bash tools/gpu_round_job.sh r2t
# Outputs land in gpurun_out/<tag>_*; copy what is to be judged into profiles/.  GPU tests, bench lines of every config, ncu --set full
# of k_astar_lane (+ source page), launch list of the bench command, racecheck.
tag=${1:-rX}
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $out/${tag}_pytest.log
echo "== bench"; timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; cut -c1-300 $out/${tag}_bench.json
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; cut -c1-200 $out/${tag}_bench_ref.json
for c in c2 c3 c4snap c5wall c5rand; do echo "== bench --config $c"; timeout 900 python bench.py --config $c > $out/${tag}_bench_$c.json 2> $out/${tag}_bench_$c.err; cut -c1-400 $out/${tag}_bench_$c.json; tail -2 $out/${tag}_bench_$c.err; done
echo "== ncu full k_astar_lane"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_astar_lane -s 3 -c 1 -o $out/${tag}_astar_lane -f \
  python bench.py --steps 1 --warmup 3 --queries 1000000 --no-cpu-baseline --no-strong > $out/${tag}_ncu.log 2>&1
ncu -i $out/${tag}_astar_lane.ncu-rep --page raw --csv > $out/${tag}_astar_lane_raw.csv 2>/dev/null
ncu -i $out/${tag}_astar_lane.ncu-rep --page source --csv --print-source=cuda,sass > $out/${tag}_astar_lane_src.csv 2>/dev/null
rm -f $out/${tag}_astar_lane.ncu-rep
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 3 --queries 1000000 --no-cpu-baseline --no-strong > $out/${tag}_launch.log 2>&1
echo "== racecheck"
timeout 600 compute-sanitizer --tool racecheck python tools/small_fp.py c4_building 20000 > $out/${tag}_racecheck.log 2>&1
tail -4 $out/${tag}_racecheck.log
ls -la $out/${tag}_*
