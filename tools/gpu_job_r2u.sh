#!/bin/bash
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
echo "== gpu tests (wall distance)"; timeout 900 python -m pytest tests -m gpu -x -q -k "wall or obstacle or ordinary or config or follower" 2>&1 | tail -4 | tee $out/r2u_pytest.log
echo "== bench c5wall"; timeout 900 python bench.py --config c5wall > $out/r2u_bench_c5wall.json 2> $out/r2u_bench_c5wall.err; cut -c1-200 $out/r2u_bench_c5wall.json; tail -2 $out/r2u_bench_c5wall.err
echo "== ncu k_wall_lane"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_wall -s 3 -c 3 -o $out/r2u_wall -f python bench.py --config c5wall --queries 2000000 --steps 1 --no-cpu-baseline > /dev/null 2>&1
ncu -i $out/r2u_wall.ncu-rep --page raw --csv > $out/r2u_wall_raw.csv 2>/dev/null
rm -f $out/r2u_wall.ncu-rep
python tools/ncu_summary.py $out/r2u_wall_raw.csv | grep -E "^==|time_duration|dram__bytes|thread_inst_executed_per|issue_active|inst_executed.sum|warps_active|long_scoreboard|lts__throughput"
