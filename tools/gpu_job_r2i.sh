#!/bin/bash
# Round 2: full GPU test suite; per-kind L2 hit/miss of the search kernel (diagnostic build); bench line.
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $out/r2i_pytest.log
echo "== bench"; timeout 600 python bench.py > $out/r2i_bench.json 2> $out/r2i_bench.err; cut -c1-1200 $out/r2i_bench.json; tail -3 $out/r2i_bench.err
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for c in first last normal; do for op in read write; do for hm in hit miss; do M=$M,lts__t_sectors_srcunit_tex_op_${op}_evict_${c}_lookup_${hm}.sum; done; done; done
for cfg in 44 40 41 42 43; do
  HBN_LIBRARY=$PWD/habitat-sim_b200/lib_diag/libhbn.so HBN_LANE_CFG=$cfg timeout 600 ncu --metrics $M --clock-control none -k regex:k_astar_lane -s 3 -c 1 --csv --log-file $out/r2i_ncu_cfg$cfg.csv python bench.py --steps 1 --warmup 3 --queries 1000000 --no-cpu-baseline --no-strong > /dev/null 2>&1
  echo "== cfg $cfg (per expansion)"; python tools/ncu_csv.py $out/r2i_ncu_cfg$cfg.csv 863971677 | cut -c30- | grep -v " 0 sector"
done 2>&1 | tee $out/r2i_kinds.log
