#!/bin/bash
# Round 2: heap evict_last (cfg 35) + L2 hit/miss by eviction class (= by data structure) under ncu.
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
for cfg in 0 32 35; do
  r=$(HBN_LANE_CFG=$cfg timeout 300 python bench.py --steps 2 --warmup 3 --queries 1000000 --no-cpu-baseline 2>/dev/null | tail -1)
  echo "cfg $cfg: $(echo "$r" | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("value %.0f q/s path_ms %.2f snap_ms %.2f" % (j["value"], j["roofline"]["kernel_ms_per_step"], j["roofline"]["snap_ms_per_step"]))')"
done 2>&1 | tee $out/r2f_sweep.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for c in first last normal normal_demote; do for op in read write; do for hm in hit miss; do M=$M,lts__t_sectors_srcunit_tex_op_${op}_evict_${c}_lookup_${hm}.sum; done; done; done
M=$M,lts__t_sectors_srcunit_ltcfabric_lookup_hit.sum,lts__t_sectors_srcunit_ltcfabric_lookup_miss.sum,smsp__inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum
for cfg in 0 35; do
  HBN_LANE_CFG=$cfg timeout 600 ncu --metrics $M --clock-control none -k regex:k_astar_lane -s 3 -c 1 --csv --log-file $out/r2f_ncu_cfg$cfg.csv python bench.py --steps 1 --warmup 3 --queries 1000000 --no-cpu-baseline > /dev/null 2>&1
  echo "== cfg $cfg (per expansion)"; python tools/ncu_csv.py $out/r2f_ncu_cfg$cfg.csv 863971677 | cut -c30-
done 2>&1 | tee $out/r2f_classes.log
echo "== new gpu tests"; timeout 900 python -m pytest tests/test_gpu_boundary.py -x -q 2>&1 | tail -15 | tee $out/r2f_pytest.log
