#!/bin/bash
# Round 2: L2-residency variants of k_astar_lane (cfg 30-34) against the shipped one; new boundary tests.
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
echo "== variants (200k, bit-exactness)"; VARIANT_CFGS=0,30,31,32,33,34 timeout 240 python tools/variant_check.py 2>&1 | grep "^C4" | tee $out/r2e_variants.log
echo "== 1M"
for cfg in 0 30 31 32 33 34; do
  r=$(HBN_LANE_CFG=$cfg timeout 300 python bench.py --steps 2 --warmup 3 --queries 1000000 --no-cpu-baseline 2>/dev/null | tail -1)
  echo "cfg $cfg: $(echo "$r" | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print("value %.0f q/s path_ms %.2f snap_ms %.2f" % (j["value"], j["roofline"]["kernel_ms_per_step"], j["roofline"]["snap_ms_per_step"]))')"
done 2>&1 | tee $out/r2e_sweep.log
echo "== dram bytes per variant (ncu, 1M)"
for cfg in 0 31 32; do
  HBN_LANE_CFG=$cfg timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum --clock-control none -k regex:k_astar_lane -s 3 -c 1 --csv python bench.py --steps 1 --warmup 3 --queries 1000000 --no-cpu-baseline 2>/dev/null | grep -E "dram__|gpu__time|lts__" | cut -d, -f5,13-15 | sed "s/^/cfg $cfg: /"
done 2>&1 | tee $out/r2e_dram.log
echo "== new gpu tests"; timeout 900 python -m pytest tests/test_gpu_boundary.py -x -q 2>&1 | tail -15 | tee $out/r2e_pytest.log
