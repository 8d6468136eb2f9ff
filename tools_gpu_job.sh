set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "search_variants and lane" 2>&1 | tail -15 > gpurun_out/r1i_pytest_lane.log; cat gpurun_out/r1i_pytest_lane.log
timeout 900 python tools/sweep_fp.py 1000000 > gpurun_out/r1i_sweep.log 2>&1; cat gpurun_out/r1i_sweep.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_astar_lane -s 3 -c 1 -o gpurun_out/r1i_astar_lane -f python bench.py --steps 1 --warmup 3 --queries 200000 --no-cpu-baseline > gpurun_out/r1i_ncu.log 2>&1
