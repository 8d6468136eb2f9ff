set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "search_variants and lane" 2>&1 | tail -15 > gpurun_out/r1g_pytest_lane.log; cat gpurun_out/r1g_pytest_lane.log
HBN_FP_G=lane timeout 600 compute-sanitizer --tool memcheck python tools/small_fp.py > gpurun_out/r1g_memcheck.log 2>&1; tail -8 gpurun_out/r1g_memcheck.log
timeout 900 python tools/sweep_fp.py 1000000 > gpurun_out/r1g_sweep.log 2>&1; cat gpurun_out/r1g_sweep.log
