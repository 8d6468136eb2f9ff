set -x
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1c_pytest_gpu.log; cat gpurun_out/r1c_pytest_gpu.log
timeout 300 python bench.py --steps 3 > gpurun_out/r1c_bench.json 2> gpurun_out/r1c_bench.err; tail -3 gpurun_out/r1c_bench.err; cat gpurun_out/r1c_bench.json
timeout 200 compute-sanitizer --tool racecheck python tools/small_fp.py t_building 1000 2>&1 | grep -v "^=========     at\|^=========     by\|Host Frame" | tail -12
