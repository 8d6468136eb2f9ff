set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r1d_pytest_gpu.log; cat gpurun_out/r1d_pytest_gpu.log
timeout 300 python bench.py --steps 3 > gpurun_out/r1d_bench.json 2> gpurun_out/r1d_bench.err; tail -3 gpurun_out/r1d_bench.err; cat gpurun_out/r1d_bench.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1d_launches.csv python bench.py --steps 2 --warmup 3 --queries 200000 --no-cpu-baseline > gpurun_out/r1d_ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_findpath_w -s 6 -c 1 -o gpurun_out/r1d_findpath -f python bench.py --steps 1 --warmup 3 --queries 100000 --no-cpu-baseline > gpurun_out/r1d_ncu_full.log 2>&1
tail -2 gpurun_out/r1d_ncu_full.log | cut -c1-300
