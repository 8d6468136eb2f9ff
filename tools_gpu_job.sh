set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_pytest_gpu.log; cat gpurun_out/r1_pytest_gpu.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_ref.json 2> gpurun_out/r1_bench_ref.err; cat gpurun_out/r1_bench_ref.json
python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; tail -3 gpurun_out/r1_bench.err; cat gpurun_out/r1_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --queries 200000 --no-cpu-baseline > gpurun_out/r1_ncu_launches.log 2>&1
tail -20 gpurun_out/r1_launches.csv
ncu --set full --clock-control none --import-source on -k regex:k_findpath -s 6 -c 2 -o gpurun_out/r1_findpath -f python bench.py --steps 1 --warmup 3 --queries 100000 --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1
tail -5 gpurun_out/r1_ncu_full.log
ls -la gpurun_out
