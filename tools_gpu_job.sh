set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1n_pytest.log; cat gpurun_out/r1n_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r1n_bench.json 2> gpurun_out/r1n_bench.err; cut -c1-300 gpurun_out/r1n_bench.json; grep -o '"e2e[^}]*}' gpurun_out/r1n_bench.json; grep -o '"kernel_ms_per_step[^,]*,[^,]*' gpurun_out/r1n_bench.json; grep -o '"cpu_baseline.*' gpurun_out/r1n_bench.json | cut -c1-200;  tail -5 gpurun_out/r1n_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r1n_launches.csv python bench.py --steps 1 --warmup 3 --queries 1000000 --no-cpu-baseline > gpurun_out/r1n_launch.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r1n_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-22:]: print(r[4][:60], r[7], r[8], float(r[-1])/1e6, "ms")
PY
