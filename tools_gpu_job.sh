set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1l_launches.csv python bench.py --steps 1 --warmup 3 --queries 1000000 --no-cpu-baseline > gpurun_out/r1l_launch.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r1l_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-40:]: print(r[4][:60], r[7], r[8], float(r[-1])/1e6, "ms")
PY
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1l_bench.json 2> gpurun_out/r1l_bench.err; cut -c1-400 gpurun_out/r1l_bench.json; grep -o '"kernel_ms_per_step[^,]*,[^,]*' gpurun_out/r1l_bench.json; tail -5 gpurun_out/r1l_bench.err
timeout 600 python -m pytest tests -m gpu -x -q -k "find_path" 2>&1 | tail -3
