set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1m_pytest.log; cat gpurun_out/r1m_pytest.log
timeout 900 python tools/sweep_fp.py 1000000 > gpurun_out/r1m_sweep.log 2>&1; cat gpurun_out/r1m_sweep.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r1m_launches.csv python bench.py --steps 1 --warmup 3 --queries 1000000 --no-cpu-baseline > gpurun_out/r1m_launch.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r1m_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-26:]: print(r[4][:60], r[7], r[8], float(r[-1])/1e6, "ms")
PY
