#!/bin/bash
# Round 2: full GPU test suite, then every bench config once.
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $out/r2k_pytest.log
for c in c4 c2 c3 c4snap c5wall c5rand; do
  echo "== bench --config $c"
  timeout 900 python bench.py --config $c --steps 3 > $out/r2k_bench_$c.json 2> $out/r2k_bench_$c.err
  python - <<PY
import json
try:
    j = json.loads(open("$out/r2k_bench_$c.json").read().strip().splitlines()[-1])
    cb = j.get("cpu_baseline") or {}
    print("  %s: value %.4g %s, e2e %.4g, ms/step %.3f, launches %d, cpu %.4g (%s cores; 1 thread %.4g), parity %s, strong %s" % (
        j["metric"], j["value"], j["unit"], j["e2e"]["value"], j["ms_per_step"], j["gpu_launches"], cb.get("value", 0), cb.get("cores"),
        cb.get("one_thread_value", 0), j.get("parity_sample"), (j.get("strong_scaling") or {}).get("value")))
except Exception as ex:
    print("  failed:", ex)
    print(open("$out/r2k_bench_$c.err").read()[-1500:])
PY
done 2>&1 | tee $out/r2k_configs.log
