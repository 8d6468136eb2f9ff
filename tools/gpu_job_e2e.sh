#!/bin/bash
# end-to-end (host buffers) numbers of the throughput configs + the GPU suite
out=gpurun_out
tag=${1:-rX}
export HBN_QUERY_CACHE=/tmp/hbn_queries
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/${tag}_pytest.log
for c in c4 c4snap c5wall c5rand c3; do timeout 900 python bench.py --config $c --no-cpu-baseline > $out/${tag}_bench_$c.json 2> $out/${tag}_bench_$c.err
python - $out/${tag}_bench_$c.json $c <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], "value %.4g  e2e %.4g  ms/step %.2f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]))
PY
done
