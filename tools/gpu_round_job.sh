#!/bin/bash
# One gpurun call that re-measures everything a round's profile summary quotes (1 GPU, about 8 minutes):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_round_job.sh r2a'
# Outputs land in gpurun_out/<tag>_*; copy what is to be judged into profiles/.
# tools/variant_check.py needs tools/_q/variant_queries.npz (python tools/variant_check.py --make, on the CPU side).
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
echo "== variants" ; timeout 120 python tools/variant_check.py 2>&1 | tee $out/${tag}_variants.log | tail -30
echo "== gpu tests" ; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $out/${tag}_pytest.log
echo "== bench" ; timeout 400 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err ; tail -c 600 $out/${tag}_bench.json
echo "== 1M sweep of the kernel candidates" ; timeout 600 python tools/sweep_fp.py 1000000 2>&1 | tee $out/${tag}_sweep.log | tail -20
echo "== other configs" ; timeout 600 python tools/bench_configs.py > $out/${tag}_configs.json 2> $out/${tag}_configs.err ; cut -c1-300 $out/${tag}_configs.json
echo "== launch list" ; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --queries 1000000 --no-cpu-baseline > $out/${tag}_launch.log 2>&1
