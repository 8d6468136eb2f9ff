#!/bin/bash
# N GPUs of one box, the driver's launch of bench.py: weak line + strong sub-measurement (1 M queries in all)
N=${1:-8}
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
nvidia-smi --query-gpu=name --format=csv,noheader | head -$N | tr '\n' ';'; echo
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > $out/r3_bench_${N}gpu.json 2> $out/r3_bench_${N}gpu.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
j=json.loads(open(f"gpurun_out/r3_bench_{N}gpu.json").read().strip().splitlines()[-1])
print("weak: value %.4g e2e %.4g n_gpus %d ms/step %.2f" % (j["value"], j["e2e"]["value"], j["n_gpus"], j["ms_per_step"])); print("strong:", j.get("strong_scaling")); print("cpu:", j["cpu_baseline"]["value"], j["cpu_baseline"]["cores"])
PY
tail -3 $out/r3_bench_${N}gpu.err
