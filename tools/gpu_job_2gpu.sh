#!/bin/bash
# two GPUs of one box: the driver's launch of bench.py (weak line + strong sub-measurement), the reference arm
# under torchrun, the multi-GPU GPU tests
out=gpurun_out
export HBN_QUERY_CACHE=/tmp/hbn_queries
nvidia-smi --query-gpu=name --format=csv,noheader
echo "== gpu tests (multi-device)"; timeout 900 python -m pytest tests -m gpu -x -q -k "multi or shard or device" 2>&1 | tail -4 | tee $out/r2x_pytest.log
echo "== bench --gpus 2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $out/r2x_bench_2gpu.json 2> $out/r2x_bench_2gpu.err; tail -1 $out/r2x_bench_2gpu.json | cut -c1-300; tail -3 $out/r2x_bench_2gpu.err
echo "== bench --impl reference --gpus 2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $out/r2x_bench_ref_2gpu.json 2> $out/r2x_bench_ref_2gpu.err; tail -1 $out/r2x_bench_ref_2gpu.json | cut -c1-300
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2x_bench_2gpu.json").read().strip().splitlines()[-1])
print("weak:", j["value"], "e2e", j["e2e"]["value"], "n_gpus", j["n_gpus"]); print("strong:", j.get("strong_scaling"))
PY
