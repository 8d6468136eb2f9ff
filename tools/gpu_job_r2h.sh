#!/bin/bash
out=gpurun_out
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $out/r2h_pytest.log
