#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  echo "== LOOKAHEAD=$1 ASYNC=$2"
  LIMGCU_MERGE_LOOKAHEAD=$1 LIMGCU_PLAN_ASYNC=$2 timeout 600 python tools/phase_times.py 2>&1 | grep "total\|which\|look-ahead" | cut -c1-420
done
