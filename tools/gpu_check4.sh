#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
for v in 1 0; do
  echo "== LIMGCU_PLAN_ASYNC=$v"
  LIMGCU_PLAN_ASYNC=$v timeout 600 python tools/phase_times.py 2>&1 | grep "merged\|total" | tee gpurun_out/phase_async$v.log
done
