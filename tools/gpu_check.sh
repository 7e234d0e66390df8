#!/bin/bash
# one GPU session: parity tests, per-phase times in the merge modes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for mode in wave seq; do
  echo "== LIMGCU_MERGE_MODE=$mode"
  LIMGCU_MERGE_MODE=$mode timeout 600 python tools/phase_times.py 2>&1 | tee gpurun_out/phase_$mode.log
done
