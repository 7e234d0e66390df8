#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
for spec in 8 0 16; do
  echo "== LIMGCU_MERGE_SPEC=$spec"
  LIMGCU_MERGE_SPEC=$spec timeout 600 python tools/phase_times.py 2>&1 | grep "merged\|total\|profile" | tee gpurun_out/phase_spec$spec.log
done
