#!/bin/bash
for mode in ${MODES:-wave}; do
  echo "== LIMGCU_MERGE_MODE=$mode"
  LIMGCU_MERGE_MODE=$mode timeout 600 python tools/phase_times.py 2>&1 | tee gpurun_out/phase_$mode.log
done
