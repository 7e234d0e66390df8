#!/bin/bash
for cfg in "8 16" "8 48" "16 16" "16 48"; do
  set -- $cfg
  echo "== margin $1 gap $2"
  LIMGCU_PLAN_ASYNC=0 LIMGCU_MERGE_MARGIN=$1 LIMGCU_MERGE_GAP=$2 timeout 600 python tools/phase_times.py 2>&1 | grep "merged\|total" | cut -c1-60,150-420 | grep -A1 "which" | grep -v "^--"
done
