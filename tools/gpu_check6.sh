#!/bin/bash
for cfg in "8 16" "8 64" "16 16"; do
  set -- $cfg
  echo "== margin $1 gap $2"
  for name in c4_4k_flatui c2_4k_photo; do
  LIMGCU_PLAN_ASYNC=0 LIMGCU_MERGE_MARGIN=$1 LIMGCU_MERGE_GAP=$2 timeout 600 python tools/fail_rate.py $name 24 2>&1 | tail -1
  done
done
