#!/bin/bash
for a in 0 1 2; do for g in 1 2 3 4; do
  if [ $a = 0 ] && [ $g != 2 ]; then continue; fi
  echo -n "PLAN_ASYNC=$a PLAN_CTAS=$g: "; LIMGCU_PLAN_ASYNC=$a LIMGCU_PLAN_CTAS=$g python tools/encode_time.py 2>&1 | tail -1
done; done
