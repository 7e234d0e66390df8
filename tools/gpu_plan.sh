#!/bin/bash
# plan kernels vs the cluster scan: CTAs per SM of the asynchronous bitmap kernels, async modes
mkdir -p gpurun_out
for cfg in "1 2" "1 4" "1 6" "1 8" "2 6" "0 6"; do set -- $cfg
  echo "plan async $1 ctas $2: $(LIMGCU_MERGE_MARGIN=0 LIMGCU_PLAN_ASYNC=$1 LIMGCU_PLAN_CTAS=$2 timeout 300 python tools/encode_time.py c2_4k_photo,c4_4k_flatui,c5_1080p_frame0,c3_8k_rgba 8 2>&1 | tail -1)"
done | tee gpurun_out/plan_sweep_${1:-x}.txt
