#!/bin/bash
# safe columns (k_plan_safe) on / off: parity tests, encode times, failing bands
mkdir -p gpurun_out
TAG=${1:-x}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
for cfg in "1 2" "1 1" "1 0" "1 4" "0 8"; do set -- $cfg
  echo "safe $1 margin $2: $(LIMGCU_MERGE_SAFE=$1 LIMGCU_MERGE_MARGIN=$2 timeout 300 python tools/encode_time.py c2_4k_photo,c4_4k_flatui,c5_1080p_frame0,c3_8k_rgba 8 2>&1 | tail -1)"
done | tee gpurun_out/safe_sweep_$TAG.txt
python tools/band_tries.py c3_8k_rgba 8 | tail -4
for i in 1 2 3; do python tools/encode_time.py c4_4k_flatui 12 | tail -1; done
