#!/bin/bash
# ncu launch list (device time of every kernel, cold cache, serialised) of one encode + decode per config: shares, not absolutes
mkdir -p gpurun_out
TAG=${1:-x}
for w in c5_1080p_frame0 c2_4k_photo; do
  python tools/encode_time.py $w 2 > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${w}_$TAG.csv python tools/encode_time.py $w 2 > gpurun_out/ncu_launch.log 2>&1
  echo "$w ncu rc=$?"
done
