#!/bin/bash
# cluster size x warps per CTA of the scan: single-frame encode time and frames in flight (batch throughput)
mkdir -p gpurun_out
for cfg in "8 8" "16 8" "16 4" "8 4"; do set -- $cfg
  echo "cluster $1 warps $2: $(LIMGCU_SCAN_CLUSTER=$1 LIMGCU_SCAN_WARPS=$2 timeout 300 python tools/encode_time.py c2_4k_photo,c4_4k_flatui,c5_1080p_frame0,c3_8k_rgba 8 2>&1 | tail -1)"
  LIMGCU_SCAN_CLUSTER=$1 LIMGCU_SCAN_WARPS=$2 timeout 300 python tools/batch_time.py c5_1080p_frame0 4,8,16 2>&1 | cut -c1-110
  LIMGCU_SCAN_CLUSTER=$1 LIMGCU_SCAN_WARPS=$2 timeout 300 python tools/batch_time.py c2_4k_photo 4,8 2>&1 | cut -c1-110
done | tee gpurun_out/cfg_${1:-x}.txt
