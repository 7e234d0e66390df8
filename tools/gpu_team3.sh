#!/bin/bash
# more warps per CTA (register-capped builds) x teams
mkdir -p gpurun_out
TAG=${1:-x}
for cfg in "cu 1" "cu 2" "cu 4" "cu_w12 1" "cu_w12 2" "cu_w12 3" "cu_w12 4" "cu_w16 1" "cu_w16 2" "cu_w16 4"; do set -- $cfg
  echo "lib $1 team $2 cluster 16: $(LIMGCU_LIB=limg_b200/liblimg$1.so LIMGCU_SCAN_TEAM=$2 LIMGCU_SCAN_CLUSTER=16 timeout 120 python tools/encode_time.py c2_4k_photo,c4_4k_flatui,c5_1080p_frame0,c3_8k_rgba 6 2>&1 | tail -1)"
done | tee gpurun_out/team3_$TAG.txt
export LIMGCU_LIB=limg_b200/liblimgcu_prof.so
for t in 2 4; do
  LIMGCU_SCAN_TEAM=$t LIMGCU_SCAN_CLUSTER=16 timeout 120 python tools/team_stats.py c2_4k_photo 2>&1 | tail -1
  LIMGCU_SCAN_TEAM=$t LIMGCU_SCAN_CLUSTER=16 timeout 120 python tools/row_times.py c2_4k_photo 2>&1 | grep -E "slope|kernel span"
done | tee gpurun_out/team_stats_$TAG.txt
