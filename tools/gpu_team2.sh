#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-x}
export LIMGCU_LIB=limg_b200/liblimgcu_prof.so
for t in 1 2 4; do
  LIMGCU_SCAN_TEAM=$t LIMGCU_SCAN_CLUSTER=16 timeout 120 python tools/team_stats.py c2_4k_photo 2>&1 | tail -1
  LIMGCU_SCAN_TEAM=$t LIMGCU_SCAN_CLUSTER=16 timeout 120 python tools/row_times.py c2_4k_photo 2>&1 | grep -E "slope|kernel span"
done | tee gpurun_out/team_stats_$TAG.txt
for sp in 8 24 1000; do
  echo "spec $sp team 2: $(LIMGCU_LIB=limg_b200/liblimgcu.so LIMGCU_MERGE_SPEC=$sp LIMGCU_SCAN_TEAM=2 LIMGCU_SCAN_CLUSTER=16 timeout 120 python tools/encode_time.py c2_4k_photo 6 2>&1 | tail -1)"
  LIMGCU_MERGE_SPEC=$sp LIMGCU_SCAN_TEAM=2 LIMGCU_SCAN_CLUSTER=16 timeout 120 python tools/team_stats.py c2_4k_photo 2>&1 | tail -1
done | tee -a gpurun_out/team_stats_$TAG.txt
