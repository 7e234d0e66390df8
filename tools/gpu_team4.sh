#!/bin/bash
# stage 0 alone (experiment 8: stage 1 emits nothing; the result is wrong, only the time stamps of the first try count)
mkdir -p gpurun_out
TAG=${1:-x}
export LIMGCU_LIB=limg_b200/liblimgcu_prof.so
for cfg in c2_4k_photo c5_1080p_frame0; do for t in 1 2 4; do for cl in 8 16; do
  echo "== $cfg team $t cluster $cl, stage 0 alone"
  LIMGCU_SCAN_EXPERIMENT=8 LIMGCU_SCAN_TEAM=$t LIMGCU_SCAN_CLUSTER=$cl timeout 120 python tools/row_times.py $cfg 2>&1 | grep -E "slope|kernel span|Error|error" | head -4
done; done; done | tee gpurun_out/team4_$TAG.txt
