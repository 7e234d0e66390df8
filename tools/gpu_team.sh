#!/bin/bash
# teams of warps per block row (LIMGCU_SCAN_TEAM) x cluster size: encode times, first-try failures, then the parity tests of the merge with the given team
mkdir -p gpurun_out
TAG=${1:-x}
for t in 1 2 4; do for cl in 8 16; do
  echo "team $t cluster $cl: $(LIMGCU_SCAN_TEAM=$t LIMGCU_SCAN_CLUSTER=$cl timeout 120 python tools/encode_time.py c2_4k_photo,c4_4k_flatui,c5_1080p_frame0,c3_8k_rgba 6 2>&1 | tail -1)"
done; done | tee gpurun_out/team_$TAG.txt
LIMGCU_SCAN_TEAM=${2:-2} LIMGCU_SCAN_CLUSTER=16 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/team_$TAG.txt
