#!/bin/bash
mkdir -p gpurun_out
for e in 0 1 2 3; do for cl in 8 2 1; do echo "experiment $e cluster $cl: $(LIMGCU_SCAN_EXPERIMENT=$e LIMGCU_SCAN_CLUSTER=$cl timeout 300 python tools/encode_time.py c2_4k_photo,c3_8k_rgba 6 2>&1 | tail -1)"; done; done | tee gpurun_out/exp_${1:-x}.txt
