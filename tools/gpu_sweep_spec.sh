#!/bin/bash
# knob sweep of the scan: speculation lookahead (LIMGCU_MERGE_SPEC), margin, cluster size
mkdir -p gpurun_out
TAG=${1:-x}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
for cl in 8 16; do for spec in 8 32 128 100000; do for m in 8 4; do
echo "cluster $cl spec $spec margin $m: $(LIMGCU_SCAN_CLUSTER=$cl LIMGCU_MERGE_SPEC=$spec LIMGCU_MERGE_MARGIN=$m timeout 300 python tools/encode_time.py c2_4k_photo,c4_4k_flatui,c3_8k_rgba 6 2>&1 | tail -1)"
done; done; done | tee gpurun_out/sweep_spec_$TAG.txt
