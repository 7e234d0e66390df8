#!/bin/bash
# GPU pass for a change of the merge scan (short form): parity tests, encode times, profile counters of the 4K photo frame
mkdir -p gpurun_out
TAG=${1:-x}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
for cfg in ${CONFIGS:-"8 0" "16 0" "8 4"}; do set -- $cfg; echo "cluster $1 experiment $2: $(LIMGCU_SCAN_CLUSTER=$1 LIMGCU_SCAN_EXPERIMENT=$2 timeout 300 python tools/encode_time.py c2_4k_photo,c4_4k_flatui,c5_1080p_frame0,c3_8k_rgba 8 2>&1 | tail -1)"; done | tee gpurun_out/scan_sweep_$TAG.txt
LIMGCU_LIB=limg_b200/liblimgcu_prof.so timeout 300 python tools/phase_times.py c2_4k_photo,c4_4k_flatui > gpurun_out/phase_$TAG.txt 2>&1; grep -E "expansion parts|profile kcycles|emitting|total|four-way attempts" gpurun_out/phase_$TAG.txt
