#!/bin/bash
# GPU pass for a change of the merge scan: parity tests, encode times per scan back end (cluster size; 0 = scan over global memory), profile counters, row time stamps
mkdir -p gpurun_out
TAG=${1:-x}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
for cl in ${CLUSTERS:-16 8 4 0}; do echo "cluster $cl: $(LIMGCU_SCAN_CLUSTER=$cl timeout 300 python tools/encode_time.py c2_4k_photo,c4_4k_flatui,c5_1080p_frame0,c3_8k_rgba 8 2>&1 | tail -1)"; done | tee gpurun_out/scan_sweep_$TAG.txt
if [ -f limg_b200/liblimgcu_prof.so ]; then
  LIMGCU_LIB=limg_b200/liblimgcu_prof.so timeout 300 python tools/phase_times.py c2_4k_photo,c4_4k_flatui,c3_8k_rgba > gpurun_out/phase_$TAG.txt 2>&1; grep -v "^   \(expansion\|strips\|on-demand\)" gpurun_out/phase_$TAG.txt | tail -30
  LIMGCU_PLAN_ASYNC=1 LIMGCU_LIB=limg_b200/liblimgcu_prof.so timeout 300 python tools/row_times.py c2_4k_photo > gpurun_out/rows_$TAG.txt 2>&1; tail -70 gpurun_out/rows_$TAG.txt
fi
