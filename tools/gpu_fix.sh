#!/bin/bash
# after a scan change: first-try failure sweep, encode times, row bands of the 8K frame, the merge parity tests
mkdir -p gpurun_out
TAG=${1:-x}
timeout 600 python tools/fail_sweep.py 2>&1 | tail -8 | tee gpurun_out/fail_sweep_$TAG.txt
timeout 300 python tools/encode_time.py c2_4k_photo,c4_4k_flatui,c5_1080p_frame0,c3_8k_rgba 6 2>&1 | tail -6 | tee gpurun_out/encode_time_$TAG.txt
timeout 300 python tools/band_tries.py c3_8k_rgba 8 2>&1 | tail -3 | tee gpurun_out/band_tries_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
