#!/bin/bash
# source-level profile of the scan's per-seed work: in sequential mode one warp works at a time, so its samples are not drowned by pollers
export LIMGCU_MERGE_MODE=seq
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload c5_1080p_photo > gpurun_out/plain_seq.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_merge_wave -s 11 -c 1 -o gpurun_out/prof_wave_seq python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload c5_1080p_photo > gpurun_out/ncu_wave_seq.log 2>&1
echo rc=$?
