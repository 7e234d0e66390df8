#!/bin/bash
# bench line (no profiler), the ncu launch list of the same command, then one full capture of the top kernels
mkdir -p gpurun_out
TAG=${1:-r1_c}
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_merge_wave -s 9 -c 1 -f -o gpurun_out/prof_wave_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_wave.log 2>&1
echo "ncu wave rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_decode_tile" -s 4 -c 1 -f -o gpurun_out/prof_decode_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_dec.log 2>&1
echo "ncu decode rc=$?"
ls -la gpurun_out/*.ncu-rep
