#!/bin/bash
# GPU pass of a change: parity tests, bench line (no profiler), optional extra command
mkdir -p gpurun_out
TAG=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 10 --warmup 3 ${2:-} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
