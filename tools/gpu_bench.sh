#!/bin/bash
# GPU pass: parity tests, the default bench line, the batch mode (short) and the exact row bands on one GPU
mkdir -p gpurun_out
TAG=${1:-x}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$TAG.json'))
for k in ('value','ms_per_step','encode_ms','decode_ms','e2e','e2e_dropin','roofline','batch_1080p','north_star_8k_rgb','cpu_baseline','merge'):
    print(k, json.dumps(d.get(k))[:700])
"
timeout 600 python bench.py --mode batch --frames 128 --lanes 8 --steps 3 --warmup 3 > gpurun_out/bench_batch_$TAG.json 2> gpurun_out/bench_batch_$TAG.err; echo "batch rc=$?"; tail -2 gpurun_out/bench_batch_$TAG.err; cut -c1-1500 gpurun_out/bench_batch_$TAG.json
timeout 600 python bench.py --mode rowband_exact --workload c3_8k_rgba --steps 3 --warmup 3 > gpurun_out/bench_exact_$TAG.json 2> gpurun_out/bench_exact_$TAG.err; echo "exact rc=$?"; tail -2 gpurun_out/bench_exact_$TAG.err; cut -c1-900 gpurun_out/bench_exact_$TAG.json
