#!/bin/bash
# GPU pass: parity tests and the default bench line (what the driver runs at round end), both arms
mkdir -p gpurun_out
TAG=${1:-x}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$TAG.json'))
for k in ('value','ms_per_step','encode_ms','decode_ms','e2e','e2e_dropin','roofline','roofline_decode','batch_1080p','north_star_8k_rgb','cpu_baseline','merge','unmerged_encode'):
    print(k, json.dumps(d.get(k))[:900])
"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_ref_$TAG.json
