#!/bin/bash
# decode kernel variants (tools/decode_time.py) + quick GPU test pass + a short bench line (clock sampler check)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python tools/decode_time.py 0,40,41,42,43,80,81,82,83 > gpurun_out/decode_variants.log 2>&1; echo "decode rc=$?"; cat gpurun_out/decode_variants.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cat gpurun_out/bench_quick.json
