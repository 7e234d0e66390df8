#!/bin/bash
# multi-GPU pass (N = $1): the driver's frames mode, BASELINE.json config 5 (batch of 1024 1080p frames), config 3 (8K RGBA row bands, independent and whole-image-exact)
N=${1:-2}; TAG=${2:-x}; PORT=29511
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N "$@" > gpurun_out/bench_n${N}_${name}_$TAG.json 2> gpurun_out/bench_n${N}_${name}_$TAG.err; echo "$name rc=$?"; PORT=$((PORT+1)); python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_n${N}_${name}_$TAG.json"))
    print("  ", {k: d.get(k) for k in ("value","ms_per_step","n_gpus","identical_to_single_gpu_encode","bands_equal_the_reference_run_per_band")}, "e2e", (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("   no line:", e); print(open("gpurun_out/bench_n${N}_${name}_$TAG.err").read()[-1500:])
PY
}
for m in ${MODES:-frames batch rowband exact}; do
  case $m in
    frames) run frames --steps 10 --warmup 3 ;;
    batch) run batch --mode batch --frames 1024 --lanes 8 --steps 3 --warmup 3 ;;
    rowband) run rowband --mode rowband --workload c3_8k_rgba --steps 5 --warmup 3 ;;
    exact) run exact --mode rowband_exact --workload c3_8k_rgba --steps 5 --warmup 3 ;;
  esac
done
