#!/bin/bash
# multi-GPU smoke of bench.py as the driver launches it (torchrun, one rank per GPU), both sharding modes + the reference arm
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_frames.json 2> gpurun_out/bench_n${N}_frames.err; echo "frames rc=$?"; cat gpurun_out/bench_n${N}_frames.json; tail -3 gpurun_out/bench_n${N}_frames.err
$TR bench.py --gpus $N --steps 5 --warmup 3 --workload c3_8k_rgba --mode rowband > gpurun_out/bench_n${N}_rowband.json 2> gpurun_out/bench_n${N}_rowband.err; echo "rowband rc=$?"; cat gpurun_out/bench_n${N}_rowband.json; tail -3 gpurun_out/bench_n${N}_rowband.err
python bench.py --gpus 1 --steps 5 --warmup 3 --workload c3_8k_rgba --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_n1_c3.json
$TR bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_n${N}_ref.json 2> gpurun_out/bench_n${N}_ref.err; echo "reference rc=$?"; cat gpurun_out/bench_n${N}_ref.json; tail -3 gpurun_out/bench_n${N}_ref.err
