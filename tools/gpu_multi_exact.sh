#!/bin/bash
# whole-image-exact row bands on N GPUs (NCCL), next to the single-GPU run of the same mode
N=${1:-2}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 5 --warmup 2 --workload c3_8k_rgba --mode rowband_exact 2>gpurun_out/exact_n1.err | tee gpurun_out/bench_exact_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 5 --warmup 2 --workload c3_8k_rgba --mode rowband_exact 2>gpurun_out/exact_n$N.err | tee gpurun_out/bench_exact_n$N.json
tail -3 gpurun_out/exact_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --steps 5 --warmup 2 --workload c2_4k_photo --mode rowband_exact 2>>gpurun_out/exact_n$N.err | tee gpurun_out/bench_exact_4k_n$N.json
