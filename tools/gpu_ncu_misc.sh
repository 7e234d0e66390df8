#!/bin/bash
# full ncu capture of the throughput kernels of one encode (after the same command ran without ncu)
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_pass1|k_finalize|k_encode_small|k_merge_verify$|k_pred_window" -s 10 -c 6 -f -o gpurun_out/prof_misc python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_misc.log 2>&1
echo "ncu rc=$?"
