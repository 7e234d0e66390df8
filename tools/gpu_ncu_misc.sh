#!/bin/bash
# full ncu capture of throughput kernels of one 4K encode (after the same command ran without ncu): regex $1, skip $2, count $3, tag $4
mkdir -p gpurun_out
python tools/encode_time.py c2_4k_photo 3 > gpurun_out/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"${1:-k_pred_window|k_encode_large}" -s ${2:-6} -c ${3:-3} -f -o gpurun_out/prof_${4:-misc} python tools/encode_time.py c2_4k_photo 3 > gpurun_out/ncu_misc.log 2>&1
echo "ncu rc=$?"
