#!/bin/bash
# full ncu capture of one kernel of the encode (regex $1, skip $2 launches), after the same command ran without ncu
mkdir -p gpurun_out
python tools/encode_time.py c2_4k_photo 2 > gpurun_out/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-2} -c 1 -f -o gpurun_out/prof_${3:-one} python tools/encode_time.py c2_4k_photo 2 > gpurun_out/ncu_one.log 2>&1
echo "ncu rc=$?"
