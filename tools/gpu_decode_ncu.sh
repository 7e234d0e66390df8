#!/bin/bash
# one full ncu capture of the decode kernel at 8K RGB (after the same command exited 0 without ncu)
mkdir -p gpurun_out
V=${1:-40}
python tools/decode_time.py $V "8K RGB" > gpurun_out/decode_plain.log 2>&1 || exit 1
cat gpurun_out/decode_plain.log
ncu --set full --clock-control none --import-source on -k regex:k_decode -s 5 -c 1 -f -o gpurun_out/prof_decode_v$V python tools/decode_time.py $V "8K RGB" > gpurun_out/ncu_decode.log 2>&1
echo "ncu rc=$?"
