#!/bin/bash
# final single-GPU pass of a round: default bench line (both arms), phase times of all five configs, ncu launch lists, one ncu --set full capture of the scan
mkdir -p gpurun_out
TAG=${1:-x}
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
timeout 300 python tools/phase_times.py c1_512_gradient,c2_4k_photo,c3_8k_rgba,c4_4k_flatui,c5_1080p_frame0 > gpurun_out/phase_$TAG.txt 2>&1; echo "phase rc=$?"
for w in c5_1080p_frame0 c2_4k_photo; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${w}_$TAG.csv python tools/encode_time.py $w 2 > gpurun_out/ncu_launch.log 2>&1
  echo "$w launch list rc=$?"
done
ncu --set full --clock-control none --import-source on -k regex:k_merge_cta -c 1 -o gpurun_out/prof_${TAG}_merge_cta python tools/encode_time.py c2_4k_photo 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
for k in ('value','ms_per_step','encode_ms','decode_ms','e2e','e2e_dropin','roofline','north_star_8k_rgb','cpu_baseline','merge'):
    print(k, json.dumps(d.get(k))[:600])
PY
