#!/bin/bash
# pass 1 staged by the TMA (k_pass1_tma) against plain loads (k_pass1): parity tests, phase times
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for t in 1 0; do echo "pass1 tma $t:"; LIMGCU_PASS1_TMA=$t python tools/phase_times.py c2_4k_photo,c3_8k_rgba,c1_512_gradient 2>&1 | grep -E "^c[0-9]" | cut -c1-220; done
