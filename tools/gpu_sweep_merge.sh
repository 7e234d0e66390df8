#!/bin/bash
# one-at-a-time sweep of the scan's knobs (all exact: they only move work between the speculative pass and its verification)
run() { echo -n "$1: "; env $1 python tools/encode_time.py c2_4k_photo,c4_4k_flatui,c3_8k_rgba 6 2>&1 | tail -1; }
run LIMGCU_MERGE_MARGIN=8
for v in 2 4 6 12; do run LIMGCU_MERGE_MARGIN=$v; done
for v in 0 4 16 32; do run LIMGCU_MERGE_SPEC=$v; done
for v in 4 8 32; do run LIMGCU_MERGE_GAP=$v; done
for v in 8 24 32; do run LIMGCU_PLAN_EXTW=$v; done
for v in 3 8; do run LIMGCU_PLAN_SYML=$v; done
for v in 8 20; do run LIMGCU_PLAN_SYMR=$v; done
for v in 8 24; do run LIMGCU_PLAN_SYMD=$v; done
