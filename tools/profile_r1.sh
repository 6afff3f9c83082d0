#!/bin/bash
# ncu evidence for round 1 (run under gpurun, ONE GPU).  Numbers printed under ncu are never bench values.
set -x
MODEL=${1:-kip320_3x4_r4e2}
ARGS="table_log2=26 max_states=20000000"
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${MODEL}.csv \
    python tools/run_model.py $MODEL $ARGS > gpurun_out/launches_${MODEL}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_expand -s 21 -c 2 -f -o gpurun_out/prof_expand_${MODEL} \
    python tools/run_model.py $MODEL $ARGS > gpurun_out/prof_expand_${MODEL}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_insert -s 22 -c 2 -f -o gpurun_out/prof_insert_${MODEL} \
    python tools/run_model.py $MODEL $ARGS > gpurun_out/prof_insert_${MODEL}.log 2>&1
ls -la gpurun_out/
ncu --set full --clock-control none --import-source on -k regex:k_invariants -s 22 -c 2 -f -o gpurun_out/prof_invariants_${MODEL} \
    python tools/run_model.py $MODEL $ARGS > gpurun_out/prof_invariants_${MODEL}.log 2>&1
