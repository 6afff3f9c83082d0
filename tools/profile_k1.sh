#!/bin/bash
MODEL=${1:-kip320_3x4_r4e2}
ncu --set full --clock-control none --import-source on -k regex:k_expand -s 21 -c 1 -f -o gpurun_out/prof_k1_${MODEL} \
    python tools/run_model.py $MODEL table_log2=26 max_states=20000000 > gpurun_out/prof_k1_${MODEL}.log 2>&1
ls -la gpurun_out | grep prof_k1
