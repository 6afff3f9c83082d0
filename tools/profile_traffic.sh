#!/bin/bash
# one ncu --set full capture per kernel on the headline model's widest level (for roofline.traffic)
MODEL=${1:-kip320_3x4_r4e3}
ARGS="table_log2=30 max_states=400000000"
for k in k_expand:24 k_insert:25 k_invariants:25; do
  name=${k%%:*}; skip=${k##*:}
  ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -f -o gpurun_out/traffic_${name}_${MODEL} \
      python tools/run_model.py $MODEL $ARGS > gpurun_out/traffic_${name}_${MODEL}.log 2>&1
done
ls -la gpurun_out | grep traffic
