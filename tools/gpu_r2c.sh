#!/bin/bash
# Round 2, third GPU pass: 128-bit layout (W=2) + exact 16-byte-key set; CTA-shape / group-size / bucket-size variants.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -3 gpurun_out/r2c_pytest.log
M=kip320_3x4_r4e3
timeout 900 python tools/bench_variants.py $M 3 '{"tag":"w2_exact"}' '{"tag":"b512","lib":"b512"}' '{"tag":"bs4","lib":"bs4"}' \
    '{"tag":"g200","model":"kip320_3x4_r4e3@g200"}' '{"tag":"g800","model":"kip320_3x4_r4e3@g800"}' '{"tag":"load_hi","table_log2":29}' \
    > gpurun_out/r2c_variants.jsonl 2> gpurun_out/r2c_variants.err
timeout 300 python tools/bench_variants.py asyncisr_deep 2 '{"tag":"w2_exact"}' >> gpurun_out/r2c_variants.jsonl 2>> gpurun_out/r2c_variants.err
timeout 300 python tools/bench_variants.py kip320sym_3x4_r4e3 2 '{"tag":"w2_exact"}' >> gpurun_out/r2c_variants.jsonl 2>> gpurun_out/r2c_variants.err
cat gpurun_out/r2c_variants.jsonl
ARGS="table_log2=30 max_states=347300000"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_expand -s 24 -c 1 -f -o gpurun_out/r2c_expand_$M \
    python tools/run_model.py $M $ARGS > gpurun_out/r2c_prof_expand.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_insert -s 25 -c 1 -f -o gpurun_out/r2c_insert_$M \
    python tools/run_model.py $M $ARGS > gpurun_out/r2c_prof_insert.log 2>&1
ls -la gpurun_out/ | grep r2c
