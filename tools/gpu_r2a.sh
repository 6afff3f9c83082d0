#!/bin/bash
# Round 2, first GPU pass: parity suite on the new two-phase expand kernel, A/B against the round-1 one-phase
# kernel, duplicate-filter sweep, ncu launch list + full captures of K1 and K2.  ONE GPU.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
M=kip320_3x4_r4e3
timeout 600 python tools/bench_variants.py $M 3 '{"tag":"two_phase"}' '{"tag":"one_phase","lib":"1p","one_phase":true}' \
   '{"tag":"dcache20","dcache_log2":20}' '{"tag":"dcache22","dcache_log2":22}' '{"tag":"dcache23","dcache_log2":23}' '{"tag":"dcache24","dcache_log2":24}' \
   > gpurun_out/r2a_variants.jsonl 2> gpurun_out/r2a_variants.err
cat gpurun_out/r2a_variants.jsonl
timeout 300 python tools/bench_variants.py asyncisr_deep 2 '{"tag":"two_phase"}' '{"tag":"dcache23","dcache_log2":23}' >> gpurun_out/r2a_variants.jsonl 2>> gpurun_out/r2a_variants.err
timeout 300 python tools/bench_variants.py kip320sym_3x4_r4e3 2 '{"tag":"two_phase"}' >> gpurun_out/r2a_variants.jsonl 2>> gpurun_out/r2a_variants.err
tail -3 gpurun_out/r2a_variants.jsonl
ARGS="table_log2=30 max_states=347300000"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2a_launches_$M.csv \
    python tools/run_model.py $M $ARGS > gpurun_out/r2a_launches_$M.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_expand -s 24 -c 1 -f -o gpurun_out/r2a_expand_$M \
    python tools/run_model.py $M $ARGS > gpurun_out/r2a_prof_expand.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_insert -s 25 -c 1 -f -o gpurun_out/r2a_insert_$M \
    python tools/run_model.py $M $ARGS > gpurun_out/r2a_prof_insert.log 2>&1
ls -la gpurun_out/
