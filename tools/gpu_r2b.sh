#!/bin/bash
# Round 2, second GPU pass: K1 with per-thread pair bookkeeping (no ballot census).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -3 gpurun_out/r2b_pytest.log
M=kip320_3x4_r4e3
timeout 600 python tools/bench_variants.py $M 3 '{"tag":"two_phase_v2"}' > gpurun_out/r2b_variants.jsonl 2> gpurun_out/r2b_variants.err
timeout 300 python tools/bench_variants.py asyncisr_deep 2 '{"tag":"two_phase_v2"}' >> gpurun_out/r2b_variants.jsonl 2>> gpurun_out/r2b_variants.err
timeout 300 python tools/bench_variants.py kip320sym_3x4_r4e3 2 '{"tag":"two_phase_v2"}' >> gpurun_out/r2b_variants.jsonl 2>> gpurun_out/r2b_variants.err
timeout 300 python tools/bench_variants.py frl_3x4x3 2 '{"tag":"two_phase_v2"}' >> gpurun_out/r2b_variants.jsonl 2>> gpurun_out/r2b_variants.err
cat gpurun_out/r2b_variants.jsonl
ARGS="table_log2=30 max_states=347300000"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_expand -s 24 -c 1 -f -o gpurun_out/r2b_expand_$M \
    python tools/run_model.py $M $ARGS > gpurun_out/r2b_prof_expand.log 2>&1
ls -la gpurun_out/ | grep r2b
