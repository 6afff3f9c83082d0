#!/bin/bash
# Round 2, fourth GPU pass: oracle-validated traces + new models; site-group sizes; cross-kernel L2 prefetch with small chunks.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -3 gpurun_out/r2d_pytest.log
M=kip320_3x4_r4e3
timeout 1200 python tools/bench_variants.py $M 3 '{"tag":"base"}' '{"tag":"g800","model":"kip320_3x4_r4e3@g800"}' '{"tag":"g3000","model":"kip320_3x4_r4e3@g3000"}' \
    '{"tag":"g1600s32","model":"kip320_3x4_r4e3@g1600s32"}' \
    '{"tag":"chunk2M","chunk_states":2000000}' '{"tag":"chunk500k","chunk_states":500000}' \
    '{"tag":"pf_2M","prefetch":true,"chunk_states":2000000}' '{"tag":"pf_1M","prefetch":true,"chunk_states":1000000}' \
    '{"tag":"pf_500k","prefetch":true,"chunk_states":500000}' '{"tag":"pf_250k","prefetch":true,"chunk_states":250000}' \
    '{"tag":"pf_nochunk","prefetch":true}' \
    > gpurun_out/r2d_variants.jsonl 2> gpurun_out/r2d_variants.err
timeout 300 python tools/bench_variants.py asyncisr_deep 2 '{"tag":"base"}' '{"tag":"pf_500k","prefetch":true,"chunk_states":500000}' >> gpurun_out/r2d_variants.jsonl 2>> gpurun_out/r2d_variants.err
cat gpurun_out/r2d_variants.jsonl
ls -la gpurun_out/ | grep r2d
