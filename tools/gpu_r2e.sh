#!/bin/bash
# Round 2, fifth GPU pass: full suite; K1/K2 two-stream overlap x CTA shapes; config #4 (5 brokers, symmetry) sizes; bench.py
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -3 gpurun_out/r2e_pytest.log
M=kip320_3x4_r4e3
timeout 1200 python tools/bench_variants.py $M 3 '{"tag":"base"}' '{"tag":"overlap","overlap":true}' \
    '{"tag":"b768","lib":"b768"}' '{"tag":"b768_overlap","lib":"b768","overlap":true}' \
    '{"tag":"b512x1","lib":"b512x1"}' '{"tag":"b512x1_overlap","lib":"b512x1","overlap":true}' \
    > gpurun_out/r2e_variants.jsonl 2> gpurun_out/r2e_variants.err
timeout 300 python tools/bench_variants.py asyncisr_deep 2 '{"tag":"base"}' '{"tag":"overlap","overlap":true}' >> gpurun_out/r2e_variants.jsonl 2>> gpurun_out/r2e_variants.err
timeout 300 python tools/bench_variants.py kip320sym_3x4_r4e3 2 '{"tag":"base"}' '{"tag":"overlap","overlap":true}' >> gpurun_out/r2e_variants.jsonl 2>> gpurun_out/r2e_variants.err
timeout 300 python tools/bench_variants.py kip320sym_5brokers_r1e2 2 '{"tag":"base"}' >> gpurun_out/r2e_variants.jsonl 2>> gpurun_out/r2e_variants.err
timeout 900 python tools/bench_variants.py kip320sym_5brokers 1 '{"tag":"c4_r2e2","table_log2":32,"max_states":1600000000}' >> gpurun_out/r2e_variants.jsonl 2>> gpurun_out/r2e_variants.err
cat gpurun_out/r2e_variants.jsonl
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
cat gpurun_out/r2e_bench.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2e_bench_ref.json 2>> gpurun_out/r2e_bench.err
cat gpurun_out/r2e_bench_ref.json
ls -la gpurun_out/ | grep r2e
