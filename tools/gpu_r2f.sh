#!/bin/bash
# Round 2, NOT RUN (the round's GPU budget was spent before it): the first pass to make when a GPU is available again.
#   gpurun --timeout 2400 -- 'bash tools/gpu_r2f.sh'          (one B200)
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_r2f.sh multi'
set -x
mkdir -p gpurun_out
if [ "$1" = "multi" ]; then
  # device-side round/level synchronisation on 2 GPUs: the 2-GPU tests, then the bench line (its first warm-up run is the
  # acceptance run; config.round_sync / round_sync_note say which path ran), then the barrier path for comparison
  timeout 900 python -m pytest tests -m gpu -x -q -k "two_gpu or two_gpus or workers" > gpurun_out/r2f_pytest_multi.log 2>&1
  tail -3 gpurun_out/r2f_pytest_multi.log
  N=$(nvidia-smi -L | wc -l)
  for mode in "" "--no-p2p"; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
        bench.py --gpus $N --steps 10 --warmup 3 $mode 2> gpurun_out/r2f_scale_n${N}${mode}.err | grep '^{' > gpurun_out/r2f_scale_n${N}${mode}.json
    cut -c1-400 gpurun_out/r2f_scale_n${N}${mode}.json
  done
  exit 0
fi
# 1. the whole suite; the tests written after the last GPU pass run last (tests/conftest.py, tests/test_zz_round2_late.py)
timeout 2000 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -5 gpurun_out/r2f_pytest.log
# 2. the headline line, and the opt-in variants whose round-2 timings were lost (two-stream K1/K2 overlap, L2 prefetch)
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
M=kip320_3x4_r4e3
timeout 900 python tools/bench_variants.py $M 3 '{"tag":"base"}' '{"tag":"overlap","overlap":true}' \
    '{"tag":"pf_1M","prefetch":true,"chunk_states":1000000}' '{"tag":"pf_500k","prefetch":true,"chunk_states":500000}' \
    > gpurun_out/r2f_variants.jsonl 2> gpurun_out/r2f_variants.err
# 3. the four protocol variants at headline bounds and config #5 (throughput next to the headline)
for m in trunchw_3x4_r3e3 kip101_3x4_r3e3 kip279_3x4_r3e3 firsttry_3x4_r3e3 asyncisr_deep; do
  timeout 300 python tools/bench_variants.py $m 2 '{"tag":"base"}' >> gpurun_out/r2f_variants.jsonl 2>> gpurun_out/r2f_variants.err
done
cat gpurun_out/r2f_variants.jsonl
# 4. launch list + one full capture of each dominant kernel (same launches as r2c: second chunk of level 22)
ARGS="table_log2=30 max_states=347300000"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f_launches_$M.csv \
    python tools/run_model.py $M $ARGS > gpurun_out/r2f_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_expand -s 24 -c 1 -f -o gpurun_out/r2f_expand_$M \
    python tools/run_model.py $M $ARGS > gpurun_out/r2f_prof_expand.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_insert -s 25 -c 1 -f -o gpurun_out/r2f_insert_$M \
    python tools/run_model.py $M $ARGS > gpurun_out/r2f_prof_insert.log 2>&1
ls -la gpurun_out | grep r2f
