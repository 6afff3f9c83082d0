#!/bin/bash
# 1 -> 8 GPU scaling run (one box).  Prints one JSON line per N into gpurun_out/scale_nN.json
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
      bench.py --gpus $N --steps 4 --warmup 3 2> gpurun_out/scale_n$N.err | grep '^{' > gpurun_out/scale_n$N.json
  cut -c1-260 gpurun_out/scale_n$N.json; tail -2 gpurun_out/scale_n$N.err | cut -c1-300
done
