#!/bin/bash
# 8 GPUs: the default bench (C3 z-decomposed) at N = 4 and 8 with per-rank phase times, the reference arm under torchrun
set -u
mkdir -p gpurun_out
for n in 8 4; do
  BENCH_RANK_PHASES=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n \
     bench.py --gpus $n --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_c3_n$n.json 2> gpurun_out/bench_c3_n$n.err
  echo "== N=$n"; cut -c1-1400 gpurun_out/bench_c3_n$n.json; grep "^rank" gpurun_out/bench_c3_n$n.err | sort -u; tail -2 gpurun_out/bench_c3_n$n.err
done
