#!/bin/bash
# 2 GPUs: peer-memory mailboxes -- NCCL-transport parity tests (mailboxes on), the same with CYLGPU_P2P=0, bench N=2 both ways
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_nccl.py -m gpu -q > gpurun_out/pytest_nccl.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_nccl.log
tail -30 gpurun_out/pytest_nccl.log | cut -c1-600
for p in 1 0; do
  CYLGPU_P2P=$p CYLGPU_PRESORT=1 BENCH_RANK_PHASES=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$p \
     bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --workload lwfa_1024x128_m2_ppc16 > gpurun_out/bench_small_n2_p2p$p.json 2> gpurun_out/bench_small_n2_p2p$p.err
  echo "== small N=2 P2P=$p"; cut -c1-200 gpurun_out/bench_small_n2_p2p$p.json; grep "^rank" gpurun_out/bench_small_n2_p2p$p.err | sort -u; tail -2 gpurun_out/bench_small_n2_p2p$p.err
done
BENCH_RANK_PHASES=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err
cut -c1-200 gpurun_out/bench_c3_n2.json; grep "^rank" gpurun_out/bench_c3_n2.err | sort -u; tail -2 gpurun_out/bench_c3_n2.err
