#!/bin/bash
# 2 GPUs: the whole GPU suite (fabric + NCCL + mailboxes), then the N=2 bench with per-rank phases
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | cut -c1-400 | tee gpurun_out/pytest_gpu_n2.txt
run() {  # name n env...
  name=$1; n=$2; shift 2
  env "$@" BENCH_RANK_PHASES=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29577 \
     bench.py --gpus $n --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_y_$name.json 2> gpurun_out/bench_y_$name.err
  echo "== $name"; python - gpurun_out/bench_y_$name.json <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value %.4e ms/step %.3f"%(d['value'],d['ms_per_step']), d['config'].get('neighbour_links'), d['config'].get('exchange_capacity'))
except Exception as e:
    print("ERR", e)
P
  grep "^rank" gpurun_out/bench_y_$name.err | sort -u | cut -c1-200
  grep -i "error\|overflow" gpurun_out/bench_y_$name.err | head -3 | cut -c1-300
}
run n2_wide 2 CYLGPU_P2P=particles
