#!/bin/bash
# 4 GPUs, final build, defaults (no environment knobs): C3 strong scaling at N=4 and N=2
set -u
mkdir -p gpurun_out
run() {  # name n env...
  name=$1; n=$2; shift 2
  env "$@" BENCH_RANK_PHASES=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29577 \
     bench.py --gpus $n --steps 20 --warmup 5 --no-e2e > gpurun_out/bench_c3_$name.json 2> gpurun_out/bench_c3_$name.err
  echo "== $name"; python - gpurun_out/bench_c3_$name.json <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value %.4e ms/step %.3f"%(d['value'],d['ms_per_step']), d['config'].get('neighbour_links'), d['config'].get('exchange_capacity'))
except Exception as e:
    print("ERR", e)
P
  grep "^rank" gpurun_out/bench_c3_$name.err | sort -u | cut -c1-200
  grep -i "error\|overflow" gpurun_out/bench_c3_$name.err | head -3 | cut -c1-300
}
run r2g_n4_default 4 DUMMY=1
run r2g_n2_default 2 DUMMY=1
