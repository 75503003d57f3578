#!/bin/bash
# 8 GPUs: default bench at N=8 with mailboxes on / off and pre-sort on / off, then N=4
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
    print("value %.4e ms/step %.3f"%(d['value'],d['ms_per_step']), {k:round(v,3) for k,v in d['phase_ms_per_step'].items()}, d['config'].get('neighbour_links'))
except Exception as e:
    print("ERR", e)
P
  tail -2 gpurun_out/bench_c3_$name.err | cut -c1-300
}
run n8 8 CYLGPU_P2P=1
run n8_nop2p 8 CYLGPU_P2P=0
run n8_nopresort 8 CYLGPU_PRESORT=0
run n4 4 CYLGPU_P2P=1
grep "^rank" gpurun_out/bench_c3_n8.err | sort -u
