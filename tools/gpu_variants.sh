#!/bin/bash
# kernel-tuning experiment: strip width / patch staging variants of the fused push -- parity subset, then C3 and C2 kernel times
set -u
mkdir -p gpurun_out
: > gpurun_out/variants.txt
for v in base s16 s24 s29 s32; do
  export CYLGPU_LIB=$PWD/cylindrical_epoch_b200/libcylgpu_$v.so
  python -m pytest tests/test_gpu_parity.py -m gpu -q -k "lwfa_steps or thermal_periodic or every_mode_count or sparse" > gpurun_out/pytest_$v.log 2>&1
  echo "$v pytest: $(tail -1 gpurun_out/pytest_$v.log)" >> gpurun_out/variants.txt
  for wl in lwfa_8192x512_m2_ppc32 thermal_2048x256_m2_ppc64 modes5_4096x512_m5_ppc16; do
    timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload $wl 2> gpurun_out/bench_$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v $wl value %.3e  ms/step %.3f push_kernel %.3f ms sort %.3f frac %.4f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['phase_ms_per_step']['sort'], d['roofline']['frac']))" >> gpurun_out/variants.txt 2>&1
  done
done
cat gpurun_out/variants.txt
