#!/bin/bash
# usage: tools/bench_variants.sh "1 2 3" [workload]   -- push-kernel comparison of variants
WL=${2:-thermal_2048x256_m2_ppc64}
for v in $1; do
  BENCH_VARIANT=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload $WL | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('variant $v $WL value %.3e  push_kernel %.3f ms  phases %s  frac %.4f drift %.2e' % (d['value'], d['roofline']['kernel_ms_per_launch'], {k: round(v,3) for k,v in d['phase_ms_per_step'].items()}, d['roofline']['frac'], d['energy']['relative_drift_over_timed_steps']))"
done
