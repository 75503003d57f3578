#!/bin/bash
# 1 GPU, final build: the whole GPU suite (now with the top-hat / B-spline builds, tests/test_zz9_gpu_shapes.py), then
# the bench on the builds of the other particle shapes (generic per-particle kernel) beside variant 4 of the default build
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/r2h_pytest_gpu_n1.txt
run() {  # name workload env...
  name=$1; wl=$2; shift 2
  env "$@" timeout 400 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2h_bench_$name.json 2> gpurun_out/r2h_bench_$name.err
  python - gpurun_out/r2h_bench_$name.json $name <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "value %.4e ms/step %.3f kernel %.3f ms frac %.4f"%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_launch'],d['roofline']['frac']), d['config'].get('particle_shape'))
except Exception as e:
    print(sys.argv[2], "ERR", e)
P
}
run c2_triangle_v4 thermal_2048x256_m2_ppc64 BENCH_VARIANT=4
run c2_tophat thermal_2048x256_m2_ppc64 CYL_SHAPE=tophat
run c2_bspline3 thermal_2048x256_m2_ppc64 CYL_SHAPE=bspline3
run c3_bspline3 lwfa_8192x512_m2_ppc32 CYL_SHAPE=bspline3
run c3_tophat lwfa_8192x512_m2_ppc32 CYL_SHAPE=tophat
