#!/bin/bash
# 1 GPU: generic kernel with register-resident weights: parity of the shape builds and of variant 4, then the benches
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz9_gpu_shapes.py tests/test_gpu_parity.py -m gpu -q -k "shape or 4 or variants_agree" 2>&1 | tail -3 | cut -c1-300 | tee gpurun_out/r2i_pytest_shapes.txt
run() {  # name workload env...
  name=$1; wl=$2; shift 2
  env "$@" timeout 400 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2i_bench_$name.json 2> gpurun_out/r2i_bench_$name.err
  python - gpurun_out/r2i_bench_$name.json $name <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[2], "value %.4e ms/step %.3f kernel %.3f ms frac %.4f"%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms_per_launch'],d['roofline']['frac']), d['config'].get('particle_shape'))
except Exception as e:
    print(sys.argv[2], "ERR", e)
P
}
run c2_triangle_v4 thermal_2048x256_m2_ppc64 BENCH_VARIANT=4
run c2_triangle_v0 thermal_2048x256_m2_ppc64 BENCH_VARIANT=0
run c2_tophat thermal_2048x256_m2_ppc64 CYL_SHAPE=tophat
run c2_bspline3 thermal_2048x256_m2_ppc64 CYL_SHAPE=bspline3
run c3_bspline3 lwfa_8192x512_m2_ppc32 CYL_SHAPE=bspline3
run c3_tophat lwfa_8192x512_m2_ppc32 CYL_SHAPE=tophat
