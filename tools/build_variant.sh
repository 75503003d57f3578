#!/bin/bash
# tools/build_variant.sh NAME "-DFOO=1 ..."  -> cylindrical_epoch_b200/libcylgpu_NAME.so (tuning experiments; CYLGPU_LIB selects it)
set -e
cd "$(dirname "$0")/../cylindrical_epoch_b200"
mkdir -p build/$1
for f in api fields bcs particles transport window_insert sdf_io driver balance; do
  extra=""; [ $f = window_insert ] && extra="-fmad=false"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $extra $2 -c csrc/$f.cu -o build/$1/$f.o &
done
wait
nvcc -shared -o libcylgpu_$1.so build/$1/*.o -ldl -lpthread
echo built libcylgpu_$1.so
