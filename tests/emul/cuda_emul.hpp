// cuda_emul.hpp -- TEST INFRASTRUCTURE: a minimal CPU stand-in for the CUDA execution model, enough
// to compile the product's barrier-free kernel headers (csrc/bc_kernels.cuh, moments_kernels.cuh,
// insert_kernel.cuh, philox.cuh) with g++ and run them thread by thread.  There is no GPU in the
// development container: this lets the arithmetic and indexing of kernels written there be checked
// against the oracle before their first run on a B200.  Not a CPU fallback: nothing in the product
// uses it, and kernels with barriers, shuffles or tensor ops (the push) cannot be run this way.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>

struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;

// read-only cache load: a plain load
template <class T> inline T __ldg(const T* p) { return *p; }

// one emulated thread at a time: a plain read-modify-write is the atomic
inline double atomicAdd(double* p, double v) { const double old = *p; *p = old + v; return old; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  const unsigned long long old = *p;
  *p = old + v;
  return old;
}

template <class Kernel, class... Args>
void emul_launch(Kernel kernel, dim3 grid, dim3 block, Args... args) {
  gridDim = grid;
  blockDim = block;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
        for (unsigned tz = 0; tz < block.z; ++tz)
          for (unsigned ty = 0; ty < block.y; ++ty)
            for (unsigned tx = 0; tx < block.x; ++tx) {
              threadIdx.x = tx; threadIdx.y = ty; threadIdx.z = tz;
              kernel(args...);
            }
      }
}
