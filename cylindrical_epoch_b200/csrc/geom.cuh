// geom.cuh -- what device kernels need of the array geometry: ghost widths, physical constants,
// complex helpers on double2 and the (ix, ir, im) indexing of the mode arrays.  No CUDA runtime
// dependency beyond double2 / make_double2 and the __host__ __device__ qualifiers, so that the
// kernel headers (moments_kernels.cuh, bc_kernels.cuh, insert_kernel.cuh) can also be compiled by
// a plain C++ compiler against the emulation shim of tests/emul/ (include after cuda_runtime.h
// or after that shim).  Product code: never includes, links or calls anything under oracle/.
#pragma once
#include <stddef.h>
#include <stdint.h>

// The particle shape is a compile-time choice of the reference and of this library (shape.cuh): -DCYL_SHAPE=0
// triangle (default), 1 top-hat, 2 third-order B-spline.  It sets the ghost widths, constants.F90:524-545.
#ifndef CYL_SHAPE
#define CYL_SHAPE 0
#endif
#if CYL_SHAPE == 2
#define PNG 4
#elif CYL_SHAPE == 1
#define PNG 2
#else
#define PNG 3    // constants.F90:537
#endif
#define NG (PNG + 2)   // constants.F90:544
#define JNG NG         // constants.F90:545  MAX(ng, png)
#define CELL_PAD 3   // ghost cells on each side covered by the particle sort buckets

namespace cylgpu {

// physical constants, constants.F90:171-201
constexpr double PI = 3.141592653589793238462643383279503;
constexpr double C_LIGHT = 2.99792458e8;
constexpr double EPSILON0 = 8.854187817620389850536563031710750e-12;

// ---- complex helpers on double2 (x = re, y = im) ----
typedef double2 cplx;
__host__ __device__ __forceinline__ cplx C(double re, double im) { return make_double2(re, im); }
__host__ __device__ __forceinline__ cplx operator+(cplx a, cplx b) { return C(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx operator-(cplx a, cplx b) { return C(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx operator-(cplx a) { return C(-a.x, -a.y); }
__host__ __device__ __forceinline__ cplx operator*(cplx a, cplx b) {
  return C(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx operator*(double s, cplx a) { return C(s * a.x, s * a.y); }
__host__ __device__ __forceinline__ cplx operator*(cplx a, double s) { return C(a.x * s, a.y * s); }
__host__ __device__ __forceinline__ cplx operator/(cplx a, double s) { return C(a.x / s, a.y / s); }
__host__ __device__ __forceinline__ cplx& operator+=(cplx& a, cplx b) { a.x += b.x; a.y += b.y; return a; }
// i * a
__host__ __device__ __forceinline__ cplx mul_i(cplx a) { return C(-a.y, a.x); }

// geometry of the mode arrays: (ix, ir, im) with bounds (1-ng:nx+ng, 1-ng:ny+ng, 0:M-1)
struct Geom {
  int nx, ny, M;
  int SX, SY;          // nx + 2ng, ny + 2ng
  size_t plane;        // SX * SY
  __host__ __device__ __forceinline__ size_t at(int ix, int ir, int im) const {
    return ((size_t)im * SY + (size_t)(ir + NG - 1)) * SX + (size_t)(ix + NG - 1);
  }
};

#include "shape.cuh"

// particle_bcs with device-resident counts (compact_kernels.cuh): the plan of one compaction and the per-step
// statistics kept on the device
enum { PST_SENT_L = 0, PST_SENT_R = 1, PST_REMOVED = 2, PST_RECV = 3, PST_WINDOW_REMOVED = 4, PST_OVERFLOW = 5, PST_N = 8 };
struct CompactPlan { long long nholes, n_new, n_left, n_right; };

}  // namespace cylgpu
