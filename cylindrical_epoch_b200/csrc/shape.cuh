// shape.cuh -- the particle shape functions of the reference: triangle (default), top-hat
// (-DPARTICLE_SHAPE_TOPHAT) and third-order B-spline (-DPARTICLE_SHAPE_BSPLINE3); include/<shape>/gx.inc,
// hx_dcell.inc, gxfac.inc and include/particle_to_grid.inc.  As in the reference the shape is a compile-time
// choice of the whole library (-DCYL_SHAPE=0 / 1 / 2): it sets ng = png + 2 (geom.cuh) and so the layout of every
// array.  The default build keeps its specialised triangle code (push.cuh, the strip kernels); the other shapes
// run the generic per-particle kernel of push_shapes.cuh and use the helpers below in the moment kernels too.
// Kernel-only header, included inside namespace cylgpu after geom.cuh.  Product code: no oracle here.
#pragma once

#if CYL_SHAPE == 2
#define SF_MIN (-2)
#define SF_MAX 2
#define SHAPE_FAC ((1.0 / 24.0) * (1.0 / 24.0))   // particles.F90:145-153
#define SHAPE_CELL_SHIFT 0.0
#elif CYL_SHAPE == 1
#define SF_MIN 0
#define SF_MAX 1
#define SHAPE_FAC 1.0
#define SHAPE_CELL_SHIFT 0.5                        // particles.F90:336-342: positions count from the cell edge
#else
#define SF_MIN (-1)
#define SF_MAX 1
#define SHAPE_FAC 0.25
#define SHAPE_CELL_SHIFT 0.0
#endif
#define WO 3    // weight arrays hold offsets -3..3 (sf_min-1 : sf_max+1 of the widest shape): w[k + WO]
#define NWT 7

__host__ __device__ __forceinline__ double pow4(double x) { const double t = x * x; return t * t; }

// <shape>/gx.inc, hx_dcell.inc: UNNORMALISED weights of one direction at shift+sf_min .. shift+sf_max
__host__ __device__ __forceinline__ void shape_weights(double cf, int shift, double* w) {
#if CYL_SHAPE == 2
  const double cf2 = cf * cf;
  w[shift - 2 + WO] = pow4(0.5 + cf);
  w[shift - 1 + WO] = 4.75 + 11.0 * cf + 4.0 * cf2 * (1.5 - cf - cf2);
  w[shift + WO] = 14.375 + 6.0 * cf2 * (cf2 - 2.5);
  w[shift + 1 + WO] = 4.75 - 11.0 * cf + 4.0 * cf2 * (1.5 + cf - cf2);
  w[shift + 2 + WO] = pow4(0.5 - cf);
#elif CYL_SHAPE == 1
  w[shift + WO] = 0.5 + cf;
  w[shift + 1 + WO] = 0.5 - cf;
#else
  const double cf2 = cf * cf;
  w[shift - 1 + WO] = 0.25 + cf2 + cf;
  w[shift + WO] = 1.5 - 2.0 * cf2;
  w[shift + 1 + WO] = 0.25 + cf2 - cf;
#endif
}

// The same weights placed at shift+sf_min .. shift+sf_max of a 7-vector through selects, every index a compile-time
// constant (shift = dcell is -1, 0 or 1): the vector stays in registers where shape_weights' run-time index would
// send it to local memory.  Entries outside the support are set to zero.
#define NSUP (SF_MAX - SF_MIN + 1)
__host__ __device__ __forceinline__ void shape_weights_placed(double cf, int shift, double* w) {
  double t[NWT];
#pragma unroll
  for (int k = 0; k < NWT; ++k) t[k] = 0.0;
  shape_weights(cf, 0, t);   // t[SF_MIN + WO .. SF_MAX + WO]
#pragma unroll
  for (int k = 0; k < NWT; ++k) {
    double v = 0.0;
#pragma unroll
    for (int d = -1; d <= 1; ++d) {
      const int src = k - d;   // w[k] = t[k - shift]
      if (src >= SF_MIN + WO && src <= SF_MAX + WO) v = (shift == d) ? t[src] : v;
    }
    w[k] = v;
  }
}

// <shape>/gxfac.inc without its fold at the axis: the NORMALISED weights of particle_to_grid.inc
__host__ __device__ __forceinline__ void shape_weights_fac(double cf, double* w) {
#if CYL_SHAPE == 2
  const double third = 1.0 / 3.0;
  const double fac1 = 0.125 * third, fac2 = 0.5 * third, fac3 = 7.1875 * third;   // particle_head.inc
  const double c2 = cf * cf;
  w[-2 + WO] = fac1 * pow4(0.5 + cf);
  w[-1 + WO] = fac2 * (1.1875 + 2.75 * cf + c2 * (1.5 - cf - c2));
  w[0 + WO] = 0.25 * (fac3 + c2 * (c2 - 2.5));
  w[1 + WO] = fac2 * (1.1875 - 2.75 * cf + c2 * (1.5 + cf - c2));
  w[2 + WO] = fac1 * pow4(0.5 - cf);
#elif CYL_SHAPE == 1
  w[0 + WO] = 0.5 + cf;
  w[1 + WO] = 0.5 - cf;
#else
  const double c2 = cf * cf;
  w[-1 + WO] = 0.5 * (0.25 + c2 + cf);
  w[0 + WO] = 0.75 - c2;
  w[1 + WO] = 0.5 * (0.25 + c2 - cf);
#endif
}

// the fold of the radial weights of gxfac.inc for a particle next to the axis
__host__ __device__ __forceinline__ void shape_axis_fold(double part_r, double dy, double* gy) {
#if CYL_SHAPE == 2
  if (part_r < 2.0 * dy) {
    if (part_r < dy) {
      gy[0 + WO] = gy[0 + WO] + gy[-1 + WO];
      gy[1 + WO] = gy[1 + WO] + gy[-2 + WO];
      gy[-1 + WO] = 0.0;
      gy[-2 + WO] = 0.0;
    } else {
      gy[-1 + WO] = gy[-1 + WO] + gy[-2 + WO];
      gy[-2 + WO] = 0.0;
    }
  }
#elif CYL_SHAPE == 1
  if (part_r < 0.5 * dy) {
    gy[1 + WO] = 1.0;
    gy[0 + WO] = 0.0;
  }
#else
  if (part_r < dy) {
    gy[0 + WO] = gy[0 + WO] + gy[-1 + WO];
    gy[-1 + WO] = 0.0;
  }
#endif
}

// include/particle_to_grid.inc: nearest cell and normalised weights of a particle at (x, r) relative to the grid
__host__ __device__ __forceinline__ void shape_particle_to_grid(double x_local, double r_local, double part_r,
                                                                double dx, double dy, int* cell_x, int* cell_y,
                                                                double* gx, double* gy) {
  const double cell_x_r = x_local / dx - SHAPE_CELL_SHIFT;
  const double cell_y_r = r_local / dy - SHAPE_CELL_SHIFT;
  const int cx = (int)floor(cell_x_r + 0.5);
  const int cy = (int)floor(cell_y_r + 0.5);
  const double cfx = (double)cx - cell_x_r;
  const double cfy = (double)cy - cell_y_r;
  *cell_x = cx + 1;
  *cell_y = cy + 1;
  for (int k = 0; k < NWT; ++k) { gx[k] = 0.0; gy[k] = 0.0; }
  shape_weights_fac(cfx, gx);
  shape_weights_fac(cfy, gy);
  shape_axis_fold(part_r, dy, gy);
}
