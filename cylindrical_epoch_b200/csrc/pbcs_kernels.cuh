// pbcs_kernels.cuh -- particle_bcs (boundary.F90:1541-1889): the per-particle boundary rules shared by the
// fused strip push and the stand-alone classification kernel, and that kernel.  Kernel-only header,
// included inside namespace cylgpu; also compiled for the CPU by the kernel-emulation tests (tests/emul/),
// which check the integer outputs (leavers per direction) against the oracle.  Product code: no oracle here.
#pragma once

// ------------------------------------------------------------------------------------------
// particle_bcs, boundary.F90:1541-1889 (non-cpml, non-thermal branches), one particle.  Used by
// the stand-alone classification kernel (on the arrays) and, fused, by the strip push kernel
// (on the registers it is about to store): the deposit of a push always sees the position
// before the boundary treatment, as in the reference where particle_bcs follows the push.
// ------------------------------------------------------------------------------------------
struct BcsConst {
  double x_min, x_max, x_min_local, x_max_local, y_max;
  double x_min_outer, x_max_outer, y_max_outer, x_shift;
  double y_max2_inside;   // r^2 below this is inside y_max whatever the rounding of the square root
  int x_min_boundary, x_max_boundary;
  int bc[4];
  // the stand-alone classification after a window shift (k_pbcs_classify_dev): remove_particles (window.F90:304-325)
  // rides on it (x < remove_x leaves the list), and a list that was inside in r before the shift is only tested in x
  double remove_x;
  int x_only;
};

enum { FL_KEEP = 0, FL_LEFT = 1, FL_RIGHT = 2, FL_GONE = 3, FL_GONE_WINDOW = 4 };
enum { CNT_HOLE = 0, CNT_LEFT = 1, CNT_RIGHT = 2, CNT_GONE = 3, CNT_GONE_WINDOW = 4, CNT_LOW = 5, CNT_PACK_L = 6, CNT_PACK_R = 7 };

struct MemParticle {   // a particle in the SoA arrays: components are touched only when a rule needs them
  double *x, *y, *z, *px, *py, *pz;
  __device__ __forceinline__ double gx() const { return *x; }
  __device__ __forceinline__ double gy() const { return *y; }
  __device__ __forceinline__ double gz() const { return *z; }
  __device__ __forceinline__ double gpx() const { return *px; }
  __device__ __forceinline__ double gpy() const { return *py; }
  __device__ __forceinline__ double gpz() const { return *pz; }
  __device__ __forceinline__ void sx(double v) { *x = v; }
  __device__ __forceinline__ void sy(double v) { *y = v; }
  __device__ __forceinline__ void sz(double v) { *z = v; }
  __device__ __forceinline__ void spx(double v) { *px = v; }
  __device__ __forceinline__ void spy(double v) { *py = v; }
  __device__ __forceinline__ void spz(double v) { *pz = v; }
};
struct RegParticle {   // a particle in registers
  double &x, &y, &z, &px, &py, &pz;
  __device__ __forceinline__ double gx() const { return x; }
  __device__ __forceinline__ double gy() const { return y; }
  __device__ __forceinline__ double gz() const { return z; }
  __device__ __forceinline__ double gpx() const { return px; }
  __device__ __forceinline__ double gpy() const { return py; }
  __device__ __forceinline__ double gpz() const { return pz; }
  __device__ __forceinline__ void sx(double v) { x = v; }
  __device__ __forceinline__ void sy(double v) { y = v; }
  __device__ __forceinline__ void sz(double v) { z = v; }
  __device__ __forceinline__ void spx(double v) { px = v; }
  __device__ __forceinline__ void spy(double v) { py = v; }
  __device__ __forceinline__ void spz(double v) { pz = v; }
};

template <class A>
__device__ __forceinline__ uint8_t particle_bcs_one(const BcsConst& B, A& a) {
  int xbd = 0;
  bool out_of_bounds = false;
  double part_pos = a.gx();
  if (part_pos < B.x_min_local) {
    xbd = -1;
    int bc = -1;
    if (B.x_min_boundary) {
      xbd = 0;
      bc = B.bc[CYLGPU_BD_X_MIN];
      if (bc == CYLGPU_BC_REFLECT) {
        a.sx(2.0 * B.x_min - part_pos);
        a.spx(-a.gpx());
      } else if (bc == CYLGPU_BC_PERIODIC) {
        xbd = -1;
        a.sx(part_pos - (-1.0) * B.x_shift);
      }
    }
    if (part_pos < B.x_min_outer && bc != CYLGPU_BC_PERIODIC) out_of_bounds = true;
  }
  if (part_pos >= B.x_max_local) {
    xbd = 1;
    int bc = -1;
    if (B.x_max_boundary) {
      xbd = 0;
      bc = B.bc[CYLGPU_BD_X_MAX];
      if (bc == CYLGPU_BC_REFLECT) {
        a.sx(2.0 * B.x_max - part_pos);
        a.spx(-a.gpx());
      } else if (bc == CYLGPU_BC_PERIODIC) {
        xbd = 1;
        a.sx(part_pos - B.x_shift);
      }
    }
    if (part_pos >= B.x_max_outer && bc != CYLGPU_BC_PERIODIC) out_of_bounds = true;
  }
  const double Y = a.gy(), Z = a.gz();
  const double r2 = Y * Y + Z * Z;
  // the square root (boundary.F90:1744) only where it can matter: all but the outermost particles
  // are inside by a margin no rounding can bridge
  if (r2 >= B.y_max2_inside && (part_pos = sqrt(r2)) >= B.y_max) {
    const int bc = B.bc[CYLGPU_BD_Y_MAX];
    if (bc == CYLGPU_BC_REFLECT) {
      const double radial_reduction = 2.0 * B.y_max / part_pos - 1.0;
      const double Yn = Y * radial_reduction, Zn = Z * radial_reduction;
      a.sy(Yn);
      a.sz(Zn);
      const double inv_final_r = 1.0 / sqrt(Yn * Yn + Zn * Zn);
      const double cos_theta = Yn * inv_final_r, sin_theta = Zn * inv_final_r;
      const double PY = a.gpy(), PZ = a.gpz();
      const double part_pr = PY * cos_theta + PZ * sin_theta;
      const double part_pt = -PY * sin_theta + PZ * cos_theta;
      a.spy(-part_pr * cos_theta - part_pt * sin_theta);
      a.spz(-part_pr * sin_theta + part_pt * cos_theta);
    }
    if (part_pos >= B.y_max_outer && bc != CYLGPU_BC_PERIODIC) out_of_bounds = true;
  }
  uint8_t f = FL_KEEP;
  if (out_of_bounds) f = FL_GONE;
  else if (xbd == -1) f = FL_LEFT;
  else if (xbd == 1) f = FL_RIGHT;
  return f;
}

// a particle that leaves the list: its slot becomes a hole, its fate (left / right / gone) is kept
__device__ __forceinline__ void record_leaver(uint32_t* __restrict__ hole_list, uint8_t* __restrict__ hole_flag,
                                              unsigned long long* cnt, uint32_t i, uint8_t f) {
  const unsigned long long h = atomicAdd(&cnt[CNT_HOLE], 1ULL);
  hole_list[h] = i;
  hole_flag[h] = f;
  atomicAdd(&cnt[f], 1ULL);   // CNT_LEFT / CNT_RIGHT / CNT_GONE share the flag value
}

// the stand-alone classification pass (cylgpu_particle_bcs; the strip push does this on its registers)
__global__ void __launch_bounds__(256) k_pbcs_classify(BcsConst B, double* __restrict__ x, double* __restrict__ y,
                                                       double* __restrict__ z, double* __restrict__ px,
                                                       double* __restrict__ py, double* __restrict__ pz,
                                                       uint32_t* __restrict__ hole_list, uint8_t* __restrict__ hole_flag,
                                                       unsigned long long* cnt, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  MemParticle a{x + i, y + i, z + i, px + i, py + i, pz + i};
  const uint8_t f = particle_bcs_one(B, a);
  if (f != FL_KEEP) record_leaver(hole_list, hole_flag, cnt, (uint32_t)i, f);
}
