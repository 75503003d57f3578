// field_kernels.cuh -- the per-mode FDTD kernels (update_e_field fields.f90:53-182, update_b_field :186-312:
// bulk stencils, the r = 0 rows and the mirror rows below the axis).  Kernel-only header (host side:
// fields.cu), included inside namespace cylgpu; also compiled for the CPU by the kernel-emulation tests
// (tests/emul/).  Product code: no oracle here.
#pragma once

// Divisions: the reference divides by dx, dy, epsilon0 and the radii at every point (17 IEEE divisions per E
// point, 12 per B point); an FP64 division is ~25 instructions, which made these streaming kernels as much
// FP64-bound as memory-bound (43 % of the HBM rate under ncu).  Here each thread forms 1 / r_d and 1 / r_p once
// (correctly rounded) and multiplies by them and by the reciprocals of dx, dy, epsilon0 passed in: every factor
// stays within 1 ulp of the reference's quotient, far inside the 1e-12 the field solver is held to.
struct FieldRecips {
  double idx, idy, ieps0;
};

// E bulk: fields.f90:67-108.  ix = ix_lo..ix_hi (the reference's 0..nx, or the wider range of
// field_ranges.cuh), ir = 1..ny (y_min_boundary is always true for x-slab decomposition), all modes.
__global__ void __launch_bounds__(128) k_update_e_bulk(
    Geom g, cplx* __restrict__ exm, cplx* __restrict__ erm, cplx* __restrict__ etm,
    const cplx* __restrict__ bxm, const cplx* __restrict__ brm, const cplx* __restrict__ btm,
    const cplx* __restrict__ jxm, const cplx* __restrict__ jrm, const cplx* __restrict__ jtm,
    FieldRecips R, double dy, double dt, double y_grid_min_local, int ix_lo, int ix_hi) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x + ix_lo;
  const int ir = blockIdx.y + 1;
  const int im = blockIdx.z;
  if (ix > ix_hi) return;
  const double c = C_LIGHT;
  const double c2 = c * c;
  const double r_d = fabs((double)(ir - 1) * dy + y_grid_min_local);
  const double r_p = r_d + 0.5 * dy;
  const double fac_x = c2 / r_p;
  const cplx im_fac_x = C(0.0, (double)im) * fac_x;
  const cplx im_fac_r = C(0.0, (double)im) * (c2 / r_d);
  const double c2_dx = c2 * R.idx, c2_dy = c2 * R.idy;

  const size_t o = g.at(ix, ir, im);
  const size_t SX = g.SX;
  const cplx bt = btm[o], bt_rp = btm[o + SX], bt_xp = btm[o + 1];
  const cplx br = brm[o], br_xp = brm[o + 1];
  const cplx bx = bxm[o], bx_rp = bxm[o + SX];

  exm[o] = exm[o] + (((fac_x * 0.5) * (bt_rp + bt) + im_fac_x * br + c2_dy * (bt_rp - bt)
                      - jxm[o] * R.ieps0) * 0.5) * dt;
  erm[o] = erm[o] + (((-im_fac_r) * bx - c2_dx * (bt_xp - bt) - jrm[o] * R.ieps0) * 0.5) * dt;
  etm[o] = etm[o] + ((c2_dx * (br_xp - br) - c2_dy * (bx_rp - bx) - jtm[o] * R.ieps0) * 0.5) * dt;
}

// E axis rows and below-axis mirror: fields.f90:116-180, over the FULL extent 1-ng..nx+ng (the
// reference uses whole-array sections here).  One thread per (column, mode, task): task 0 is
// the axis row (statement order of the reference kept inside the thread), task k = 1..ng-1 the
// mirror row ir = -k, which only reads rows >= 1 that the axis task never writes except
// erm(ix,1,m>=2) -- and the erm mirrors read rows 2..ng.
__global__ void __launch_bounds__(128) k_update_e_axis(
    Geom g, cplx* __restrict__ exm, cplx* __restrict__ erm, cplx* __restrict__ etm,
    const cplx* __restrict__ btm, const cplx* __restrict__ jxm, double dy, double dt) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x + 1 - NG;
  if (ix > g.nx + NG) return;
  const int im = blockIdx.y, task = blockIdx.z;
  const double c2 = C_LIGHT * C_LIGHT;
  // mirror parity of the modes m >= 2 (fields.f90:171-175); m = 0, 1 are written out below
  const double mode_sign = (im & 1) ? -1.0 : 1.0;
  if (task > 0) {
    const int ir = -task;
    if (im == 1) {
      etm[g.at(ix, ir, 1)] = etm[g.at(ix, -ir, 1)];
      erm[g.at(ix, ir, 1)] = erm[g.at(ix, -ir + 1, 1)];
      exm[g.at(ix, ir, 1)] = -exm[g.at(ix, -ir, 1)];
    } else if (im == 0) {
      etm[g.at(ix, ir, 0)] = -etm[g.at(ix, -ir, 0)];
      erm[g.at(ix, ir, 0)] = -erm[g.at(ix, -ir + 1, 0)];
      exm[g.at(ix, ir, 0)] = exm[g.at(ix, -ir, 0)];
    } else {
      etm[g.at(ix, ir, im)] = (-mode_sign) * etm[g.at(ix, -ir, im)];
      erm[g.at(ix, ir, im)] = (-mode_sign) * erm[g.at(ix, -ir + 1, im)];
      exm[g.at(ix, ir, im)] = mode_sign * exm[g.at(ix, -ir, im)];
    }
    return;
  }
  const size_t a0 = g.at(ix, 0, im);
  if (im == 0) {
    exm[a0] = exm[a0] + ((((4.0 * c2) / dy) * btm[g.at(ix, 1, 0)] - jxm[a0] / EPSILON0) * 0.5) * dt;
    etm[a0] = C(0.0, 0.0);
    erm[a0] = -erm[g.at(ix, 1, 0)];
  } else if (im == 1) {
    exm[a0] = C(0.0, 0.0);
    const cplx er1 = erm[g.at(ix, 1, 1)];
    // uses the OLD etm(ix,0,1), then overwrites it (statement order of fields.f90:146-149)
    erm[a0] = C(0.0, 2.0) * etm[a0] - er1;
    etm[a0] = (C(0.0, -1.0) / 8.0) * (9.0 * er1 - erm[g.at(ix, 2, 1)]);
  } else {
    exm[a0] = C(0.0, 0.0);
    etm[a0] = C(0.0, 0.0);
    erm[a0] = -erm[g.at(ix, 1, im)];
    erm[g.at(ix, 1, im)] = erm[g.at(ix, 2, im)] / 9.0;
  }
}

// B bulk: fields.f90:203-241.  ix = ix_lo..ix_hi (0..nx in the reference), ir = 1..ny-1.  SAVE_OLD: the b*_old = b* copies of
// update_eb_fields_half (fields.f90:326-328) ride on this sweep -- the old values are in registers anyway -- for
// the points it visits; the remaining rows and ghost columns are copied by k_copy_b_old_rim.
template <bool SAVE_OLD>
__global__ void __launch_bounds__(128) k_update_b_bulk(
    Geom g, cplx* __restrict__ bxm, cplx* __restrict__ brm, cplx* __restrict__ btm,
    const cplx* __restrict__ exm, const cplx* __restrict__ erm, const cplx* __restrict__ etm,
    cplx* __restrict__ bxo, cplx* __restrict__ bro, cplx* __restrict__ bto,
    FieldRecips R, double dy, double dt, double y_grid_min_local, int ix_lo, int ix_hi) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x + ix_lo;
  const int ir = blockIdx.y + 1;
  const int im = blockIdx.z;
  if (ix > ix_hi) return;
  const double r_d = fabs((double)(ir - 1) * dy + y_grid_min_local);
  const double r_p = r_d + 0.5 * dy;
  const double ir_d = 1.0 / r_d;
  const cplx im_fac_x = C(0.0, (double)im) * ir_d;
  const cplx im_fac_r = C(0.0, (double)im) * (1.0 / r_p);
  const size_t o = g.at(ix, ir, im);
  const size_t SX = g.SX;
  const cplx et = etm[o], et_rm = etm[o - SX], et_xm = etm[o - 1];
  const cplx er = erm[o], er_xm = erm[o - 1];
  const cplx ex = exm[o], ex_rm = exm[o - SX];
  const cplx bx = bxm[o], br = brm[o], bt = btm[o];
  if (SAVE_OLD) { bxo[o] = bx; bro[o] = br; bto[o] = bt; }

  bxm[o] = bx - ((im_fac_x * er + (0.5 * (et + et_rm)) * ir_d + (et - et_rm) * R.idy) * 0.5) * dt;
  brm[o] = br + ((im_fac_r * ex + (et - et_xm) * R.idx) * 0.5) * dt;
  btm[o] = bt + (((-(er - er_xm)) * R.idx + (ex - ex_rm) * R.idy) * 0.5) * dt;
}

// b*_old = b* for everything k_update_b_bulk<true> does not visit: rows outside 1..ny-1 and the ghost columns
// outside ix_lo..ix_hi.  Runs BEFORE the B sweep (the axis / mirror rows are rewritten by k_update_b_axis afterwards).
__global__ void __launch_bounds__(128) k_copy_b_old_rim(Geom g, const cplx* __restrict__ bxm, const cplx* __restrict__ brm,
                                                        const cplx* __restrict__ btm, cplx* __restrict__ bxo,
                                                        cplx* __restrict__ bro, cplx* __restrict__ bto, int ix_lo,
                                                        int ix_hi) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;   // 0..SX-1
  const int row = blockIdx.y;                              // 0..SY-1
  const int im = blockIdx.z;
  if (col >= g.SX) return;
  const int ix = col + 1 - NG, ir = row + 1 - NG;
  const bool swept = ix >= ix_lo && ix <= ix_hi && ir >= 1 && ir <= g.ny - 1;
  if (swept) return;
  const size_t o = ((size_t)im * g.SY + row) * g.SX + col;
  bxo[o] = bxm[o]; bro[o] = brm[o]; bto[o] = btm[o];
}

// B axis rows and mirror: fields.f90:249-310, one thread per (column, mode, task) as for E.  The
// m = 1 Brm(ix,0) FDTD update reads etm(ix-1,0,1), which this kernel never writes; the mirrors
// read rows >= 1 only.
__global__ void __launch_bounds__(128) k_update_b_axis(
    Geom g, cplx* __restrict__ bxm, cplx* __restrict__ brm, cplx* __restrict__ btm,
    const cplx* __restrict__ exm, const cplx* __restrict__ etm, double dx, double dy, double dt) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x + 1 - NG;
  if (ix > g.nx + NG) return;
  const int im = blockIdx.y, task = blockIdx.z;
  const double mode_sign = (im & 1) ? -1.0 : 1.0;
  if (task > 0) {
    const int ir = -task;
    if (im == 0) {
      btm[g.at(ix, ir, 0)] = -btm[g.at(ix, -ir + 1, 0)];
      brm[g.at(ix, ir, 0)] = -brm[g.at(ix, -ir, 0)];
      bxm[g.at(ix, ir, 0)] = bxm[g.at(ix, -ir + 1, 0)];
    } else if (im == 1) {
      btm[g.at(ix, ir, 1)] = btm[g.at(ix, -ir + 1, 1)];
      brm[g.at(ix, ir, 1)] = brm[g.at(ix, -ir, 1)];
      bxm[g.at(ix, ir, 1)] = -bxm[g.at(ix, -ir + 1, 1)];
    } else {
      btm[g.at(ix, ir, im)] = (-mode_sign) * btm[g.at(ix, -ir + 1, im)];
      brm[g.at(ix, ir, im)] = (-mode_sign) * brm[g.at(ix, -ir, im)];
      bxm[g.at(ix, ir, im)] = mode_sign * bxm[g.at(ix, -ir + 1, im)];
    }
    return;
  }
  const size_t a0 = g.at(ix, 0, im);
  if (im == 0) {
    brm[a0] = C(0.0, 0.0);
    bxm[a0] = bxm[g.at(ix, 1, 0)];
    btm[a0] = -btm[g.at(ix, 1, 0)];
  } else if (im == 1) {
    bxm[a0] = -bxm[g.at(ix, 1, 1)];
    if (ix >= 2 - NG) {   // fields.f90:272-274 section 2-ng:nx+ng
      brm[a0] = brm[a0] + (((C(0.0, 1.0) / dy) * exm[a0] + (etm[a0] - etm[a0 - 1]) / dx) * 0.5) * dt;
    }
    btm[a0] = C(0.0, -2.0) * brm[a0] - btm[g.at(ix, 1, 1)];
  } else {
    bxm[a0] = -bxm[g.at(ix, 1, im)];
    brm[a0] = C(0.0, 0.0);
    btm[a0] = -btm[g.at(ix, 1, im)];
  }
}
