// push.cuh -- per-particle arithmetic of push_particles (particles.F90:296-665) shared by
// the push/deposit kernels: half-step drift, triangle shape weights (include/triangle/
// gx.inc, hx_dcell.inc), azimuthal-mode field gather (include/triangle/e_part.inc,
// b_part.inc), Boris rotation, and the charge-conserving mode deposit weights.
#pragma once
#include "ctx.cuh"

namespace cylgpu {

struct PushConst {
  Geom g;
  const cplx *exm, *erm, *etm, *bxm, *brm, *btm;
  double *jx, *jr, *jt;      // interleaved re/im view of jxm, jrm, jtm
  const double* tab;         // [4][ntab]: inv_area_rt, inv_area_xt, inv_volume, ratio_area_xt
  int ntab;
  double x_grid_min_local, y_grid_min_local;
  double idx, idy, idt, dtco2;
  double cmratio, ccmratio, part_mc, ipart_mc;
  double q_fac;              // part_q * fac  (q_weight_fac = q_fac * weight)
  int deposit;
};

// triangle weights (unnormalised, sum = 2): gx.inc / hx_dcell.inc
__device__ __forceinline__ void tri3(double cf, double& w0, double& w1, double& w2) {
  const double cf2 = cf * cf;
  w0 = 0.25 + cf2 + cf;
  w1 = 1.5 - 2.0 * cf2;
  w2 = 0.25 + cf2 - cf;
}

// 3x3 weighted complex sum with the reference's association order (e_part.inc:7-16)
__device__ __forceinline__ cplx gather9(const cplx* __restrict__ F, size_t o, size_t SX, double wy0, double wy1,
                                        double wy2, double wx0, double wx1, double wx2) {
  cplx s0 = wx0 * __ldg(&F[o]);
  s0 = s0 + wx1 * __ldg(&F[o + 1]);
  s0 = s0 + wx2 * __ldg(&F[o + 2]);
  cplx s1 = wx0 * __ldg(&F[o + SX]);
  s1 = s1 + wx1 * __ldg(&F[o + SX + 1]);
  s1 = s1 + wx2 * __ldg(&F[o + SX + 2]);
  cplx s2 = wx0 * __ldg(&F[o + 2 * SX]);
  s2 = s2 + wx1 * __ldg(&F[o + 2 * SX + 1]);
  s2 = s2 + wx2 * __ldg(&F[o + 2 * SX + 2]);
  return (wy0 * s0 + wy1 * s1) + wy2 * s2;
}

// Everything the deposit needs from the push of one particle
struct DepositIn {
  double gx[5], gy[5], hx[5], hy[5];   // index k <-> offset k-2; h = new - old weights
  int cell_x2, cell_y2;
  int xmin, xmax, ymin, ymax;
  double dtheta;
  cplx exp_itheta_05, exp_idtheta;
  double fcx, fcz;                      // fcy == fcx
};

// place three weights at offsets d-1, d, d+1 (d = dcell in {-1,0,1}) of a 5-vector
__device__ __forceinline__ void place3(double* v, int d, double w0, double w1, double w2) {
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int r = k - 1 - d;   // 0,1,2 -> w0,w1,w2
    v[k] = (r == 0) ? w0 : (r == 1) ? w1 : (r == 2) ? w2 : 0.0;
  }
}

// particles.F90:296-582.  Advances (x,y,z,px,py,pz) by one step and fills `D`.
__device__ __forceinline__ void push_one(const PushConst& P, double& part_x, double& part_y, double& part_z,
                                         double& px, double& py, double& pz, double part_weight, DepositIn& D) {
  const Geom& g = P.g;
  const double c = C_LIGHT;
  double part_ux = px * P.ipart_mc, part_uy = py * P.ipart_mc, part_uz = pz * P.ipart_mc;

  double gamma_rel = sqrt(part_ux * part_ux + part_uy * part_uy + part_uz * part_uz + 1.0);
  double root = P.dtco2 / gamma_rel;
  part_x = part_x + part_ux * root;
  part_y = part_y + part_uy * root;
  part_z = part_z + part_uz * root;

  double part_x_local = part_x - P.x_grid_min_local;
  double part_r = sqrt(part_y * part_y + part_z * part_z);
  double part_r_local = part_r - P.y_grid_min_local;

  const cplx exp_min_itheta = C(part_y, -part_z) / part_r;
  const double theta_05 = atan2(part_z, part_y);
  // exp_itheta_05 = 1 / exp_min_itheta_05: real/complex division (Smith's algorithm, as
  // emitted by gfortran/libgcc for COMPLEX division)
  {
    const double zr = exp_min_itheta.x, zi = exp_min_itheta.y;
    if (fabs(zr) >= fabs(zi)) {
      const double r = zi / zr, den = zr + zi * r;
      D.exp_itheta_05 = C(1.0 / den, -r / den);
    } else {
      const double r = zr / zi, den = zi + zr * r;
      D.exp_itheta_05 = C(r / den, -1.0 / den);
    }
  }

  double cell_x_r = part_x_local * P.idx;
  double cell_y_r = part_r_local * P.idy;
  int cell_x1 = (int)floor(cell_x_r + 0.5);
  double cell_frac_x = (double)cell_x1 - cell_x_r;
  cell_x1 += 1;
  int cell_y1 = (int)floor(cell_y_r + 0.5);
  double cell_frac_y = (double)cell_y1 - cell_y_r;
  cell_y1 += 1;

  double gx0, gx1, gx2, gy0, gy1, gy2;
  tri3(cell_frac_x, gx0, gx1, gx2);
  tri3(cell_frac_y, gy0, gy1, gy2);

  int cell_x2 = (int)floor(cell_x_r);
  cell_frac_x = (double)cell_x2 - cell_x_r + 0.5;
  cell_x2 += 1;
  int cell_y2 = (int)floor(cell_y_r);
  cell_frac_y = (double)cell_y2 - cell_y_r + 0.5;
  cell_y2 += 1;

  double hx0, hx1, hx2, hy0, hy1, hy2;
  tri3(cell_frac_x, hx0, hx1, hx2);
  tri3(cell_frac_y, hy0, hy1, hy2);

  // gather (e_part.inc / b_part.inc)
  const size_t SX = g.SX;
  double ex_part = 0, er_part = 0, et_part = 0, bx_part = 0, br_part = 0, bt_part = 0;
  {
    cplx e = C(1.0, 0.0);
    for (int im = 0; im < g.M; ++im) {
      const size_t o12 = g.at(cell_x1 - 1, cell_y2 - 1, im);
      const size_t o21 = g.at(cell_x2 - 1, cell_y1 - 1, im);
      const size_t o22 = g.at(cell_x2 - 1, cell_y2 - 1, im);
      const size_t o11 = g.at(cell_x1 - 1, cell_y1 - 1, im);
      ex_part = ex_part + (e * gather9(P.exm, o12, SX, hy0, hy1, hy2, gx0, gx1, gx2)).x;
      er_part = er_part + (e * gather9(P.erm, o21, SX, gy0, gy1, gy2, hx0, hx1, hx2)).x;
      et_part = et_part + (e * gather9(P.etm, o22, SX, hy0, hy1, hy2, hx0, hx1, hx2)).x;
      bx_part = bx_part + (e * gather9(P.bxm, o21, SX, gy0, gy1, gy2, hx0, hx1, hx2)).x;
      br_part = br_part + (e * gather9(P.brm, o12, SX, hy0, hy1, hy2, gx0, gx1, gx2)).x;
      bt_part = bt_part + (e * gather9(P.btm, o11, SX, gy0, gy1, gy2, gx0, gx1, gx2)).x;
      e = e * exp_min_itheta;
    }
  }
  const double ey_part = er_part * exp_min_itheta.x + et_part * exp_min_itheta.y;
  const double ez_part = -er_part * exp_min_itheta.y + et_part * exp_min_itheta.x;
  const double by_part = br_part * exp_min_itheta.x + bt_part * exp_min_itheta.y;
  const double bz_part = -br_part * exp_min_itheta.y + bt_part * exp_min_itheta.x;

  // Boris rotation (particles.F90:405-451)
  const double uxm = part_ux + P.cmratio * ex_part;
  const double uym = part_uy + P.cmratio * ey_part;
  const double uzm = part_uz + P.cmratio * ez_part;
  gamma_rel = sqrt(uxm * uxm + uym * uym + uzm * uzm + 1.0);
  root = P.ccmratio / gamma_rel;
  const double taux = bx_part * root, tauy = by_part * root, tauz = bz_part * root;
  const double taux2 = taux * taux, tauy2 = tauy * tauy, tauz2 = tauz * tauz;
  const double tau = 1.0 / (1.0 + taux2 + tauy2 + tauz2);
  const double uxp = ((1.0 + taux2 - tauy2 - tauz2) * uxm
                      + 2.0 * ((taux * tauy + tauz) * uym + (taux * tauz - tauy) * uzm)) * tau;
  const double uyp = ((1.0 - taux2 + tauy2 - tauz2) * uym
                      + 2.0 * ((tauy * tauz + taux) * uzm + (tauy * taux - tauz) * uxm)) * tau;
  const double uzp = ((1.0 - taux2 - tauy2 + tauz2) * uzm
                      + 2.0 * ((tauz * taux + tauy) * uxm + (tauz * tauy - taux) * uym)) * tau;
  part_ux = uxp + P.cmratio * ex_part;
  part_uy = uyp + P.cmratio * ey_part;
  part_uz = uzp + P.cmratio * ez_part;

  const double part_u2 = part_ux * part_ux + part_uy * part_uy + part_uz * part_uz;
  gamma_rel = sqrt(part_u2 + 1.0);
  const double igamma = 1.0 / gamma_rel;
  root = P.dtco2 * igamma;
  const double delta_x = part_ux * root, delta_y = part_uy * root, delta_z = part_uz * root;
  part_x = part_x + delta_x;
  part_y = part_y + delta_y;
  part_z = part_z + delta_z;
  px = P.part_mc * part_ux;
  py = P.part_mc * part_uy;
  pz = P.part_mc * part_uz;

  if (!P.deposit) return;

  const double part_vy = part_uy * c * igamma;
  const double part_vz = part_uz * c * igamma;
  part_r = sqrt(part_y * part_y + part_z * part_z);
  const cplx exp_itheta_10 = C(part_y, part_z) / part_r;
  const double part_vt = -part_vy * exp_itheta_10.y + part_vz * exp_itheta_10.x;

  // position at t + 1.5 dt (particles.F90:515-523); the stored position is not touched
  part_x_local = part_x + delta_x - P.x_grid_min_local;
  const double y15 = part_y + delta_y, z15 = part_z + delta_z;
  part_r = sqrt(y15 * y15 + z15 * z15);
  part_r_local = part_r - P.y_grid_min_local;
  const cplx exp_itheta_15 = C(y15, z15) / part_r;
  const double theta_15 = atan2(z15, y15);
  D.exp_idtheta = exp_itheta_15 * exp_min_itheta;
  D.dtheta = theta_15 - theta_05;

  cell_x_r = part_x_local * P.idx;
  cell_y_r = part_r_local * P.idy;
  int cell_x3 = (int)floor(cell_x_r);
  cell_frac_x = (double)cell_x3 - cell_x_r + 0.5;
  cell_x3 += 1;
  int cell_y3 = (int)floor(cell_y_r);
  cell_frac_y = (double)cell_y3 - cell_y_r + 0.5;
  cell_y3 += 1;
  const int dcellx = cell_x3 - cell_x2, dcelly = cell_y3 - cell_y2;

  // gx = old hx (placed at offsets -1..1), hx = new - old
  place3(D.gx, 0, hx0, hx1, hx2);
  place3(D.gy, 0, hy0, hy1, hy2);
  double n0, n1, n2;
  tri3(cell_frac_x, n0, n1, n2);
  place3(D.hx, dcellx, n0, n1, n2);
  tri3(cell_frac_y, n0, n1, n2);
  place3(D.hy, dcelly, n0, n1, n2);
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    D.hx[k] = D.hx[k] - D.gx[k];
    D.hy[k] = D.hy[k] - D.gy[k];
  }
  // sf_min = -1, sf_max = 1; Fortran integer division truncates toward zero like C
  D.xmin = -1 + (dcellx - 1) / 2;
  D.xmax = 1 + (dcellx + 1) / 2;
  D.ymin = -1 + (dcelly - 1) / 2;
  D.ymax = 1 + (dcelly + 1) / 2;
  D.cell_x2 = cell_x2;
  D.cell_y2 = cell_y2;
  const double q_weight_fac = P.q_fac * part_weight;
  D.fcx = q_weight_fac * P.idt;
  D.fcz = q_weight_fac * part_vt;
}

// the four theta-integrated mode factors m_fac_1..4 of particles.F90:588-626 for mode im > 0
struct ModeFac {
  cplx f2, f3, f4;
};
__device__ __forceinline__ ModeFac mode_factors(int im, double dtheta, cplx exp_imtheta0, cplx exp_imdtheta) {
  const double third = 1.0 / 3.0, sixth = 0.5 * third;
  const double mdth = (double)im * dtheta;
  const double m2dth2 = mdth * mdth;
  ModeFac F;
  if (fabs(mdth) < 1.0e-4) {
    const cplx f1 = 2.0 * exp_imtheta0;
    F.f2 = f1 * C(1.0 - sixth * m2dth2, 0.5 * mdth);
    F.f3 = f1 * C(0.5 - 0.125 * m2dth2, third * mdth);
    F.f4 = f1 * C(third - 0.1 * m2dth2, 0.25 * mdth);
  } else {
    const double inv_mdth = 1.0 / mdth;
    const double inv_m2dth2 = inv_mdth * inv_mdth;
    const cplx f1 = (2.0 * inv_mdth) * exp_imtheta0;
    F.f2 = f1 * (C(0.0, -1.0) * (exp_imdtheta - C(1.0, 0.0)));
    F.f3 = f1 * (inv_mdth * (exp_imdtheta * C(1.0, -mdth) - C(1.0, 0.0)));
    F.f4 = f1 * (C(0.0, inv_m2dth2) * (exp_imdtheta * C(2.0 - m2dth2, -2.0 * mdth) - C(2.0, 0.0)));
  }
  return F;
}

}  // namespace cylgpu
