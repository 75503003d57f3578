// push_shapes.cuh -- push_particles (particles.F90:296-665) for ANY of the reference's particle shapes, one thread
// per particle: half-step drift, <shape>/gx.inc and hx_dcell.inc weights, <shape>/e_part.inc and b_part.inc gather
// through L1/L2, Boris (or Higuera-Cary), full-step move, the charge-conserving mode deposit with sf_min..sf_max of
// the shape and FP64 reductions at L2.  The loops run over the shape's own support (2, 3 or 5 nodes a direction).
// The default (triangle) build has its specialised kernels (push.cuh, push_v0.cuh, the strips of particles.cu) and
// uses this one only where a test asks for it (push variant 4); the top-hat and B-spline builds (-DCYL_SHAPE=1 / 2,
// shape.cuh) run it always.  Barrier-free: tests/emul/ runs it on the CPU.  Kernel-only header, included inside
// namespace cylgpu after push.cuh.  Product code: no oracle here.
#pragma once

struct DepositG {
  double gx[NWT], gy[NWT], hx[NWT], hy[NWT];   // index k <-> offset k - WO; h = new - old weights
  int cell_x2, cell_y2;
  int xmin, xmax, ymin, ymax;
  double dtheta;
  cplx exp_itheta_05, exp_idtheta;
  double fcx, fcz;
};

// <shape>/e_part.inc, b_part.inc: rows in r weighted by wy, columns in x by wx, summed in the includes' order
__device__ __forceinline__ cplx gather_g(const cplx* __restrict__ F, const Geom& g, const double* wy, const double* wx,
                                         int cx, int cy, int im) {
  cplx total = C(0.0, 0.0);
#pragma unroll
  for (int iy = SF_MIN; iy <= SF_MAX; ++iy) {
    const size_t o = g.at(cx, cy + iy, im);
    cplx s = wx[SF_MIN + WO] * __ldg(&F[o + SF_MIN]);
#pragma unroll
    for (int ix = SF_MIN + 1; ix <= SF_MAX; ++ix) s = s + wx[ix + WO] * __ldg(&F[o + ix]);
    const cplx row = wy[iy + WO] * s;
    total = (iy == SF_MIN) ? row : total + row;
  }
  return total;
}

template <int M>
__device__ __forceinline__ void push_one_g(const PushConst& P, double& part_x, double& part_y, double& part_z,
                                           double& px, double& py, double& pz, double part_weight, DepositG& D) {
  const Geom& g = P.g;
  const double c = C_LIGHT;
  double part_ux = px * P.ipart_mc, part_uy = py * P.ipart_mc, part_uz = pz * P.ipart_mc;
  double gamma_rel, igamma;
  sqrt_rsqrt(part_ux * part_ux + part_uy * part_uy + part_uz * part_uz + 1.0, gamma_rel, igamma);
  double root = P.dtco2 * igamma;
  part_x = part_x + part_ux * root;
  part_y = part_y + part_uy * root;
  part_z = part_z + part_uz * root;

  double part_x_local = part_x - P.x_grid_min_local;
  double part_r, ipart_r;
  sqrt_rsqrt(part_y * part_y + part_z * part_z, part_r, ipart_r);
  double part_r_local = part_r - P.y_grid_min_local;
  const cplx emi = C(part_y * ipart_r, -part_z * ipart_r);   // e^{-i theta} at t + dt/2
  D.exp_itheta_05 = C(emi.x, -emi.y);

  double cell_x_r = part_x_local * P.idx - SHAPE_CELL_SHIFT;
  double cell_y_r = part_r_local * P.idy - SHAPE_CELL_SHIFT;
  int cell_x1 = (int)floor(cell_x_r + 0.5);
  double cell_frac_x = (double)cell_x1 - cell_x_r;
  cell_x1 += 1;
  int cell_y1 = (int)floor(cell_y_r + 0.5);
  double cell_frac_y = (double)cell_y1 - cell_y_r;
  cell_y1 += 1;
  double gx[NWT], gy[NWT], hx[NWT], hy[NWT];
#pragma unroll
  for (int k = 0; k < NWT; ++k) { gx[k] = 0.0; gy[k] = 0.0; hx[k] = 0.0; hy[k] = 0.0; }
  shape_weights(cell_frac_x, 0, gx);
  shape_weights(cell_frac_y, 0, gy);
  int cell_x2 = (int)floor(cell_x_r);
  cell_frac_x = (double)cell_x2 - cell_x_r + 0.5;
  cell_x2 += 1;
  int cell_y2 = (int)floor(cell_y_r);
  cell_frac_y = (double)cell_y2 - cell_y_r + 0.5;
  cell_y2 += 1;
  shape_weights(cell_frac_x, 0, hx);
  shape_weights(cell_frac_y, 0, hy);

  // which cell pair each B component is read at: the top-hat include pairs them differently from the other two
  // (tophat/b_part.inc: bxm at (cell_x1, cell_y2), brm at (cell_x2, cell_y1), btm at (cell_x2, cell_y2))
#if CYL_SHAPE == 1
  const int bx_cx = cell_x1, bx_cy = cell_y2, br_cx = cell_x2, br_cy = cell_y1, bt_cx = cell_x2, bt_cy = cell_y2;
#else
  const int bx_cx = cell_x2, bx_cy = cell_y1, br_cx = cell_x1, br_cy = cell_y2, bt_cx = cell_x1, bt_cy = cell_y1;
#endif
  double ex_part = 0.0, er_part = 0.0, et_part = 0.0, bx_part = 0.0, br_part = 0.0, bt_part = 0.0;
  {
    cplx e = C(1.0, 0.0);
#pragma unroll
    for (int im = 0; im < M; ++im) {
      ex_part = ex_part + re_mul(e, gather_g(P.exm, g, hy, gx, cell_x1, cell_y2, im));
      er_part = er_part + re_mul(e, gather_g(P.erm, g, gy, hx, cell_x2, cell_y1, im));
      et_part = et_part + re_mul(e, gather_g(P.etm, g, hy, hx, cell_x2, cell_y2, im));
      bx_part = bx_part + re_mul(e, gather_g(P.bxm, g, gy, hx, bx_cx, bx_cy, im));
      br_part = br_part + re_mul(e, gather_g(P.brm, g, hy, gx, br_cx, br_cy, im));
      bt_part = bt_part + re_mul(e, gather_g(P.btm, g, gy, gx, bt_cx, bt_cy, im));
      e = e * emi;
    }
  }
  const double ey_part = er_part * emi.x + et_part * emi.y;
  const double ez_part = -er_part * emi.y + et_part * emi.x;
  const double by_part = br_part * emi.x + bt_part * emi.y;
  const double bz_part = -br_part * emi.y + bt_part * emi.x;

  // Boris rotation (particles.F90:405-451)
  const double uxm = part_ux + P.cmratio * ex_part;
  const double uym = part_uy + P.cmratio * ey_part;
  const double uzm = part_uz + P.cmratio * ez_part;
  if (P.hc_push) {   // Higuera-Cary (particles.F90:409-421)
    gamma_rel = uxm * uxm + uym * uym + uzm * uzm + 1.0;
    const double beta_x = P.hc_alpha * bx_part, beta_y = P.hc_alpha * by_part, beta_z = P.hc_alpha * bz_part;
    const double beta2 = beta_x * beta_x + beta_y * beta_y + beta_z * beta_z;
    const double sigma = gamma_rel - beta2;
    const double beta_dot_u = beta_x * uxm + beta_y * uym + beta_z * uzm;
    gamma_rel = sigma + sqrt(sigma * sigma + 4.0 * (beta2 + beta_dot_u * beta_dot_u));
    gamma_rel = sqrt(0.5 * gamma_rel);
    igamma = 1.0 / gamma_rel;
  } else {
    sqrt_rsqrt(uxm * uxm + uym * uym + uzm * uzm + 1.0, gamma_rel, igamma);
  }
  root = P.ccmratio * igamma;
  const double taux = bx_part * root, tauy = by_part * root, tauz = bz_part * root;
  const double taux2 = taux * taux, tauy2 = tauy * tauy, tauz2 = tauz * tauz;
  const double tau = rcp_nr(1.0 + taux2 + tauy2 + tauz2);
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
  sqrt_rsqrt(part_u2 + 1.0, gamma_rel, igamma);
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
  sqrt_rsqrt(part_y * part_y + part_z * part_z, part_r, ipart_r);
  const cplx exp_itheta_10 = C(part_y * ipart_r, part_z * ipart_r);
  const double part_vt = -part_vy * exp_itheta_10.y + part_vz * exp_itheta_10.x;

  // position at t + 1.5 dt (particles.F90:515-523); the stored position is not touched
  part_x_local = part_x + delta_x - P.x_grid_min_local;
  const double y15 = part_y + delta_y, z15 = part_z + delta_z;
  sqrt_rsqrt(y15 * y15 + z15 * z15, part_r, ipart_r);
  part_r_local = part_r - P.y_grid_min_local;
  const cplx exp_itheta_15 = C(y15 * ipart_r, z15 * ipart_r);
  D.exp_idtheta = exp_itheta_15 * emi;
#ifndef CYL_REFERENCE_MATH
  D.dtheta = delta_theta(D.exp_idtheta, emi.y <= 0.0, z15 >= 0.0);
#else
  D.dtheta = atan2(z15, y15) - atan2(-emi.y, emi.x);
#endif

#pragma unroll
  for (int k = 0; k < NWT; ++k) { D.gx[k] = hx[k]; D.gy[k] = hy[k]; }
  cell_x_r = part_x_local * P.idx - SHAPE_CELL_SHIFT;
  cell_y_r = part_r_local * P.idy - SHAPE_CELL_SHIFT;
  int cell_x3 = (int)floor(cell_x_r);
  cell_frac_x = (double)cell_x3 - cell_x_r + 0.5;
  cell_x3 += 1;
  int cell_y3 = (int)floor(cell_y_r);
  cell_frac_y = (double)cell_y3 - cell_y_r + 0.5;
  cell_y3 += 1;
  const int dcellx = cell_x3 - cell_x2, dcelly = cell_y3 - cell_y2;
  shape_weights_placed(cell_frac_x, dcellx, D.hx);   // (every index a constant: the vectors stay in registers)
  shape_weights_placed(cell_frac_y, dcelly, D.hy);
#pragma unroll
  for (int k = 0; k < NWT; ++k) { D.hx[k] = D.hx[k] - D.gx[k]; D.hy[k] = D.hy[k] - D.gy[k]; }
  // Fortran integer division truncates toward zero like C
  D.xmin = SF_MIN + (dcellx - 1) / 2;
  D.xmax = SF_MAX + (dcellx + 1) / 2;
  D.ymin = SF_MIN + (dcelly - 1) / 2;
  D.ymax = SF_MAX + (dcelly + 1) / 2;
  D.cell_x2 = cell_x2;
  D.cell_y2 = cell_y2;
  const double q_weight_fac = P.q_fac * part_weight;
  D.fcx = q_weight_fac * P.idt;
  D.fcz = q_weight_fac * part_vt;
}

// deposit, particles.F90:584-665, one particle, straight into HBM with FP64 reductions at L2
__device__ __forceinline__ void deposit_global_g(const PushConst& P, const DepositG& D) {
  const Geom& g = P.g;
  const double third = 1.0 / 3.0;
  const double* inv_area_rt = P.tab + JNG;              // index by cy directly
  const double* inv_area_xt = P.tab + P.ntab + JNG;
  const double* inv_volume = P.tab + 2 * P.ntab + JNG;
  const double* ratio_area_xt = P.tab + 3 * P.ntab + JNG;
  cplx exp_imtheta0 = C(1.0, 0.0), exp_imdtheta = C(1.0, 0.0);
  for (int im = 0; im < g.M; ++im) {
    ModeFac mf;
    mf.f2 = mf.f3 = mf.f4 = C(0.0, 0.0);
    if (im > 0) {
      exp_imtheta0 = exp_imtheta0 * D.exp_itheta_05;
      exp_imdtheta = exp_imdtheta * D.exp_idtheta;
      mf = mode_factors(im, D.dtheta, exp_imtheta0, exp_imdtheta, P.taylor_switch);
    }
    cplx jyh[NWT];
#pragma unroll
    for (int k = 0; k < NWT; ++k) jyh[k] = C(0.0, 0.0);
    // sf_min - 1 .. sf_max + 1 with compile-time indices; the run-time range of this particle as a predicate
#pragma unroll
    for (int ky = SF_MIN - 1 + WO; ky <= SF_MAX + 1 + WO; ++ky) {
      const int iy = ky - WO;
      if (iy < D.ymin || iy > D.ymax) continue;
      const int cy = D.cell_y2 + iy;
      cplx w_rt, ym_fac_1;
      if (im == 0) {
        w_rt = C(D.gy[ky] + 0.5 * D.hy[ky], 0.0);
        ym_fac_1 = C(0.5 * D.gy[ky] + third * D.hy[ky], 0.0);
      } else {
        w_rt = mf.f2 * D.gy[ky] + mf.f3 * D.hy[ky];
        ym_fac_1 = mf.f3 * D.gy[ky] + mf.f4 * D.hy[ky];
      }
      const double fjx = D.fcx * __ldg(&inv_area_rt[cy]);
      const double fjy = D.fcx * D.hy[ky] * __ldg(&inv_area_xt[cy]);
      const double fjz = D.fcz * __ldg(&inv_volume[cy]);
      const double ratio = __ldg(&ratio_area_xt[cy]);
      cplx jxh = C(0.0, 0.0);
#pragma unroll
      for (int kx = SF_MIN - 1 + WO; kx <= SF_MAX + 1 + WO; ++kx) {
        const int ix = kx - WO;
        if (ix < D.xmin || ix > D.xmax) continue;
        const int cx = D.cell_x2 + ix;
        cplx w_xt;
        if (im == 0) w_xt = C(D.gx[kx] + 0.5 * D.hx[kx], 0.0);
        else w_xt = mf.f2 * D.gx[kx] + mf.f3 * D.hx[kx];
        const cplx w_xr = D.gx[kx] * w_rt + D.hx[kx] * ym_fac_1;
        jxh = jxh - (fjx * D.hx[kx]) * w_rt;
        jyh[kx] = jyh[kx] * ratio - fjy * w_xt;
        const cplx jzh = fjz * w_xr;
        const size_t o = g.at(cx, cy, im);
        red_add(P.jx, o + 1, jxh, im > 0);
        red_add(P.jr, o + g.SX, jyh[kx], im > 0);
        red_add(P.jt, o, jzh, im > 0);
      }
    }
  }
}

template <int M>
__global__ void __launch_bounds__(128) k_push_generic(PushConst P, double* __restrict__ x, double* __restrict__ y,
                                                      double* __restrict__ z, double* __restrict__ px,
                                                      double* __restrict__ py, double* __restrict__ pz,
                                                      const double* __restrict__ w, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double X = x[i], Y = y[i], Z = z[i], PX = px[i], PY = py[i], PZ = pz[i];
  const double W = w[i];
  DepositG D;
  push_one_g<M>(P, X, Y, Z, PX, PY, PZ, W, D);
  x[i] = X; y[i] = Y; z[i] = Z;
  px[i] = PX; py[i] = PY; pz[i] = PZ;
  if (P.deposit) deposit_global_g(P, D);
}
