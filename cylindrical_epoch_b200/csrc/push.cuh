// push.cuh -- per-particle arithmetic of push_particles (particles.F90:296-665) shared by
// the push/deposit kernels: half-step drift, triangle shape weights (include/triangle/
// gx.inc, hx_dcell.inc), azimuthal-mode field gather (include/triangle/e_part.inc,
// b_part.inc), Boris rotation, and the charge-conserving mode deposit weights.
#pragma once
#ifndef CYL_EMUL   // tests/emul/ compiles this header for the CPU against its own shim + geom.cuh
#include "ctx.cuh"
#endif

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
  double hc_alpha;           // 0.5 * part_q * dt / part_m (Higuera-Cary, particles.F90:413)
  int deposit;
  int hc_push;               // the reference's -DHC_PUSH build
  double taylor_switch;      // |m dtheta| below which m_fac_1..4 use the series: 1.0e-4 (particles.F90:593)
};

// triangle weights (unnormalised, sum = 2): gx.inc / hx_dcell.inc
__device__ __forceinline__ void tri3(double cf, double& w0, double& w1, double& w2) {
  const double cf2 = cf * cf;
  w0 = 0.25 + cf2 + cf;
  w1 = 1.5 - 2.0 * cf2;
  w2 = 0.25 + cf2 - cf;
}

// ------------------------------------------------------------------------------------------
// Division / square root.  The reference divides and takes square roots with IEEE
// semantics; CUDA's correctly rounded FP64 div/sqrt cost a MUFU seed, ~8 dependent DFMAs, a
// range check and a slow-path call each, and the push has 6 + 15 of them per particle.  The
// kernels use branch-free Newton/Goldschmidt forms instead (all operands here are finite,
// normal and non-zero): results agree with the IEEE ones to <= 1 ulp, far inside the parity
// tolerance.  -DCYL_REFERENCE_MATH restores the correctly rounded operations.
// ------------------------------------------------------------------------------------------
#ifndef CYL_REFERENCE_MATH
__device__ __forceinline__ double rcp_nr(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);     // seed is good to ~2^-20: two Newton steps reach rounding level
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ double div_nr(double a, double b) {
  const double r = rcp_nr(b);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}
// s = sqrt(x), is = 1 / sqrt(x)  (coupled Goldschmidt iteration + one correction each)
__device__ __forceinline__ void sqrt_rsqrt(double x, double& s, double& is) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  const double d = fma(-g, g, x);
  s = fma(d, h, g);
  is = h + h;
}
#else
__device__ __forceinline__ double rcp_nr(double x) { return 1.0 / x; }
__device__ __forceinline__ double div_nr(double a, double b) { return a / b; }
__device__ __forceinline__ void sqrt_rsqrt(double x, double& s, double& is) { s = sqrt(x); is = 1.0 / s; }
#endif

// dtheta = ATAN2(z15, y15) - ATAN2(z05, y05)  (particles.F90:520-523) from ed = e^{i dtheta},
// which the push has anyway: the principal angle of ed by the arcsine series when it is small
// (|sin| < 1/16: relative error < 2^-53; virtually every particle away from the axis), one
// ATAN2 otherwise.  The reference's difference of two principal angles jumps by 2 pi when the
// particle crosses the theta = pi cut (the negative y axis); that jump enters m_fac_1..4
// through m*dtheta and is reproduced: the straight path from (y05, z05) to (y15, z15) crosses
// the cut iff z changes sign and the rotation sense matches (sign of sin dtheta).
__device__ __forceinline__ double delta_theta(cplx ed, bool up05, bool up15) {
  const double s = ed.y, s2 = s * s;
  double d;
  if (ed.x > 0.0 && fabs(s) < 0.0625) {
    double p = 143.0 / 10240.0;
    p = fma(p, s2, 231.0 / 13312.0);
    p = fma(p, s2, 63.0 / 2816.0);
    p = fma(p, s2, 35.0 / 1152.0);
    p = fma(p, s2, 5.0 / 112.0);
    p = fma(p, s2, 3.0 / 40.0);
    p = fma(p, s2, 1.0 / 6.0);
    d = fma(p * s2, s, s);
  } else {
    d = atan2(s, ed.x);
  }
  const double two_pi = 6.283185307179586476925286766559;
  if (up05 && !up15 && s > 0.0) d -= two_pi;
  else if (!up05 && up15 && s < 0.0) d += two_pi;
  return d;
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

// ------------------------------------------------------------------------------------------
// push_particles, particles.F90:296-582, split in three so that the gather can read either
// the mode arrays in HBM/L2 (gather_global) or the strip patch staged in shared memory
// (gather_patch).  All arithmetic keeps the reference's operation order.
// ------------------------------------------------------------------------------------------
struct W3 { double w0, w1, w2; };
struct Fields6 { double ex, er, et, bx, br, bt; };

// state carried from the half-step (pre) to the Boris rotation and deposit set-up (post)
struct PushMid {
  double ux, uy, uz;              // u = p / mc at time t
  cplx exp_min_itheta;            // e^{-i theta} at t + dt/2
  double theta_05;
  int cell_x1, cell_y1, cell_x2, cell_y2;
  W3 gx, gy, hx, hy;              // centred (g) and staggered (h) triangle weights
};

// particles.F90:296-388: half-step drift, r / theta, cells, shape weights
__device__ __forceinline__ void push_pre(const PushConst& P, double& part_x, double& part_y, double& part_z,
                                         double px, double py, double pz, PushMid& S, DepositIn& D) {
  S.ux = px * P.ipart_mc; S.uy = py * P.ipart_mc; S.uz = pz * P.ipart_mc;
#ifndef CYL_REFERENCE_MATH
  double gamma_rel, igamma;
  sqrt_rsqrt(S.ux * S.ux + S.uy * S.uy + S.uz * S.uz + 1.0, gamma_rel, igamma);
  const double root = P.dtco2 * igamma;
#else
  const double gamma_rel = sqrt(S.ux * S.ux + S.uy * S.uy + S.uz * S.uz + 1.0);
  const double root = P.dtco2 / gamma_rel;
#endif
  part_x = part_x + S.ux * root;
  part_y = part_y + S.uy * root;
  part_z = part_z + S.uz * root;

  const double part_x_local = part_x - P.x_grid_min_local;
#ifndef CYL_REFERENCE_MATH
  double part_r, ipart_r;
  sqrt_rsqrt(part_y * part_y + part_z * part_z, part_r, ipart_r);
  const double part_r_local = part_r - P.y_grid_min_local;
  S.exp_min_itheta = C(part_y * ipart_r, -part_z * ipart_r);
  S.theta_05 = 0.0;   // not needed: delta theta comes from e^{i dtheta} (push_post)
  // exp_itheta_05 = 1 / exp_min_itheta_05 = conj / |.|^2, and |.|^2 = 1 to rounding
  D.exp_itheta_05 = C(S.exp_min_itheta.x, -S.exp_min_itheta.y);
#else
  const double part_r = sqrt(part_y * part_y + part_z * part_z);
  const double part_r_local = part_r - P.y_grid_min_local;

  S.exp_min_itheta = C(part_y, -part_z) / part_r;
  S.theta_05 = atan2(part_z, part_y);
  // exp_itheta_05 = 1 / exp_min_itheta_05: real/complex division (Smith's algorithm, as
  // emitted by gfortran/libgcc for COMPLEX division)
  {
    const double zr = S.exp_min_itheta.x, zi = S.exp_min_itheta.y;
    if (fabs(zr) >= fabs(zi)) {
      const double r = zi / zr, den = zr + zi * r;
      D.exp_itheta_05 = C(1.0 / den, -r / den);
    } else {
      const double r = zr / zi, den = zi + zr * r;
      D.exp_itheta_05 = C(r / den, -1.0 / den);
    }
  }
#endif

  const double cell_x_r = part_x_local * P.idx;
  const double cell_y_r = part_r_local * P.idy;
  int cell_x1 = (int)floor(cell_x_r + 0.5);
  double cell_frac_x = (double)cell_x1 - cell_x_r;
  S.cell_x1 = cell_x1 + 1;
  int cell_y1 = (int)floor(cell_y_r + 0.5);
  double cell_frac_y = (double)cell_y1 - cell_y_r;
  S.cell_y1 = cell_y1 + 1;
  tri3(cell_frac_x, S.gx.w0, S.gx.w1, S.gx.w2);
  tri3(cell_frac_y, S.gy.w0, S.gy.w1, S.gy.w2);

  int cell_x2 = (int)floor(cell_x_r);
  cell_frac_x = (double)cell_x2 - cell_x_r + 0.5;
  S.cell_x2 = cell_x2 + 1;
  int cell_y2 = (int)floor(cell_y_r);
  cell_frac_y = (double)cell_y2 - cell_y_r + 0.5;
  S.cell_y2 = cell_y2 + 1;
  tri3(cell_frac_x, S.hx.w0, S.hx.w1, S.hx.w2);
  tri3(cell_frac_y, S.hy.w0, S.hy.w1, S.hy.w2);
}

// 3x3 weighted sum of a REAL array with element stride ES (mode 0: the imaginary part of the
// m = 0 gather never reaches the particle, e_part.inc:18 takes the real part of 1 * sum)
template <int ES, bool LDG>
__device__ __forceinline__ double gather9_r(const double* __restrict__ F, size_t o, size_t SX, const W3& wy,
                                            const W3& wx) {
  auto L = [&](size_t k) -> double { return LDG ? __ldg(&F[ES * k]) : F[ES * k]; };
  double s0 = wx.w0 * L(o);
  s0 = s0 + wx.w1 * L(o + 1);
  s0 = s0 + wx.w2 * L(o + 2);
  double s1 = wx.w0 * L(o + SX);
  s1 = s1 + wx.w1 * L(o + SX + 1);
  s1 = s1 + wx.w2 * L(o + SX + 2);
  double s2 = wx.w0 * L(o + 2 * SX);
  s2 = s2 + wx.w1 * L(o + 2 * SX + 1);
  s2 = s2 + wx.w2 * L(o + 2 * SX + 2);
  return (wy.w0 * s0 + wy.w1 * s1) + wy.w2 * s2;
}
template <bool LDG>
__device__ __forceinline__ cplx gather9_c(const cplx* __restrict__ F, size_t o, size_t SX, const W3& wy,
                                          const W3& wx) {
  auto L = [&](size_t k) -> cplx { return LDG ? __ldg(&F[k]) : F[k]; };
  cplx s0 = wx.w0 * L(o);
  s0 = s0 + wx.w1 * L(o + 1);
  s0 = s0 + wx.w2 * L(o + 2);
  cplx s1 = wx.w0 * L(o + SX);
  s1 = s1 + wx.w1 * L(o + SX + 1);
  s1 = s1 + wx.w2 * L(o + SX + 2);
  cplx s2 = wx.w0 * L(o + 2 * SX);
  s2 = s2 + wx.w1 * L(o + 2 * SX + 1);
  s2 = s2 + wx.w2 * L(o + 2 * SX + 2);
  return (wy.w0 * s0 + wy.w1 * s1) + wy.w2 * s2;
}
// real part of e * g  (cplx product, e_part.inc:18)
__device__ __forceinline__ double re_mul(cplx e, cplx g) { return e.x * g.x - e.y * g.y; }

// e_part.inc / b_part.inc straight from the mode arrays (through L1/L2)
template <int M>
__device__ __forceinline__ Fields6 gather_global(const PushConst& P, const PushMid& S) {
  const Geom& g = P.g;
  const size_t SX = g.SX;
  Fields6 F;
  {
    const size_t o12 = g.at(S.cell_x1 - 1, S.cell_y2 - 1, 0), o21 = g.at(S.cell_x2 - 1, S.cell_y1 - 1, 0);
    const size_t o22 = g.at(S.cell_x2 - 1, S.cell_y2 - 1, 0), o11 = g.at(S.cell_x1 - 1, S.cell_y1 - 1, 0);
    F.ex = 0.0 + gather9_r<2, true>((const double*)P.exm, o12, SX, S.hy, S.gx);
    F.er = 0.0 + gather9_r<2, true>((const double*)P.erm, o21, SX, S.gy, S.hx);
    F.et = 0.0 + gather9_r<2, true>((const double*)P.etm, o22, SX, S.hy, S.hx);
    F.bx = 0.0 + gather9_r<2, true>((const double*)P.bxm, o21, SX, S.gy, S.hx);
    F.br = 0.0 + gather9_r<2, true>((const double*)P.brm, o12, SX, S.hy, S.gx);
    F.bt = 0.0 + gather9_r<2, true>((const double*)P.btm, o11, SX, S.gy, S.gx);
  }
  cplx e = S.exp_min_itheta;
#pragma unroll
  for (int im = 1; im < M; ++im) {
    const size_t o12 = g.at(S.cell_x1 - 1, S.cell_y2 - 1, im), o21 = g.at(S.cell_x2 - 1, S.cell_y1 - 1, im);
    const size_t o22 = g.at(S.cell_x2 - 1, S.cell_y2 - 1, im), o11 = g.at(S.cell_x1 - 1, S.cell_y1 - 1, im);
    F.ex = F.ex + re_mul(e, gather9_c<true>(P.exm, o12, SX, S.hy, S.gx));
    F.er = F.er + re_mul(e, gather9_c<true>(P.erm, o21, SX, S.gy, S.hx));
    F.et = F.et + re_mul(e, gather9_c<true>(P.etm, o22, SX, S.hy, S.hx));
    F.bx = F.bx + re_mul(e, gather9_c<true>(P.bxm, o21, SX, S.gy, S.hx));
    F.br = F.br + re_mul(e, gather9_c<true>(P.brm, o12, SX, S.hy, S.gx));
    F.bt = F.bt + re_mul(e, gather9_c<true>(P.btm, o11, SX, S.gy, S.gx));
    e = e * S.exp_min_itheta;
  }
  return F;
}

// Strip patch in shared memory: PR rows x PC columns per (component, mode); mode 0 real only.
//   s0[(comp*PR + row)*PC + col]                        double
//   sm[((comp*(M-1) + im-1)*PR + row)*PC + col]         cplx
// patch (row 0, col 0) is the node (c0 - 1, row0 - 1) of the mode arrays.
#define PATCH_ROWS 4
template <int M, int PC>
__device__ __forceinline__ Fields6 gather_patch(const double* __restrict__ s0, const cplx* __restrict__ sm,
                                                const PushMid& S, int c0, int row0) {
  const int lx1 = S.cell_x1 - c0, ly1 = S.cell_y1 - row0, lx2 = S.cell_x2 - c0;   // ly2 == 0
  const int o12 = lx1, o21 = ly1 * PC + lx2, o22 = lx2, o11 = ly1 * PC + lx1;
  constexpr int CS = PATCH_ROWS * PC;
  Fields6 F;
  F.ex = 0.0 + gather9_r<1, false>(s0 + 0 * CS, o12, PC, S.hy, S.gx);
  F.er = 0.0 + gather9_r<1, false>(s0 + 1 * CS, o21, PC, S.gy, S.hx);
  F.et = 0.0 + gather9_r<1, false>(s0 + 2 * CS, o22, PC, S.hy, S.hx);
  F.bx = 0.0 + gather9_r<1, false>(s0 + 3 * CS, o21, PC, S.gy, S.hx);
  F.br = 0.0 + gather9_r<1, false>(s0 + 4 * CS, o12, PC, S.hy, S.gx);
  F.bt = 0.0 + gather9_r<1, false>(s0 + 5 * CS, o11, PC, S.gy, S.gx);
  cplx e = S.exp_min_itheta;
#pragma unroll
  for (int im = 1; im < M; ++im) {
    const cplx* b = sm + (size_t)(im - 1) * CS;
    constexpr int MS = (M - 1) * CS;
    F.ex = F.ex + re_mul(e, gather9_c<false>(b + 0 * MS, o12, PC, S.hy, S.gx));
    F.er = F.er + re_mul(e, gather9_c<false>(b + 1 * MS, o21, PC, S.gy, S.hx));
    F.et = F.et + re_mul(e, gather9_c<false>(b + 2 * MS, o22, PC, S.hy, S.hx));
    F.bx = F.bx + re_mul(e, gather9_c<false>(b + 3 * MS, o21, PC, S.gy, S.hx));
    F.br = F.br + re_mul(e, gather9_c<false>(b + 4 * MS, o12, PC, S.hy, S.gx));
    F.bt = F.bt + re_mul(e, gather9_c<false>(b + 5 * MS, o11, PC, S.gy, S.gx));
    e = e * S.exp_min_itheta;
  }
  return F;
}

// particles.F90:393-582: rotate (r,theta)->(y,z), Boris, full-step move, deposit set-up
__device__ __forceinline__ void push_post(const PushConst& P, const PushMid& S, const Fields6& F, double& part_x,
                                          double& part_y, double& part_z, double& px, double& py, double& pz,
                                          double part_weight, DepositIn& D) {
  const double c = C_LIGHT;
  const cplx emi = S.exp_min_itheta;
  const double ex_part = F.ex, bx_part = F.bx;
  const double ey_part = F.er * emi.x + F.et * emi.y;
  const double ez_part = -F.er * emi.y + F.et * emi.x;
  const double by_part = F.br * emi.x + F.bt * emi.y;
  const double bz_part = -F.br * emi.y + F.bt * emi.x;

  // Boris rotation (particles.F90:405-451)
  const double uxm = S.ux + P.cmratio * ex_part;
  const double uym = S.uy + P.cmratio * ey_part;
  const double uzm = S.uz + P.cmratio * ez_part;
  double gamma_rel, igamma;
  if (P.hc_push) {
    // Higuera-Cary gamma (particles.F90:409-421); not the default build, plain IEEE operations
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
  double root = P.ccmratio * igamma;
  const double taux = bx_part * root, tauy = by_part * root, tauz = bz_part * root;
  const double taux2 = taux * taux, tauy2 = tauy * tauy, tauz2 = tauz * tauz;
  const double tau = rcp_nr(1.0 + taux2 + tauy2 + tauz2);
  const double uxp = ((1.0 + taux2 - tauy2 - tauz2) * uxm
                      + 2.0 * ((taux * tauy + tauz) * uym + (taux * tauz - tauy) * uzm)) * tau;
  const double uyp = ((1.0 - taux2 + tauy2 - tauz2) * uym
                      + 2.0 * ((tauy * tauz + taux) * uzm + (tauy * taux - tauz) * uxm)) * tau;
  const double uzp = ((1.0 - taux2 - tauy2 + tauz2) * uzm
                      + 2.0 * ((tauz * taux + tauy) * uxm + (tauz * tauy - taux) * uym)) * tau;
  const double part_ux = uxp + P.cmratio * ex_part;
  const double part_uy = uyp + P.cmratio * ey_part;
  const double part_uz = uzp + P.cmratio * ez_part;

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
  double part_r, ipart_r;
  sqrt_rsqrt(part_y * part_y + part_z * part_z, part_r, ipart_r);
  const cplx exp_itheta_10 = C(part_y * ipart_r, part_z * ipart_r);
  const double part_vt = -part_vy * exp_itheta_10.y + part_vz * exp_itheta_10.x;

  // position at t + 1.5 dt (particles.F90:515-523); the stored position is not touched
  const double part_x_local = part_x + delta_x - P.x_grid_min_local;
  const double y15 = part_y + delta_y, z15 = part_z + delta_z;
  sqrt_rsqrt(y15 * y15 + z15 * z15, part_r, ipart_r);
  const double part_r_local = part_r - P.y_grid_min_local;
  const cplx exp_itheta_15 = C(y15 * ipart_r, z15 * ipart_r);
  D.exp_idtheta = exp_itheta_15 * emi;
#ifndef CYL_REFERENCE_MATH
  D.dtheta = delta_theta(D.exp_idtheta, emi.y <= 0.0, z15 >= 0.0);
#else
  const double theta_15 = atan2(z15, y15);
  D.dtheta = theta_15 - S.theta_05;
#endif

  const double cell_x_r = part_x_local * P.idx;
  const double cell_y_r = part_r_local * P.idy;
  int cell_x3 = (int)floor(cell_x_r);
  const double cell_frac_x = (double)cell_x3 - cell_x_r + 0.5;
  cell_x3 += 1;
  int cell_y3 = (int)floor(cell_y_r);
  const double cell_frac_y = (double)cell_y3 - cell_y_r + 0.5;
  cell_y3 += 1;
  const int dcellx = cell_x3 - S.cell_x2, dcelly = cell_y3 - S.cell_y2;

  // gx = old hx (placed at offsets -1..1), hx = new - old
  place3(D.gx, 0, S.hx.w0, S.hx.w1, S.hx.w2);
  place3(D.gy, 0, S.hy.w0, S.hy.w1, S.hy.w2);
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
  D.cell_x2 = S.cell_x2;
  D.cell_y2 = S.cell_y2;
  const double q_weight_fac = P.q_fac * part_weight;
  D.fcx = q_weight_fac * P.idt;
  D.fcz = q_weight_fac * part_vt;
}

// the whole push of one particle against the mode arrays in HBM/L2
template <int M>
__device__ __forceinline__ void push_one(const PushConst& P, double& part_x, double& part_y, double& part_z,
                                         double& px, double& py, double& pz, double part_weight, DepositIn& D) {
  PushMid S;
  push_pre(P, part_x, part_y, part_z, px, py, pz, S, D);
  const Fields6 F = gather_global<M>(P, S);
  push_post(P, S, F, part_x, part_y, part_z, px, py, pz, part_weight, D);
}

// the four theta-integrated mode factors m_fac_1..4 of particles.F90:588-626 for mode im > 0
struct ModeFac {
  cplx f2, f3, f4;
};
__device__ __forceinline__ ModeFac mode_factors(int im, double dtheta, cplx exp_imtheta0, cplx exp_imdtheta,
                                                 double taylor_switch) {
  const double third = 1.0 / 3.0, sixth = 0.5 * third;
  const double mdth = (double)im * dtheta;
  const double m2dth2 = mdth * mdth;
  ModeFac F;
  if (fabs(mdth) < taylor_switch) {
    const cplx f1 = 2.0 * exp_imtheta0;
    F.f2 = f1 * C(1.0 - sixth * m2dth2, 0.5 * mdth);
    F.f3 = f1 * C(0.5 - 0.125 * m2dth2, third * mdth);
    F.f4 = f1 * C(third - 0.1 * m2dth2, 0.25 * mdth);
  } else {
    const double inv_mdth = rcp_nr(mdth);
    const double inv_m2dth2 = inv_mdth * inv_mdth;
    const cplx f1 = (2.0 * inv_mdth) * exp_imtheta0;
    F.f2 = f1 * (C(0.0, -1.0) * (exp_imdtheta - C(1.0, 0.0)));
    F.f3 = f1 * (inv_mdth * (exp_imdtheta * C(1.0, -mdth) - C(1.0, 0.0)));
    F.f4 = f1 * (C(0.0, inv_m2dth2) * (exp_imdtheta * C(2.0 - m2dth2, -2.0 * mdth) - C(2.0, 0.0)));
  }
  return F;
}

}  // namespace cylgpu
