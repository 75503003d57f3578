// deposit_mma.cuh -- charge-conserving mode deposit (particles.F90:584-665) of one warp of
// particles as a sum of outer products on the FP64 tensor pipe (DMMA.8x8x4).
//
// For the particles of one window (same staggered row cell_y2 = base_y, cell_x2 in
// base_x .. base_x+3) the deposit into the 5-row x 8-slot node window is separable
// (DOCUMENTATION eqs 98-105, particles.F90:640-662):
//
//   jx(ky, c, s) = sum_p  a_p(ky, c)            * run_p(s)        run = prefix sum of hx
//   jr(ky, c, s) = sum_p  -S_p(ky) f2_p(c)      * gx_p(s)  +  -S_p(ky) f3_p(c) * hx_p(s)
//   jt(ky, c, s) = sum_p  fjz_p(ky) wrt_p(ky,c) * gx_p(s)  +  fjz_p(ky) ym1_p(ky,c) * hx_p(s)
//
// with c running over the 2M-1 real coefficients (m = 0 real; m > 0 real, imaginary) and S
// the real, mode-independent radial recurrence S = S ratio + fjy (particles.F90:658).  That is
// a [5(2M-1)] x [n_particles] times [n_particles] x [8] matrix product per part: K = the 32
// particles of the warp, 8 DMMA k-steps of 4.  Each lane writes its U (coefficient) and V
// (x-shape) values as one COLUMN of a per-warp shared-memory tile, fragments are read back
// with LDS.128 (two k-steps per load; which particle sits on which k index is irrelevant for
// a sum over particles, so the column permutation is free), accumulators stay in registers,
// and each lane finally issues 2 REDs per 8x8 output tile.  No selects, no shuffles.
#pragma once
#include "push.cuh"

namespace cylgpu {

#define MMA_WX 8          // window slots in x: 5-point footprint + up to 3 cells of origin shift
#define MMA_PITCH 40      // doubles per tile row: 32 columns + pad (conflict-free LDS.128 fragments)
#define MMA_ROWS 24       // V tile (8 rows) + two U tiles (double buffer)
#define MMA_WARP_DOUBLES (MMA_ROWS * MMA_PITCH)

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int M>
struct MmaGeom {
  static constexpr int NC = 2 * M - 1;        // real coefficients per (row, component)
  static constexpr int R = 5 * NC;            // coefficient rows of one part
  static constexpr int T = (R + 7) / 8;       // 8-row DMMA tiles per part
};

// `active` lanes have cell_y2 == base_y and 0 <= sx = cell_x2 - base_x <= MMA_WX - 5; every lane of
// the warp must call this (inactive lanes contribute exact zeros; their D must be finite).
template <int M>
__device__ __forceinline__ void deposit_mma(const PushConst& P, const DepositIn& D, bool active, int lane,
                                            int base_x, int base_y, int sx, double* __restrict__ wbuf) {
  typedef MmaGeom<M> G;
  constexpr int NC = G::NC, R = G::R, T = G::T;
  const Geom& g = P.g;
  const double third = 1.0 / 3.0;
  const double* inv_area_rt = P.tab + JNG;
  const double* inv_area_xt = P.tab + P.ntab + JNG;
  const double* inv_volume = P.tab + 2 * P.ntab + JNG;
  const double* ratio_area_xt = P.tab + 3 * P.ntab + JNG;

  double* Vs = wbuf;
  double* Us = wbuf + 8 * MMA_PITCH;          // two tiles of 8 rows
  const int fm = lane >> 2, fk = lane & 3;    // fragment coordinates: row (A) / column (B) and k
  const int frag_off = fm * MMA_PITCH + 2 * fk;

  const double fcx = active ? D.fcx : 0.0;
  const double fcz = active ? D.fcz : 0.0;
  if (!active) sx = 0;

  // mode factors for every m > 0
  cplx f2[M > 1 ? M - 1 : 1], f3[M > 1 ? M - 1 : 1], f4[M > 1 ? M - 1 : 1];
  {
    cplx e0 = C(1.0, 0.0), ed = C(1.0, 0.0);
#pragma unroll
    for (int im = 1; im < M; ++im) {
      e0 = e0 * D.exp_itheta_05;
      ed = ed * D.exp_idtheta;
      const ModeFac mf = mode_factors(im, D.dtheta, e0, ed);
      f2[im - 1] = mf.f2; f3[im - 1] = mf.f3; f4[im - 1] = mf.f4;
    }
  }

  // V column of this particle: value k of the 5-point footprint goes to window slot sx + k, the
  // three remaining slots get zeros.  Outside [xmin, xmax] gx and hx are exact zeros already.
  auto store_v = [&](const double (&v)[5]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) Vs[((sx + k) & 7) * MMA_PITCH + lane] = (k < 5) ? v[k] : 0.0;
  };
  double bfrag[8];
  auto load_b = [&]() {
#pragma unroll
    for (int J = 0; J < 4; ++J) {
      const double2 t = *reinterpret_cast<const double2*>(Vs + frag_off + 8 * J);
      bfrag[2 * J] = t.x; bfrag[2 * J + 1] = t.y;
    }
  };
  // one 8-row tile of U: stage, read back as A fragments, 8 k-steps
  auto tile = [&](int t, const double (&u)[8], double (&acc0)[2], double (&acc1)[2]) {
    double* Ut = Us + (t & 1) * 8 * MMA_PITCH;
#pragma unroll
    for (int r = 0; r < 8; ++r) Ut[r * MMA_PITCH + lane] = u[r];
    __syncwarp();
    if (t == 0) load_b();
#pragma unroll
    for (int J = 0; J < 4; ++J) {
      const double2 a = *reinterpret_cast<const double2*>(Ut + frag_off + 8 * J);
      dmma884(acc0, a.x, bfrag[2 * J]);
      dmma884(acc1, a.y, bfrag[2 * J + 1]);
    }
  };
  // RED the accumulators of one component: lane (fm, fk) holds rows 8t + fm, slots 2fk, 2fk+1
  auto flush = [&](double* __restrict__ arr, size_t shift, double (&acc)[T][2][2]) {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int c = 8 * t + fm;
      const int ky = c / NC, coef = c - ky * NC;
      const int im = (coef + 1) >> 1, reim = coef ? ((coef + 1) & 1) : 0;
      const size_t o = g.at(base_x - 2 + 2 * fk, base_y - 2 + ky, im) + shift;
      const double v0 = acc[t][0][0] + acc[t][1][0], v1 = acc[t][0][1] + acc[t][1][1];
      if (c < R) {
        if (v0 != 0.0) atomicAdd(arr + 2 * o + reim, v0);
        if (v1 != 0.0) atomicAdd(arr + 2 * (o + 1) + reim, v1);
      }
    }
  };
  // coefficient c of a part -> (row ky, mode im, re/im), all compile-time after unrolling
#define MMA_PART(UEXPR, ACC)                                                   \
  do {                                                                         \
    _Pragma("unroll") for (int t = 0; t < T; ++t) {                            \
      double u[8];                                                             \
      _Pragma("unroll") for (int r = 0; r < 8; ++r) {                          \
        const int c = 8 * t + r;                                               \
        const int ky = (c < R) ? c / NC : 0, coef = (c < R) ? c % NC : 0;      \
        const int im = (coef + 1) >> 1;                                        \
        const bool imag = coef > 0 && ((coef + 1) & 1);                        \
        double val = 0.0;                                                      \
        if (c < R) { UEXPR; }                                                  \
        u[r] = val;                                                            \
      }                                                                        \
      tile(t, u, ACC[t][0], ACC[t][1]);                                        \
    }                                                                          \
  } while (0)

  double acc[T][2][2];
  auto zero_acc = [&]() {
#pragma unroll
    for (int t = 0; t < T; ++t) { acc[t][0][0] = acc[t][0][1] = acc[t][1][0] = acc[t][1][1] = 0.0; }
  };
  const double* gy = D.gy;
  const double* hy = D.hy;

  // ---------------- jx: a(ky, c) x run(s) ----------------
  {
    double v[5];
    double run = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) { run = run + D.hx[k]; v[k] = run; }
    if (D.xmax < 2) v[4] = 0.0;     // ix = 2 is outside the footprint (particles.F90:646 loop bound)
    __syncwarp();
    store_v(v);
    double fjx[5];
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) fjx[ky] = fcx * __ldg(&inv_area_rt[base_y - 2 + ky]);
    zero_acc();
    MMA_PART({
      if (coef == 0) val = -(fjx[ky] * (gy[ky] + 0.5 * hy[ky]));
      else {
        const cplx w_rt = f2[im - 1] * gy[ky] + f3[im - 1] * hy[ky];
        const cplx a = (-fjx[ky]) * w_rt;
        val = imag ? a.y : a.x;
      }
    }, acc);
    flush(P.jx, 1, acc);
  }
  // ---------------- jr: -S(ky) (f2 gx + f3 hx) ----------------
  {
    double S[5];
    {
      double s = 0.0;
#pragma unroll
      for (int ky = 0; ky < 5; ++ky) {
        const int cy = base_y - 2 + ky;
        s = s * __ldg(&ratio_area_xt[cy]) + (fcx * hy[ky]) * __ldg(&inv_area_xt[cy]);
        S[ky] = s;
      }
      if (D.ymax < 2) S[4] = 0.0;   // iy = 2 is outside the footprint
      if (D.ymin > -2) S[0] = 0.0;  // (already an exact zero: hy[0] == 0; kept for clarity)
    }
    __syncwarp();
    store_v(D.gx);
    zero_acc();
    MMA_PART({
      if (coef == 0) val = -S[ky];
      else { const cplx a = (-S[ky]) * f2[im - 1]; val = imag ? a.y : a.x; }
    }, acc);
    __syncwarp();
    store_v(D.hx);
    MMA_PART({
      if (coef == 0) val = -(0.5 * S[ky]);
      else { const cplx a = (-S[ky]) * f3[im - 1]; val = imag ? a.y : a.x; }
    }, acc);
    flush(P.jr, (size_t)g.SX, acc);
  }
  // ---------------- jt: fjz(ky) (w_rt gx + ym_fac_1 hx) ----------------
  {
    double fjz[5];
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) fjz[ky] = fcz * __ldg(&inv_volume[base_y - 2 + ky]);
    __syncwarp();
    store_v(D.gx);
    zero_acc();
    MMA_PART({
      if (coef == 0) val = fjz[ky] * (gy[ky] + 0.5 * hy[ky]);
      else {
        const cplx w_rt = f2[im - 1] * gy[ky] + f3[im - 1] * hy[ky];
        const cplx a = fjz[ky] * w_rt;
        val = imag ? a.y : a.x;
      }
    }, acc);
    __syncwarp();
    store_v(D.hx);
    MMA_PART({
      if (coef == 0) val = fjz[ky] * (0.5 * gy[ky] + third * hy[ky]);
      else {
        const cplx ym1 = f3[im - 1] * gy[ky] + f4[im - 1] * hy[ky];
        const cplx a = fjz[ky] * ym1;
        val = imag ? a.y : a.x;
      }
    }, acc);
    flush(P.jt, 0, acc);
  }
#undef MMA_PART
}

}  // namespace cylgpu
