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
#ifndef MMA_VSLOTS
#define MMA_VSLOTS 2      // resident V tiles: 2 = gx and run/hx side by side, 1 = one tile re-staged per part
#endif
#define MMA_ROWS (8 * MMA_VSLOTS + 16)   // V tiles + two U tiles of 8 rows
#define MMA_WARP_DOUBLES (MMA_ROWS * MMA_PITCH)

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
#ifdef CYL_KNOCK_DMMA   // tuning experiment: one DFMA per lane instead of the tensor op (wrong results)
  c[0] = fma(a, b, c[0]); c[1] = fma(b, a, c[1]);
  return;
#endif
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

  // per-warp staging: two V tiles (slot A: run, later hx; slot B: gx) and two U tiles
  double* VA = wbuf;
  double* VB = wbuf + (MMA_VSLOTS > 1 ? 1 : 0) * 8 * MMA_PITCH;
  double* VH = wbuf + (MMA_VSLOTS > 2 ? 2 : 0) * 8 * MMA_PITCH;   // hx tile (slot A unless 3 slots)
  double* Us = wbuf + MMA_VSLOTS * 8 * MMA_PITCH;
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
  int voff[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) voff[k] = ((sx + k) & 7) * MMA_PITCH + lane;
  auto store_v = [&](double* V, const double (&v)[5]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) V[voff[k]] = (k < 5) ? v[k] : 0.0;
  };
  // a pair of 8-row U tiles against one V tile: 8 k-steps, B fragment shared by both tiles
  auto mma_pair = [&](const double* V, bool two, double (&acc0)[2][2], double (&acc1)[2][2]) {
#pragma unroll
    for (int J = 0; J < 4; ++J) {
      const double2 b = *reinterpret_cast<const double2*>(V + frag_off + 8 * J);
      const double2 a0 = *reinterpret_cast<const double2*>(Us + frag_off + 8 * J);
      dmma884(acc0[0], a0.x, b.x);
      dmma884(acc0[1], a0.y, b.y);
      if (two) {
        const double2 a1 = *reinterpret_cast<const double2*>(Us + 8 * MMA_PITCH + frag_off + 8 * J);
        dmma884(acc1[0], a1.x, b.x);
        dmma884(acc1[1], a1.y, b.y);
      }
    }
  };
  // RED the accumulators of one component: lane (fm, fk) holds rows 8t + fm, slots 2fk, 2fk+1
  auto flush = [&](double* __restrict__ arr, size_t shift, double (&acc)[T + 1][2][2]) {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int c = 8 * t + fm;
      const int ky = c / NC, coef = c - ky * NC;
      const int im = (coef + 1) >> 1, reim = coef ? ((coef + 1) & 1) : 0;
      const size_t o = g.at(base_x - 2 + 2 * fk, base_y - 2 + ky, im) + shift;
      const double v0 = acc[t][0][0] + acc[t][1][0], v1 = acc[t][0][1] + acc[t][1][1];
#ifdef CYL_KNOCK_RED   // tuning experiment: keep the arithmetic alive, never issue the RED
      if (c < R) {
        if (v0 == 1.2345e300) atomicAdd(arr + 2 * o + reim, v0);
        if (v1 == 1.2345e300) atomicAdd(arr + 2 * (o + 1) + reim, v1);
      }
#else
      if (c < R) {
        if (v0 != 0.0) atomicAdd(arr + 2 * o + reim, v0);
        if (v1 != 0.0) atomicAdd(arr + 2 * (o + 1) + reim, v1);
      }
#endif
    }
  };
  // One part = all T tiles of a coefficient block against the V tile `VT`.  Coefficient c of a
  // part -> (row ky, mode im, re/im), all compile-time after unrolling.  Tiles go in pairs
  // (both staged before one __syncwarp); the leading __syncwarp retires every fragment read
  // of the previous pair before its buffers are overwritten.
#define MMA_PART(VT, UEXPR, ACC)                                               \
  do {                                                                         \
    _Pragma("unroll") for (int t0 = 0; t0 < T; t0 += 2) {                      \
      if (t0 > 0) __syncwarp();                                                \
      _Pragma("unroll") for (int tt = 0; tt < 2; ++tt) {                       \
        const int t = t0 + tt;                                                 \
        if (t < T) {                                                           \
          _Pragma("unroll") for (int r = 0; r < 8; ++r) {                      \
            const int c = 8 * t + r;                                           \
            const int ky = (c < R) ? c / NC : 0, coef = (c < R) ? c % NC : 0;  \
            const int im = (coef + 1) >> 1;                                    \
            const bool imag = coef > 0 && ((coef + 1) & 1);                    \
            double val = 0.0;                                                  \
            if (c < R) { UEXPR; }                                              \
            Us[(tt * 8 + r) * MMA_PITCH + lane] = val;                         \
          }                                                                    \
        }                                                                      \
      }                                                                        \
      __syncwarp();                                                            \
      mma_pair(VT, t0 + 1 < T, ACC[t0], ACC[t0 + 1]);                          \
    }                                                                          \
  } while (0)

  double acc[T + 1][2][2];   // one spare so that the odd tile of the last pair has a (dead) target
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
    store_v(VA, v);
    if (MMA_VSLOTS == 3) { store_v(VB, D.gx); store_v(VH, D.hx); }   // all x-shape tiles up front: gx, hx die here
    double fjx[5];
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) fjx[ky] = fcx * __ldg(&inv_area_rt[base_y - 2 + ky]);
    zero_acc();
    MMA_PART(VA, {
      if (coef == 0) val = -(fjx[ky] * (gy[ky] + 0.5 * hy[ky]));
      else {
        const cplx w_rt = f2[im - 1] * gy[ky] + f3[im - 1] * hy[ky];
        const cplx a = (-fjx[ky]) * w_rt;
        val = imag ? a.y : a.x;
      }
    }, acc);
    flush(P.jx, 1, acc);
  }
  // gx -> slot B, hx -> slot A (the run tile is dead after the sync)
  __syncwarp();
  if (MMA_VSLOTS < 3) store_v(VB, D.gx);
  if (MMA_VSLOTS == 2) store_v(VA, D.hx);
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
    }
    zero_acc();
    MMA_PART(VB, {
      if (coef == 0) val = -S[ky];
      else { const cplx a = (-S[ky]) * f2[im - 1]; val = imag ? a.y : a.x; }
    }, acc);
    __syncwarp();
    if (MMA_VSLOTS == 1) store_v(VA, D.hx);
    MMA_PART(VH, {
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
    zero_acc();
    __syncwarp();
    if (MMA_VSLOTS == 1) store_v(VB, D.gx);
    MMA_PART(VB, {
      if (coef == 0) val = fjz[ky] * (gy[ky] + 0.5 * hy[ky]);
      else {
        const cplx w_rt = f2[im - 1] * gy[ky] + f3[im - 1] * hy[ky];
        const cplx a = fjz[ky] * w_rt;
        val = imag ? a.y : a.x;
      }
    }, acc);
    __syncwarp();
    if (MMA_VSLOTS == 1) store_v(VA, D.hx);
    MMA_PART(VH, {
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
