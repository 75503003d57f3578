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
#define MMA_ROWS 40       // three V tiles (run, gx, hx) + two U tiles, 8 rows each
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
                                            int base_x, int base_y, int sx, double* __restrict__ wbuf,
                                            const double* __restrict__ stab, int stab_row) {
  typedef MmaGeom<M> G;
  constexpr int NC = G::NC, R = G::R, T = G::T;
  const Geom& g = P.g;
  const double third = 1.0 / 3.0;
  // radial tables of the 5 window rows: the strip's copy in shared memory (stab[t*5 + ky], staged
  // for base_y == stab_row) or, for a window of another row, the global tables
  // (one generic pointer, so that only one of the two loads is ever issued)
  const bool tab_sm = (base_y == stab_row);
  const double* tab_base = tab_sm ? stab : P.tab + JNG + base_y - 2;
  const int tab_stride = tab_sm ? 5 : P.ntab;
  auto tab = [&](int t, int ky) -> double { return tab_base[t * tab_stride + ky]; };

  // per-warp staging: three V tiles (A: run, B: gx, H: hx) and two U tiles
  double* VA = wbuf;
  double* VB = wbuf + 8 * MMA_PITCH;
  double* VH = wbuf + 16 * MMA_PITCH;
  double* Us = wbuf + 24 * MMA_PITCH;
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
      const ModeFac mf = mode_factors(im, D.dtheta, e0, ed, P.taylor_switch);
      f2[im - 1] = mf.f2; f3[im - 1] = mf.f3; f4[im - 1] = mf.f4;
    }
  }

  // V column of this particle: value k of the 5-point footprint goes to window slot sx + k, the
  // three remaining slots get zeros.  Outside [xmin, xmax] gx and hx are exact zeros already.
  auto store_v = [&](double* V, const double (&v)[5]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) V[((sx + k) & 7) * MMA_PITCH + lane] = (k < 5) ? v[k] : 0.0;
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
  // RED one accumulated 8-row tile: lane (fm, fk) holds row 8t + fm, slots 2fk, 2fk+1
  auto flush_tile = [&](double* __restrict__ arr, size_t shift, int t, const double (&acc)[2][2]) {
    const int c = 8 * t + fm;
    const int ky = c / NC, coef = c - ky * NC;
    const int im = (coef + 1) >> 1, reim = coef ? ((coef + 1) & 1) : 0;
    const size_t o = g.at(base_x - 2 + 2 * fk, base_y - 2 + ky, im) + shift;
    const double v0 = acc[0][0] + acc[1][0], v1 = acc[0][1] + acc[1][1];
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
  };
  // Stage the pair of U tiles (t0, t0+1) of one coefficient block.  Coefficient c of a block
  // -> (row ky, mode im, re/im), all compile-time after unrolling.
#define MMA_STAGE(T0, UEXPR)                                                   \
  do {                                                                         \
    _Pragma("unroll") for (int tt = 0; tt < 2; ++tt) {                         \
      const int t = (T0) + tt;                                                 \
      if (t < T) {                                                             \
        _Pragma("unroll") for (int r = 0; r < 8; ++r) {                        \
          const int c = 8 * t + r;                                             \
          const int ky = (c < R) ? c / NC : 0, coef = (c < R) ? c % NC : 0;    \
          const int im = (coef + 1) >> 1;                                      \
          const bool imag = coef > 0 && ((coef + 1) & 1);                      \
          double val = 0.0;                                                    \
          /* rows >= R are padding: their accumulators are never flushed, so   \
             whatever the tile holds there is harmless and need not be stored */ \
          if (c < R) { UEXPR; Us[(tt * 8 + r) * MMA_PITCH + lane] = val; }     \
        }                                                                      \
      }                                                                        \
    }                                                                          \
  } while (0)
  // One component, tile pair by tile pair: [stage U1, mma with V1] (+ [stage U2, mma with V2])
  // into the same accumulators, then RED.  A __syncwarp before each staging retires the
  // fragment reads of the previous one, a second one publishes the new tiles; only two tiles
  // of accumulators are ever live, whatever n_mode is.
#define MMA_COMPONENT(ARR, SHIFT, V1, UEXPR1, TWO, V2, UEXPR2)                 \
  do {                                                                         \
    _Pragma("unroll") for (int t0 = 0; t0 < T; t0 += 2) {                      \
      double acc0[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, acc1[2][2] = {{0.0, 0.0}, {0.0, 0.0}}; \
      __syncwarp();                                                            \
      MMA_STAGE(t0, UEXPR1);                                                   \
      __syncwarp();                                                            \
      mma_pair(V1, t0 + 1 < T, acc0, acc1);                                    \
      if (TWO) {                                                               \
        __syncwarp();                                                          \
        MMA_STAGE(t0, UEXPR2);                                                 \
        __syncwarp();                                                          \
        mma_pair(V2, t0 + 1 < T, acc0, acc1);                                  \
      }                                                                        \
      flush_tile(ARR, SHIFT, t0, acc0);                                        \
      if (t0 + 1 < T) flush_tile(ARR, SHIFT, t0 + 1, acc1);                    \
    }                                                                          \
  } while (0)

  const double* gy = D.gy;
  const double* hy = D.hy;

  // all three x-shape tiles up front (gx, hx are dead afterwards): run = prefix sum of hx
  {
    double v[5];
    double run = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) { run = run + D.hx[k]; v[k] = run; }
    if (D.xmax < 2) v[4] = 0.0;     // ix = 2 is outside the footprint (particles.F90:646 loop bound)
    __syncwarp();
    store_v(VA, v);
    store_v(VB, D.gx);
    store_v(VH, D.hx);
  }
  // ---------------- jx: a(ky, c) x run(s) ----------------
  {
    double fjx[5];
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) fjx[ky] = fcx * tab(0, ky);
    MMA_COMPONENT(P.jx, 1, VA, {
      if (coef == 0) val = -(fjx[ky] * (gy[ky] + 0.5 * hy[ky]));
      else {
        const cplx w_rt = f2[im - 1] * gy[ky] + f3[im - 1] * hy[ky];
        const cplx a = (-fjx[ky]) * w_rt;
        val = imag ? a.y : a.x;
      }
    }, false, VA, {});
  }
  // ---------------- jr: -S(ky) (f2 gx + f3 hx) ----------------
  {
    double S[5];
    {
      double s = 0.0;
#pragma unroll
      for (int ky = 0; ky < 5; ++ky) {
        s = s * tab(3, ky) + (fcx * hy[ky]) * tab(1, ky);
        S[ky] = s;
      }
      if (D.ymax < 2) S[4] = 0.0;   // iy = 2 is outside the footprint
    }
    MMA_COMPONENT(P.jr, (size_t)g.SX, VB, {
      if (coef == 0) val = -S[ky];
      else { const cplx a = (-S[ky]) * f2[im - 1]; val = imag ? a.y : a.x; }
    }, true, VH, {
      if (coef == 0) val = -(0.5 * S[ky]);
      else { const cplx a = (-S[ky]) * f3[im - 1]; val = imag ? a.y : a.x; }
    });
  }
  // ---------------- jt: fjz(ky) (w_rt gx + ym_fac_1 hx) ----------------
  {
    double fjz[5];
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) fjz[ky] = fcz * tab(2, ky);
    MMA_COMPONENT(P.jt, 0, VB, {
      if (coef == 0) val = fjz[ky] * (gy[ky] + 0.5 * hy[ky]);
      else {
        const cplx w_rt = f2[im - 1] * gy[ky] + f3[im - 1] * hy[ky];
        const cplx a = fjz[ky] * w_rt;
        val = imag ? a.y : a.x;
      }
    }, true, VH, {
      if (coef == 0) val = fjz[ky] * (0.5 * gy[ky] + third * hy[ky]);
      else {
        const cplx ym1 = f3[im - 1] * gy[ky] + f4[im - 1] * hy[ky];
        const cplx a = fjz[ky] * ym1;
        val = imag ? a.y : a.x;
      }
    });
  }
#undef MMA_STAGE
#undef MMA_COMPONENT
}

}  // namespace cylgpu
