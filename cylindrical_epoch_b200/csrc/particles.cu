// particles.cu -- particle push + gather + charge-conserving mode deposit, particle boundary
// conditions / migration, and the cell sort of the SoA particle store.
// Replaces push_particles (particles.F90:28-734), particle_bcs (boundary.F90:1541-1889),
// partlist_sendrecv + pack/unpack_particle (partlist.F90:414-564,822-876) and
// remove_particles (window.F90:304-325).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "push.cuh"
#include "deposit_mma.cuh"

namespace cylgpu {

// ------------------------------------------------------------------------------------------
// per-radius tables, particles.F90:190-217 (computed on the host in the reference's order,
// r_low accumulated by repeated addition of dy)
// ------------------------------------------------------------------------------------------
int build_tables(cylgpu_ctx* c) {
  const int ny = c->g.ny;
  const int ntab = ny + 2 * JNG + 1;   // index iy in [0-jng, ny+jng] -> iy + JNG
  const double dx = c->cfg.dx, dy = c->cfg.dy;
  std::vector<double> t(4 * (size_t)ntab, 0.0);
  double* rt = t.data();
  double* xt = rt + ntab;
  double* vol = xt + ntab;
  double* ratio = vol + ntab;
  double r_low = c->cfg.y_grid_min_local - (double)JNG * dy;
  xt[0] = 1.0 / (2.0 * PI * std::fabs(r_low) * dx);
  for (int iy = 1 - JNG; iy <= ny + JNG; ++iy) {
    if (std::lround(2.0 * r_low / dy) == -1) {
      rt[iy + JNG] = 1.0 / (PI * ((0.5 * dy) * (0.5 * dy)));
    } else {
      rt[iy + JNG] = 1.0 / (PI * std::fabs((r_low + dy) * (r_low + dy) - r_low * r_low));
    }
    xt[iy + JNG] = 1.0 / (2.0 * PI * std::fabs(r_low + dy) * dx);
    r_low = r_low + dy;
  }
  for (int iy = 1 - JNG; iy <= ny + JNG; ++iy) {
    vol[iy + JNG] = rt[iy + JNG] / dx;
    ratio[iy + JNG] = xt[iy + JNG] / xt[iy + JNG - 1];
  }
  c->ntab = ntab;
  if (!c->tables) CUDA_TRY(cudaMalloc(&c->tables, 4 * (size_t)ntab * sizeof(double)));
  CUDA_TRY(cudaMemcpy(c->tables, t.data(), 4 * (size_t)ntab * sizeof(double), cudaMemcpyHostToDevice));
  return 0;
}

#include "push_v0.cuh"
#include "push_shapes.cuh"

// ------------------------------------------------------------------------------------------
// variant 1: warp-window deposit.
//
// FP64 atomics are the bottleneck of variant 0: ~144 RED.F64 per particle at ~1 RED per
// clock per SM (measured: 8.1 of 9.9 ms).  Shared-memory FP atomics are CAS loops on
// sm_100a (ATOMS.CAST.SPIN) and slower still, so the reduction is done in REGISTERS:
// particles are sorted by the staggered cell (cell_x2, cell_y2) of their upcoming
// half-step position, so the 32 lanes of a warp deposit into (nearly) the same 5-row x
// WX-column window of nodes.  Per (row, mode) every lane evaluates its contribution to each
// window slot (zero outside its own 5-point footprint), the warp runs a shuffle
// reduce-scatter (31 exchanges per 32 values, after which lane j holds the warp total of
// value j) and each lane issues ONE RED per 32 values: ~10 REDs per particle instead of
// 144.  Lanes whose footprint does not fit the warp window (sort drift, row ends) fall back
// to per-particle REDs, so correctness never depends on the sort.
// ------------------------------------------------------------------------------------------
#define WX 7   // window width in x: 5-point footprint + up to 2 cells of base shift
#ifndef PUSH_MINB
#define PUSH_MINB 4   // CTAs of 128 threads per SM the push kernel is compiled for (128 registers)
#endif

// One exchange of the reduce-scatter: lanes with bit `bit` set keep `b`, the others keep `a`;
// each sends the one it does not keep to its partner and adds what it receives.
__device__ __forceinline__ double rs_exchange(double a, double b, bool hi, int bit) {
  const double send = hi ? a : b;
  const double keep = hi ? b : a;
  return keep + __shfl_xor_sync(0xffffffffu, send, bit);
}

// Levels 2..5 of the 64-value reduce-scatter on the 32 level-1 results.  On return w[0], w[1]
// hold the warp totals of stream positions  p_i = i + 2 b0 + 4 b1 + 8 b2 + 16 b3 + 32 b4
// (b_k = bit k of the lane id), i = 0, 1.
__device__ __forceinline__ void warp_reduce_scatter_tail(double (&w)[32], int lane) {
  const bool h8 = lane & 8, h4 = lane & 4, h2 = lane & 2, h1 = lane & 1;
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = rs_exchange(w[i], w[i + 16], h8, 8);
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = rs_exchange(w[i], w[i + 8], h4, 4);
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = rs_exchange(w[i], w[i + 4], h2, 2);
#pragma unroll
  for (int i = 0; i < 2; ++i) w[i] = rs_exchange(w[i], w[i + 2], h1, 1);
}

// Value stream of one window row, in generation order g:
//   [m=0: jx(WX) jr(WX) jt(WX) real] [m=1: jx(WX re,im) jr(...) jt(...)] [m=2: ...] ...
// decoded to (mode, comp, slot, reim); returns false for padding
__device__ __forceinline__ bool decode_row_element(int e, int M, int& im, int& comp, int& slot, int& reim) {
  if (e < 3 * WX) { im = 0; comp = e / WX; slot = e % WX; reim = 0; return true; }
  const int q = e - 3 * WX;
  im = 1 + q / (6 * WX);
  if (im >= M) return false;
  const int r = q % (6 * WX);
  comp = r / (2 * WX);
  slot = (r % (2 * WX)) >> 1;
  reim = r & 1;
  return true;
}

template <int M>
__device__ __forceinline__ void deposit_window(const PushConst& P, const DepositIn& D, bool active, int lane,
                                               int base_x, int base_y, int sx) {
  const Geom& g = P.g;
  const double third = 1.0 / 3.0;
  const double* inv_area_rt = P.tab + JNG;
  const double* inv_area_xt = P.tab + P.ntab + JNG;
  const double* inv_volume = P.tab + 2 * P.ntab + JNG;
  const double* ratio_area_xt = P.tab + 3 * P.ntab + JNG;

  // x factors in the window frame (slot s <-> cx = base_x - 2 + s); zero outside the footprint
  double gxw[WX], hxw[WX];
  const int slo = sx + D.xmin + 2, shi = sx + D.xmax + 2;
#pragma unroll
  for (int s = 0; s < WX; ++s) {
    double gv = 0.0, hv = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (s - k >= 0 && s - k <= WX - 5) {   // compile-time feasible shifts only
        const bool sel = (sx == s - k);
        gv = sel ? D.gx[k] : gv;
        hv = sel ? D.hx[k] : hv;
      }
    }
    const bool in = active && s >= slo && s <= shi;
    gxw[s] = in ? gv : 0.0;
    hxw[s] = in ? hv : 0.0;
  }
  const double fcx = active ? D.fcx : 0.0;
  const double fcz = active ? D.fcz : 0.0;

  // mode factors for every m > 0 (6 doubles per mode)
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
  // The radial prefix of particles.F90:658, jyh(ix) = jyh(ix)*ratio - fjy*w_xt(ix), is separable:
  // jyh(ix) = -w_xt(ix) * S with the real, mode-independent recurrence S = S*ratio + fjy.
  double S = 0.0;
  const bool h16 = lane & 16;
  const int b0 = lane & 1, b1 = (lane >> 1) & 1, b2 = (lane >> 2) & 1, b3 = (lane >> 3) & 1, b4 = (lane >> 4) & 1;
  const int pos0 = 2 * b0 + 4 * b1 + 8 * b2 + 16 * b3 + 32 * b4;

#pragma unroll 1
  for (int ky = 0; ky < 5; ++ky) {
    const int iy = ky - 2;
    const int cy = base_y + iy;
    const bool row_on = active && iy >= D.ymin && iy <= D.ymax;
    if (!__any_sync(0xffffffffu, row_on)) continue;
    // dynamic ky indexing of the 5-vectors through selects (keeps them in registers)
    double gyk = 0.0, hyk = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      gyk = (k == ky) ? D.gy[k] : gyk;
      hyk = (k == ky) ? D.hy[k] : hyk;
    }
    if (!row_on) { gyk = 0.0; hyk = 0.0; }
    const double iart = __ldg(&inv_area_rt[cy]), iaxt = __ldg(&inv_area_xt[cy]);
    const double ivol = __ldg(&inv_volume[cy]), ratio = __ldg(&ratio_area_xt[cy]);
    const double fjx = row_on ? fcx * iart : 0.0;
    const double fjz = row_on ? fcz * ivol : 0.0;
    if (row_on) S = S * ratio + (fcx * hyk) * iaxt;
    const double Srow = row_on ? S : 0.0;

    // 64-value super-chunks: values are generated in pairs (stream positions i and i + 32), the
    // level-1 exchange is done as soon as a pair exists, so only 32 partial sums are ever live
    double w[32];
    double pend = 0.0;
    int gen = 0;       // compile-time after unrolling
    int chunk = 0;
    auto flush = [&]() {
      warp_reduce_scatter_tail(w, lane);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int p = pos0 + i;
        const int gidx = (p < 32) ? 2 * p : 2 * (p - 32) + 1;
        const int e = chunk * 64 + gidx;
        int im, comp, slot, reim;
        if (w[i] != 0.0 && decode_row_element(e, M, im, comp, slot, reim)) {
          const int cx = base_x - 2 + slot;
          size_t o = g.at(cx, cy, im);
          double* arr = (comp == 0) ? P.jx : (comp == 1) ? P.jr : P.jt;
          o += (comp == 0) ? 1 : (comp == 1) ? (size_t)g.SX : 0;
          atomicAdd(arr + 2 * o + reim, w[i]);
        }
      }
      chunk += 1;
    };
#define PUSH_VALUE(val)                                            \
    do {                                                           \
      if ((gen & 1) == 0) {                                        \
        pend = (val);                                              \
      } else {                                                     \
        w[gen >> 1] = rs_exchange(pend, (val), h16, 16);           \
      }                                                            \
      gen += 1;                                                    \
      if (gen == 64) { flush(); gen = 0; }                         \
    } while (0)

    // ---- m = 0 (all real) ----
    {
      const double w_rt = gyk + 0.5 * hyk;
      const double ym1 = 0.5 * gyk + third * hyk;
      const double a = -(fjx * w_rt);
      double run = 0.0;
#pragma unroll
      for (int s = 0; s < WX; ++s) {
        run = run + hxw[s];
        PUSH_VALUE((s <= shi) ? a * run : 0.0);
      }
#pragma unroll
      for (int s = 0; s < WX; ++s) PUSH_VALUE(-(Srow * (gxw[s] + 0.5 * hxw[s])));
      const double bz = fjz * w_rt, cz = fjz * ym1;
#pragma unroll
      for (int s = 0; s < WX; ++s) PUSH_VALUE(gxw[s] * bz + hxw[s] * cz);
    }
    // ---- m > 0 ----
#pragma unroll
    for (int im = 1; im < M; ++im) {
      const cplx w_rt = f2[im - 1] * gyk + f3[im - 1] * hyk;
      const cplx ym1 = f3[im - 1] * gyk + f4[im - 1] * hyk;
      const cplx a = (-fjx) * w_rt;
      double run = 0.0;
#pragma unroll
      for (int s = 0; s < WX; ++s) {
        run = run + hxw[s];
        const double rr = (s <= shi) ? run : 0.0;
        PUSH_VALUE(a.x * rr);
        PUSH_VALUE(a.y * rr);
      }
      const cplx sf2 = (-Srow) * f2[im - 1], sf3 = (-Srow) * f3[im - 1];
#pragma unroll
      for (int s = 0; s < WX; ++s) {
        PUSH_VALUE(sf2.x * gxw[s] + sf3.x * hxw[s]);
        PUSH_VALUE(sf2.y * gxw[s] + sf3.y * hxw[s]);
      }
      const cplx bz = fjz * w_rt, cz = fjz * ym1;
#pragma unroll
      for (int s = 0; s < WX; ++s) {
        PUSH_VALUE(gxw[s] * bz.x + hxw[s] * cz.x);
        PUSH_VALUE(gxw[s] * bz.y + hxw[s] * cz.y);
      }
    }
    if (gen > 0) {
      // pad the last super-chunk with zeros
#pragma unroll
      for (int k = 0; k < 64; ++k)
        if (k >= gen) {
          if ((k & 1) == 0) pend = 0.0;
          else w[k >> 1] = (k == gen) ? rs_exchange(pend, 0.0, h16, 16) : 0.0;
        }
      flush();
    }
#undef PUSH_VALUE
  }
}

template <int M>
__global__ void __launch_bounds__(128, PUSH_MINB) k_push_v1(PushConst P, double* __restrict__ x, double* __restrict__ y,
                                                 double* __restrict__ z, double* __restrict__ px,
                                                 double* __restrict__ py, double* __restrict__ pz,
                                                 const double* __restrict__ w, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool valid = i < n;
  // idle lanes of the tail warp take part in the window deposit (shuffles): their D must be finite -- zero weights,
  // zero angle -- or 0 * Inf from an uninitialised register would reach the warp totals
  DepositIn D = {};
  D.exp_itheta_05 = C(1.0, 0.0);
  D.exp_idtheta = C(1.0, 0.0);
  D.cell_x2 = 0x3fffffff; D.cell_y2 = 0x3fffffff;
  if (valid) {
    double X = x[i], Y = y[i], Z = z[i], PX = px[i], PY = py[i], PZ = pz[i];
    const double W = w[i];
    push_one<M>(P, X, Y, Z, PX, PY, PZ, W, D);
    x[i] = X; y[i] = Y; z[i] = Z;
    px[i] = PX; py[i] = PY; pz[i] = PZ;
  }
  if (!P.deposit) return;
  const int base_x = __reduce_min_sync(0xffffffffu, D.cell_x2);
  const int base_y = __reduce_min_sync(0xffffffffu, D.cell_y2);
  if (base_x == 0x3fffffff) return;   // whole warp beyond n
  const int sx = D.cell_x2 - base_x;
  const bool inwin = valid && sx <= WX - 5 && D.cell_y2 == base_y;
  if (valid && !inwin) deposit_global(P, D);
  deposit_window<M>(P, D, inwin, lane, base_x, base_y, inwin ? sx : 0);
}

#include "pbcs_kernels.cuh"
#include "compact_kernels.cuh"

// ------------------------------------------------------------------------------------------
// variant 2: strip CTAs.  The sort of this push (do_sort) left the particles ordered by the
// staggered cell (cell_y2, cell_x2) they occupy after the half-step drift, with the start of
// every cell in `cell_start`.  One CTA owns a strip of STRIP_C consecutive cells of one row:
// it stages the (STRIP_C + 3) x 4 node patch of all 6 M mode arrays that those cells gather
// from in shared memory (mode 0 as reals), then walks its particles 128 at a time: gather from
// shared memory, Boris, store, warp-window deposit.  Lanes whose cell is not the predicted
// one (clamped sort keys, last-bit disagreements) gather from the mode arrays instead, and
// the window deposit runs as many passes as the warp has distinct windows (normally one), so
// nothing depends on the sort being exact.
// ------------------------------------------------------------------------------------------
// Strip width.  Measured in round 2 with the row-per-warp patch staging (profiles/r2f_push_strip_variants.txt,
// push kernel ms on C3 / C2 / C4): 16 cells 21.49 / 5.11 / 24.20 (19 of 32 lanes staging), 24 cells 20.52 / 5.07 /
// 20.98, 29 cells 20.60 / 5.23 / 20.77, 32 cells 21.22 / 5.17 / 22.38; round 1's element-wise staging at 16 cells:
// 21.07 / 5.06 / 21.50.
#ifndef STRIP_C
#define STRIP_C 24
#endif
#define STRIP_PC (STRIP_C + 3)

__device__ __forceinline__ const cplx* field_ptr(const PushConst& P, int comp) {
  return comp == 0 ? P.exm : comp == 1 ? P.erm : comp == 2 ? P.etm : comp == 3 ? P.bxm : comp == 4 ? P.brm : P.btm;
}

struct SoaIn { const double* d[7]; };
struct SoaOut { double* d[7]; };

// shared memory of one strip CTA (dynamic): [mode-0 patch: 6*CS doubles][m>0 patch: 6*(M-1)*CS cplx]
// [DMMA staging: 4 warps * MMA_WARP_DOUBLES doubles (variant 3 only)]
template <int M>
constexpr size_t strip_smem_bytes(bool mma) {
  return (size_t)6 * PATCH_ROWS * STRIP_PC * 8 + (size_t)6 * (M - 1) * PATCH_ROWS * STRIP_PC * 16 +
         (mma ? (size_t)4 * MMA_WARP_DOUBLES * 8 : 0) + 32 * 8;
}

// CTAs per SM the strip kernel is compiled for.  The kernel is bound by the LSU/shared-memory
// data pipe (82 % busy), so register spills (local memory goes through the same pipe) cost more
// than the warps they buy: 3 CTAs x 168 registers beat 4 x 128 and 5 x 96 (measured).  Above
// three modes the per-mode factors need still more registers and the patch more shared memory.
#ifndef STRIP_MINB
#define STRIP_MINB(M) ((M) <= 3 ? 3 : 2)
#endif
struct FusedBcs {   // particle_bcs fused into the strip push (enabled = 0: the push leaves it to k_pbcs_classify)
  BcsConst B;
  uint32_t* hole_list;
  uint8_t* hole_flag;
  unsigned long long* cnt;
  int enabled;
};

template <int M, bool MMA>
__global__ void __launch_bounds__(128, STRIP_MINB(M)) k_push_v2(PushConst P, SoaIn in, SoaOut out,
                                                                const uint32_t* __restrict__ src,
                                                                const int* __restrict__ cell_start, int ncx,
                                                                int nstrip_x, const __grid_constant__ FusedBcs FB) {
  constexpr int PC = STRIP_PC, CS = PATCH_ROWS * PC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s0 = reinterpret_cast<double*>(smem_raw);
  cplx* sm = reinterpret_cast<cplx*>(smem_raw + (size_t)6 * CS * 8);
  double* stab = reinterpret_cast<double*>(smem_raw + (size_t)6 * CS * 8 + (size_t)6 * (M - 1) * CS * 16);
  double* wbuf = stab + 32 + (threadIdx.x >> 5) * MMA_WARP_DOUBLES;
  const int srow = blockIdx.x / nstrip_x, scol = blockIdx.x - srow * nstrip_x;
  const int kx0 = scol * STRIP_C;
  const int nk = min(STRIP_C, ncx - kx0);
  const int key0 = srow * ncx + kx0;
  const int begin = cell_start[key0], end = cell_start[key0 + nk];
  if (begin >= end) return;
  const int c0 = kx0 + 1 - CELL_PAD, row0 = srow + 1 - CELL_PAD;   // cell_x2 / cell_y2 of the strip origin
  const Geom& g = P.g;
  // stage the patch: node (c0 - 1 + col, row0 - 1 + r).  A warp takes whole patch rows (component, mode, r): the
  // row's base address is formed once per warp, the lanes walk the columns (round 2: the element-wise index
  // arithmetic of the first version -- two divisions, two remainders and a six-way pointer select per element --
  // took 11 % of the kernel's stall samples on the 32-ppc workload, profiles/r2_push_c3_full_summary.txt)
  {
    const int ncol = nk + 3;
    const int warp = threadIdx.x >> 5, ln = threadIdx.x & 31;
    for (int row = warp; row < 6 * M * PATCH_ROWS; row += 4) {
      const int r = row % PATCH_ROWS;
      const int q = row / PATCH_ROWS;
      const int im = q % M, comp = q / M;
      const cplx* f = field_ptr(P, comp) + g.at(c0 - 1, row0 - 1 + r, im);
      if (im == 0) {
        double* d = s0 + (comp * PATCH_ROWS + r) * PC;
        for (int col = ln; col < ncol; col += 32) d[col] = __ldg((const double*)(f + col));
      } else {
        cplx* d = sm + ((comp * (M - 1) + im - 1) * PATCH_ROWS + r) * PC;
        for (int col = ln; col < ncol; col += 32) d[col] = __ldg(f + col);
      }
    }
  }
  // radial tables of the window rows row0-2 .. row0+2 (particles.F90:190-217): [table][ky]
  if (threadIdx.x < 20) {
    const int t = threadIdx.x / 5, ky = threadIdx.x - 5 * t;
    stab[threadIdx.x] = __ldg(P.tab + t * P.ntab + JNG + row0 - 2 + ky);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int base = begin; base < end; base += blockDim.x) {
    const int iraw = base + threadIdx.x;
    const bool valid = iraw < end;
    // sorted slot -> particle: the sort only built the permutation, the move happens here.
    // Idle lanes shadow the last particle of the strip: finite data, no store.
    DepositIn D;
    {
      const uint32_t j = __ldg(&src[valid ? iraw : end - 1]);
      double X = in.d[0][j], Y = in.d[1][j], Z = in.d[2][j], PX = in.d[3][j], PY = in.d[4][j], PZ = in.d[5][j];
      const double W = in.d[6][j];
      PushMid S;
      push_pre(P, X, Y, Z, PX, PY, PZ, S, D);
      Fields6 F;
      if (S.cell_y2 == row0 && S.cell_x2 >= c0 && S.cell_x2 < c0 + nk) F = gather_patch<M, PC>(s0, sm, S, c0, row0);
      else F = gather_global<M>(P, S);
      push_post(P, S, F, X, Y, Z, PX, PY, PZ, W, D);
      if (valid) {
        if (FB.enabled) {   // boundary.F90:1541-1865 on the values about to be stored
          RegParticle a{X, Y, Z, PX, PY, PZ};
          const uint8_t f = particle_bcs_one(FB.B, a);
          if (f != FL_KEEP) record_leaver(FB.hole_list, FB.hole_flag, FB.cnt, (uint32_t)iraw, f);
        }
        out.d[0][iraw] = X; out.d[1][iraw] = Y; out.d[2][iraw] = Z;
        out.d[3][iraw] = PX; out.d[4][iraw] = PY; out.d[5][iraw] = PZ;
        out.d[6][iraw] = W;
      }
    }
    if (!P.deposit) continue;
    // window passes: the first pending lane names the row, the smallest pending cell_x2 of
    // that row the window origin; every pass retires at least that lane
    unsigned pending = __ballot_sync(0xffffffffu, valid);
#pragma unroll 1
    while (pending) {
      const int leader = __ffs(pending) - 1;
      const int base_y = __shfl_sync(0xffffffffu, D.cell_y2, leader);
      const bool mine = ((pending >> lane) & 1u) && D.cell_y2 == base_y;
      const int base_x = __reduce_min_sync(0xffffffffu, mine ? D.cell_x2 : 0x3fffffff);
      const int sx = D.cell_x2 - base_x;
      const bool inwin = mine && sx <= (MMA ? MMA_WX : WX) - 5;
      if (MMA) deposit_mma<M>(P, D, inwin, lane, base_x, base_y, sx, wbuf, stab, row0);
      else deposit_window<M>(P, D, inwin, lane, base_x, base_y, inwin ? sx : 0);
      pending &= ~__ballot_sync(0xffffffffu, inwin);
    }
  }
}

__global__ void __launch_bounds__(256) k_copy_zero(cplx* __restrict__ old0, cplx* __restrict__ old1,
                                                   cplx* __restrict__ old2, cplx* __restrict__ j0,
                                                   cplx* __restrict__ j1, cplx* __restrict__ j2, size_t n) {
  const cplx zero = C(0.0, 0.0);
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    old0[t] = j0[t]; j0[t] = zero;
    old1[t] = j1[t]; j1[t] = zero;
    old2[t] = j2[t]; j2[t] = zero;
  }
}

int do_sort(cylgpu_ctx* c);
int do_sort_species(cylgpu_ctx* c, int isp, bool physical);

// particles.F90:163-169: j*_old = j*; j* = 0 (one fused streaming pass)
static int push_prologue(cylgpu_ctx* c) {
  const Geom& g = c->g;
  const size_t n = g.plane * g.M;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)148 * 8);
  k_copy_zero<<<blocks, 256, 0, c->stream>>>(c->f[CYLGPU_JXM_OLD], c->f[CYLGPU_JRM_OLD], c->f[CYLGPU_JTM_OLD],
                                             c->f[CYLGPU_JXM], c->f[CYLGPU_JRM], c->f[CYLGPU_JTM], n);
  c->stats.kernel_launches += 1;
  return 0;
}

// particles.F90:146-734 for the S.n particles of one species currently in S.d: (sort,) gather,
// Boris, store, deposit.  `time_kernel` brackets the fused kernel with events and a host sync.
static BcsConst make_bcs_const(cylgpu_ctx* c);
static int reserve_pscratch(cylgpu_ctx* c, int64_t n);

// `fuse_bcs`: the strip kernel also applies particle_bcs to what it stores and lists the leavers
// (counters zeroed here); *fused_out says whether it did (only the strip variants can).
static int push_species(cylgpu_ctx* c, int isp, bool need_sort, bool time_kernel, bool fuse_bcs = false,
                        bool* fused_out = nullptr) {
  const Geom& g = c->g;
  // strip kernels read through the sort permutation and write the second buffer set: the sort
  // then only has to build the permutation (no scatter passes)
  const bool strips = (c->push_variant == 2 || c->push_variant == 3) && c->sort_interval == 1;
  const double fac = SHAPE_FAC;   // particles.F90:145-153: (0.5)**c_ndims for the triangle
  const double dt = c->dt;
  cylgpu::SpeciesState& S = c->species[isp];
  if (fused_out) *fused_out = false;
  if (!strips && S.lazy) TRY(poll_counts(c, true));   // k_push_v0 / v1 take the exact count
  if (!S.set || S.sp.immobile || S.n == 0) return 0;
  if (c->xcap > 0 && fuse_bcs) TRY(reserve_particles(c, isp, S.n + 2 * c->xcap));   // arrivals of this step
  if (need_sort && !(c->presorted && strips)) {   // (presorted: the side stream did it behind the field phase)
    PhaseTimer sort_timer(c, &c->stats.ms_sort);
    TRY(do_sort_species(c, isp, /*physical=*/!strips));
  }
  FusedBcs FB;
  FB.enabled = 0;
  if (fuse_bcs && strips) {
    TRY(reserve_pscratch(c, S.n));
    FB.B = make_bcs_const(c);
    for (int k = 0; k < 4; ++k) FB.B.bc[k] = S.sp.bc_particle[k];
    FB.hole_list = c->hole_list;  // the sort's key / rank scratch: k_sort_src is done with it by now
    FB.hole_flag = c->flag;
    FB.cnt = c->counters;
    FB.enabled = 1;
    CUDA_TRY(cudaMemsetAsync(c->counters, 0, 8 * sizeof(unsigned long long), c->stream));
    if (fused_out) *fused_out = true;
  }
  PushConst P;
  P.g = g;
  P.exm = c->f[CYLGPU_EXM]; P.erm = c->f[CYLGPU_ERM]; P.etm = c->f[CYLGPU_ETM];
  P.bxm = c->f[CYLGPU_BXM]; P.brm = c->f[CYLGPU_BRM]; P.btm = c->f[CYLGPU_BTM];
  P.jx = (double*)c->f[CYLGPU_JXM]; P.jr = (double*)c->f[CYLGPU_JRM]; P.jt = (double*)c->f[CYLGPU_JTM];
  P.tab = c->tables; P.ntab = c->ntab;
  P.x_grid_min_local = c->x_grid_min_local;
  P.y_grid_min_local = c->cfg.y_grid_min_local;
  P.idx = 1.0 / c->cfg.dx; P.idy = 1.0 / c->cfg.dy; P.idt = 1.0 / dt;
  const double dto2 = dt / 2.0;
  P.dtco2 = C_LIGHT * dto2;
  const double dtfac = 0.5 * dt * fac;
  P.part_mc = C_LIGHT * S.sp.mass;
  P.ipart_mc = 1.0 / P.part_mc;
  P.cmratio = S.sp.charge * dtfac * P.ipart_mc;
  P.ccmratio = C_LIGHT * P.cmratio;
  P.q_fac = S.sp.charge * fac;
  P.deposit = S.sp.zero_current ? 0 : 1;
  P.hc_push = c->hc_push ? 1 : 0;
  P.taylor_switch = c->taylor_switch;
  P.hc_alpha = 0.5 * S.sp.charge * dt / S.sp.mass;
  const int64_t nb = (S.n + 127) / 128;
  PhaseTimer kernel_timer(c, &c->stats.ms_push_kernel, &c->stats.n_push_kernel, time_kernel);
  const int ncx = g.nx + 2 * CELL_PAD, ncy = g.ny + 2 * CELL_PAD;
  const int nstrip_x = (ncx + STRIP_C - 1) / STRIP_C;
  SoaIn pin; SoaOut pout;
  for (int q = 0; q < 7; ++q) { pin.d[q] = S.d[q]; pout.d[q] = S.alt[q]; }
#define PUSH_ARGS P, S.d[0], S.d[1], S.d[2], S.d[3], S.d[4], S.d[5], S.d[6], S.n
#define LAUNCH_STRIP(MM, MMA)                                                                              \
  do {                                                                                                     \
    const size_t shb = strip_smem_bytes<MM>(MMA);                                                          \
    static bool smem_opted_in[64] = {false};   /* per instantiation and device: the attribute is a driver call */ \
    if (!smem_opted_in[c->device & 63]) {                                                                  \
      CUDA_TRY(cudaFuncSetAttribute(k_push_v2<MM, MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shb)); \
      smem_opted_in[c->device & 63] = true;                                                                \
    }                                                                                                      \
    k_push_v2<MM, MMA><<<(unsigned)(nstrip_x * ncy), 128, shb, c->stream>>>(P, pin, pout, S.perm, S.cell_start, ncx, nstrip_x, FB); \
  } while (0)
#define LAUNCH_M(MM)                                                                                       \
  do {                                                                                                     \
    if (c->push_variant == 4) k_push_generic<MM><<<(unsigned)nb, 128, 0, c->stream>>>(PUSH_ARGS);         \
    else if (strips && c->push_variant == 3) LAUNCH_STRIP(MM, true);                                       \
    else if (strips) LAUNCH_STRIP(MM, false);                                                              \
    else if (c->push_variant >= 1) k_push_v1<MM><<<(unsigned)nb, 128, 0, c->stream>>>(PUSH_ARGS);         \
    else k_push_v0<MM><<<(unsigned)nb, 128, 0, c->stream>>>(PUSH_ARGS);                                   \
  } while (0)
  switch (g.M) {
    case 1: LAUNCH_M(1); break;
    case 2: LAUNCH_M(2); break;
    case 3: LAUNCH_M(3); break;
    case 4: LAUNCH_M(4); break;
    case 5: LAUNCH_M(5); break;
    case 6: LAUNCH_M(6); break;
    default: set_error("n_mode = %d not supported by the push kernels (1..6)", g.M); return 2;
  }
#undef LAUNCH_M
#undef LAUNCH_STRIP
#undef PUSH_ARGS
  if (strips) for (int q = 0; q < 7; ++q) std::swap(S.d[q], S.alt[q]);
  c->stats.kernel_launches += 1;
  kernel_timer.stop();   // per-launch device time of the fused kernel (the roofline numerator's clock)
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int do_push(cylgpu_ctx* c) {
  TRY(presort_join(c, false));
  TRY(flush_pending_remove(c));
  c->r_clean = false;   // pushed without particle_bcs
  TRY(push_prologue(c));
  const bool need_sort = c->sort_interval > 0 && (!c->sorted_valid || c->pushes_since_sort >= c->sort_interval);
  for (int isp = 0; isp < c->cfg.n_species; ++isp) {
    TRY(push_species(c, isp, need_sort, c->timing));
  }
  if (need_sort) {
    c->sorted_valid = true;
    c->pushes_since_sort = 0;
    c->stats.n_sorts += 1;
  }
  c->pushes_since_sort += 1;
  return do_r_min_final(c);
}

// ------------------------------------------------------------------------------------------
// particle storage
// ------------------------------------------------------------------------------------------
int reserve_particles(cylgpu_ctx* c, int isp, int64_t n) {
  cylgpu::SpeciesState& S = c->species[isp];
  if (n <= S.cap) return 0;
  int64_t cap = std::max<int64_t>(n, (int64_t)(S.cap * 1.25) + 1024);
  for (int k = 0; k < 7; ++k) {
    double* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, (size_t)cap * sizeof(double)));
    if (S.n > 0) CUDA_TRY(cudaMemcpyAsync(p, S.d[k], (size_t)S.n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (S.d[k]) CUDA_TRY(cudaFree(S.d[k]));
    S.d[k] = p;
  }
  S.cap = cap;
  return 0;
}

static int reserve_pscratch(cylgpu_ctx* c, int64_t n) {
  if (n <= c->pscratch_cap) return 0;
  const int64_t cap = (int64_t)(n * 1.25) + 1024;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (c->perm) cudaFree(c->perm);
  if (c->flag) cudaFree(c->flag);
  if (c->tailmark) cudaFree(c->tailmark);
  if (c->hole_list) cudaFree(c->hole_list);
  if (c->lowhole) cudaFree(c->lowhole);
  if (c->hightail) cudaFree(c->hightail);
  CUDA_TRY(cudaMalloc(&c->perm, (size_t)cap * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&c->flag, (size_t)cap));
  CUDA_TRY(cudaMalloc(&c->tailmark, (size_t)cap));
  CUDA_TRY(cudaMalloc(&c->hole_list, (size_t)cap * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&c->lowhole, (size_t)cap * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&c->hightail, (size_t)cap * sizeof(uint32_t)));
  c->pscratch_cap = cap;
  return 0;
}

// ------------------------------------------------------------------------------------------
// particle_bcs, boundary.F90:1541-1889: host side (the per-particle rules are particle_bcs_one)
// ------------------------------------------------------------------------------------------
// window.F90:304-325 remove_particles: everything behind the new x_min goes
__global__ void __launch_bounds__(256) k_flag_behind(const double* __restrict__ x, double x_min,
                                                     uint32_t* __restrict__ hole_list, uint8_t* __restrict__ hole_flag,
                                                     unsigned long long* cnt, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (x[i] < x_min) record_leaver(hole_list, hole_flag, cnt, (uint32_t)i, FL_GONE);
}


// second pass over the leavers: pack the migrants in pack_particle order (7 doubles), split the
// holes into those below the new count (to be filled) and those in the tail (marked)
__global__ void __launch_bounds__(256) k_collect(Soa s, const uint32_t* __restrict__ hole_list,
                                                 const uint8_t* __restrict__ hole_flag, int64_t nholes, int64_t n_new,
                                                 double* __restrict__ send_l, double* __restrict__ send_r,
                                                 uint32_t* __restrict__ lowhole, uint8_t* __restrict__ tailmark,
                                                 unsigned long long* cnt2) {
  // cnt2[1] = packed left, cnt2[2] = packed right, cnt2[4] = holes below the new count
  const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= nholes) return;
  const uint32_t i = hole_list[h];
  const uint8_t f = hole_flag[h];
  if (f == FL_LEFT || f == FL_RIGHT) {
    const unsigned long long k = atomicAdd(&cnt2[f], 1ULL);
    double* dst = ((f == FL_LEFT) ? send_l : send_r) + 7 * k;
#pragma unroll
    for (int q = 0; q < 7; ++q) dst[q] = s.d[q][i];
  }
  if ((int64_t)i >= n_new) tailmark[(int64_t)i - n_new] = 1;
  else lowhole[atomicAdd(&cnt2[4], 1ULL)] = i;
}

__global__ void __launch_bounds__(256) k_tail_keepers(const uint8_t* __restrict__ tailmark, int64_t n_new,
                                                      int64_t nholes, uint32_t* __restrict__ hightail,
                                                      unsigned long long* cnt2) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nholes) return;
  if (!tailmark[t]) hightail[atomicAdd(&cnt2[3], 1ULL)] = (uint32_t)(n_new + t);
}

__global__ void __launch_bounds__(256) k_fill_holes(Soa s, const uint32_t* __restrict__ lowhole,
                                                    const uint32_t* __restrict__ hightail,
                                                    const unsigned long long* cnt2, int64_t nmax) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nmax || (unsigned long long)k >= cnt2[4]) return;
  const uint32_t dst = lowhole[k], src = hightail[k];
#pragma unroll
  for (int q = 0; q < 7; ++q) s.d[q][dst] = s.d[q][src];
}

__global__ void __launch_bounds__(256) k_unpack(Soa s, int64_t base, const double* __restrict__ recv, int64_t n) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
#pragma unroll
  for (int q = 0; q < 7; ++q) s.d[q][base + k] = recv[7 * k + q];
}

static inline unsigned nblk(int64_t n, int b) { return (unsigned)((n + b - 1) / b); }

static int ensure_dbuf(double** p, int64_t* cap, int64_t need, cudaStream_t st) {
  if (need <= *cap) return 0;
  CUDA_TRY(cudaStreamSynchronize(st));
  if (*p) cudaFree(*p);
  const int64_t nc = (int64_t)(need * 1.5) + 4096;
  CUDA_TRY(cudaMalloc(p, (size_t)nc * sizeof(double)));
  *cap = nc;
  return 0;
}

// compaction once the leavers are listed and counted (counts on the host): fills the holes below
// the new count with the keepers above it (order is not preserved; the reference's list order
// only matters for floating-point summation order)
static int compact(cylgpu_ctx* c, cylgpu::SpeciesState& S, int64_t nholes, int64_t off_l, int64_t off_r,
                   unsigned long long* cnt2) {
  Soa s;
  for (int q = 0; q < 7; ++q) s.d[q] = S.d[q];
  const int64_t n = S.n, n_new = n - nholes;
  CUDA_TRY(cudaMemsetAsync(cnt2, 0, 8 * sizeof(unsigned long long), c->stream));
  CUDA_TRY(cudaMemsetAsync(c->tailmark, 0, (size_t)nholes, c->stream));
  k_collect<<<nblk(nholes, 256), 256, 0, c->stream>>>(s, c->hole_list, c->flag, nholes, n_new, c->psend_l + 7 * off_l,
                                                      c->psend_r + 7 * off_r, c->lowhole, c->tailmark, cnt2);
  c->stats.kernel_launches += 1;
  if (n_new > 0) {
    k_tail_keepers<<<nblk(nholes, 256), 256, 0, c->stream>>>(c->tailmark, n_new, nholes, c->hightail, cnt2);
    k_fill_holes<<<nblk(nholes, 256), 256, 0, c->stream>>>(s, c->lowhole, c->hightail, cnt2, nholes);
    c->stats.kernel_launches += 2;
  }
  CUDA_TRY(cudaGetLastError());
  S.n = n_new;
  return 0;   // (callers mirror the count on the device where device-resident counts are on)
}

// The step's host syncs (particle counts) wait for the whole push kernel.  With one rank per GPU on
// a shared host, a spinning wait steals cycles from the other ranks' launching threads; the
// blocking wait (one event, cudaEventBlockingSync) yields the core instead.
static int host_wait(cylgpu_ctx* c) {
  if (!c->blocking_wait) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
  }
  if (!c->ev_wait) CUDA_TRY(cudaEventCreateWithFlags(&c->ev_wait, cudaEventBlockingSync | cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(c->ev_wait, c->stream));
  CUDA_TRY(cudaEventSynchronize(c->ev_wait));
  return 0;
}

static BcsConst make_bcs_const(cylgpu_ctx* c) {
  BcsConst B;
  const double dx = c->cfg.dx, dy = c->cfg.dy;
  B.x_min = c->x_min; B.x_max = c->x_max;
  B.x_min_local = c->x_min_local; B.x_max_local = c->x_max_local;
  B.y_max = c->cfg.y_max;
  double boundary_shift = dx * (double)((1 + PNG + 0) / 2);   // boundary.F90:1561-1563
  B.x_min_outer = B.x_min - boundary_shift;
  B.x_max_outer = B.x_max + boundary_shift;
  boundary_shift = dy * (double)((1 + PNG + 0) / 2);
  B.y_max_outer = B.y_max + boundary_shift;
  B.x_shift = B.x_max - B.x_min;   // length_x
  B.y_max2_inside = B.y_max * B.y_max * (1.0 - 1.0e-14);
  B.x_min_boundary = c->cfg.x_min_boundary;
  B.x_max_boundary = c->cfg.x_max_boundary;
  B.remove_x = -1.0e300;
  B.x_only = 0;
  return B;
}

// grow a device buffer of doubles, keeping its first `keep` entries
static int grow_dbuf_keep(double** p, int64_t* cap, int64_t need, int64_t keep, cudaStream_t st) {
  if (need <= *cap) return 0;
  const int64_t nc = (int64_t)(need * 1.5) + 4096;
  double* q = nullptr;
  CUDA_TRY(cudaMalloc(&q, (size_t)nc * sizeof(double)));
  if (*p && keep > 0) CUDA_TRY(cudaMemcpyAsync(q, *p, (size_t)keep * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (*p) cudaFree(*p);
  *p = q;
  *cap = nc;
  return 0;
}

// boundary.F90:1541-1865 for the S.n particles of one species in S.d: boundary conditions,
// classification, removal of the leavers (hole filling) and packing of the migrants behind the
// `off_l` / `off_r` particles already waiting in psend_l / psend_r.  One host sync (the counts).
static int pbcs_classify_compact(cylgpu_ctx* c, int isp, BcsConst B, int64_t off_l, int64_t off_r, int64_t* nleft_out,
                                 int64_t* nright_out, bool classified_by_push = false,
                                 const int* alive_dev = nullptr) {
  unsigned long long* cnt = c->counters;        // 8 for classify
  unsigned long long* cnt2 = c->counters + 8;   // 8 for collect/compact
  cylgpu::SpeciesState& S = c->species[isp];
  for (int k = 0; k < 4; ++k) B.bc[k] = S.sp.bc_particle[k];
  if (!classified_by_push) {
    TRY(reserve_pscratch(c, S.n));
    CUDA_TRY(cudaMemsetAsync(cnt, 0, 8 * sizeof(unsigned long long), c->stream));
    if (S.n > 0) {
      k_pbcs_classify<<<nblk(S.n, 256), 256, 0, c->stream>>>(B, S.d[0], S.d[1], S.d[2], S.d[3], S.d[4], S.d[5],
                                                            c->hole_list, c->flag, cnt, S.n);
      c->stats.kernel_launches += 1;
    }
  }
  CUDA_TRY(cudaMemcpyAsync(c->h_counters, cnt, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  if (alive_dev)   // the sort sent the particles behind the window to its extra bucket: the list ends before them
    CUDA_TRY(cudaMemcpyAsync(c->h_counters + 24, alive_dev, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  TRY(host_wait(c));
  if (alive_dev) {
    const int64_t alive = (int64_t)*reinterpret_cast<const int*>(c->h_counters + 24);
    c->stats.n_window_removed += S.n - alive;
    S.n = alive;
  }
  const int64_t nholes = (int64_t)c->h_counters[CNT_HOLE];
  const int64_t nleft = (int64_t)c->h_counters[CNT_LEFT];
  const int64_t nright = (int64_t)c->h_counters[CNT_RIGHT];
  const int64_t ngone = (int64_t)c->h_counters[CNT_GONE];
  c->stats.n_sent_left += nleft;
  c->stats.n_sent_right += nright;
  c->stats.n_removed += ngone;
  TRY(grow_dbuf_keep(&c->psend_l, &c->psend_l_cap, 7 * (off_l + nleft), 7 * off_l, c->stream));
  TRY(grow_dbuf_keep(&c->psend_r, &c->psend_r_cap, 7 * (off_r + nright), 7 * off_r, c->stream));
  if (nholes > 0) TRY(compact(c, S, nholes, off_l, off_r, cnt2));
  *nleft_out = nleft;
  *nright_out = nright;
  return 0;
}

// partlist_sendrecv (count, then payload), boundary.F90:1867-1877.  The received particles end
// up in c->precv as [from_l block][from_r block] in pack_particle order.
static int pbcs_exchange(cylgpu_ctx* c, int64_t nleft, int64_t nright, int64_t* from_l_out, int64_t* from_r_out) {
  unsigned long long* xc = c->counters + 16;    // [0..1] send counts, [2..3] recv counts
  const bool has_l = c->left >= 0, has_r = c->right >= 0;
  *from_l_out = *from_r_out = 0;
  if (!has_l && !has_r) return 0;
  c->h_counters[16] = (unsigned long long)nleft;
  c->h_counters[17] = (unsigned long long)nright;
  c->h_counters[18] = c->h_counters[19] = 0;
  CUDA_TRY(cudaMemcpyAsync(xc, c->h_counters + 16, 4 * sizeof(unsigned long long), cudaMemcpyHostToDevice,
                           c->stream));
  TRY(transport_sendrecv(c, has_l ? xc + 0 : nullptr, has_l ? 8 : 0, has_l ? xc + 2 : nullptr, has_l ? 8 : 0,
                         has_r ? xc + 1 : nullptr, has_r ? 8 : 0, has_r ? xc + 3 : nullptr, has_r ? 8 : 0));
  CUDA_TRY(cudaMemcpyAsync(c->h_counters + 16, xc, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                           c->stream));
  TRY(host_wait(c));
  const int64_t from_l = has_l ? (int64_t)c->h_counters[18] : 0;
  const int64_t from_r = has_r ? (int64_t)c->h_counters[19] : 0;
  TRY(ensure_dbuf(&c->precv, &c->precv_cap, 7 * (from_l + from_r), c->stream));
  double* rl = c->precv;
  double* rr = c->precv + 7 * from_l;
  TRY(transport_sendrecv(c, c->psend_l, has_l ? 7 * nleft * sizeof(double) : 0, rl, 7 * from_l * sizeof(double),
                         c->psend_r, has_r ? 7 * nright * sizeof(double) : 0, rr,
                         7 * from_r * sizeof(double)));
  *from_l_out = from_l;
  *from_r_out = from_r;
  return 0;
}

// ------------------------------------------------------------------------------------------
// particle_bcs with DEVICE-RESIDENT counts (cylgpu_set_exchange_capacity > 0): the same steps as
// pbcs_classify_compact / pbcs_exchange / the arrivals of pbcs_species, but every count (leavers, migrants per
// direction, arrivals, the new list length) is read by the kernels from device memory, and the migrants of
// one direction travel in ONE message of fixed size -- [7-double header: count][xcap slots of 7 doubles] --
// so that the host never waits for the device inside a step.  The host keeps upper bounds of the list
// lengths (SpeciesState::n with lazy = true) and tightens them from copies that trail behind (poll_counts).
// Replaces the two MPI_SENDRECVs per direction of partlist_sendrecv (partlist.F90:842,869: count, then data).
// ------------------------------------------------------------------------------------------
// (kernels: compact_kernels.cuh)
// the host's count of species isp is exact and has just changed: mirror it on the device
__global__ void k_set_count(int64_t* n_dev, long long v) { *n_dev = v; }
int set_count_exact(cylgpu_ctx* c, int isp) {
  if (!c->n_dev) return 0;
  k_set_count<<<1, 1, 0, c->stream>>>(c->n_dev + isp, (long long)c->species[isp].n);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// enqueue the copy of the device counts and statistics to the host (ring of 8; nobody waits here)
int publish_counts(cylgpu_ctx* c) {
  const int slot = (int)(c->pub_head % 8);
  cylgpu_ctx::Publish& P = c->pub[slot];
  if (P.pending) TRY(poll_counts(c, false));
  if (P.pending) {   // the host is 8 publishes ahead of the device: wait for this slot
    CUDA_TRY(cudaEventSynchronize(P.ev));
    TRY(poll_counts(c, false));
  }
  if (!P.ev) CUDA_TRY(cudaEventCreateWithFlags(&P.ev, cudaEventDisableTiming));
  const size_t words = CYLGPU_MAX_SPECIES + PST_N;
  CUDA_TRY(cudaMemcpyAsync(c->h_pub + (size_t)slot * words, c->n_dev, words * sizeof(int64_t), cudaMemcpyDeviceToHost,
                           c->stream));
  CUDA_TRY(cudaEventRecord(P.ev, c->stream));
  for (int i = 0; i < CYLGPU_MAX_SPECIES; ++i) P.bound_at[i] = c->species[i].n;
  P.pending = true;
  c->pub_head += 1;
  return 0;
}

// Tighten the host's upper bounds with the newest copy that has arrived: bound -= (bound then - exact then).
// block = true waits for the newest copy, after which every count is exact again (every device-side change of
// a count is followed by a publish).
int poll_counts(cylgpu_ctx* c, bool block) {
  if (!c->n_dev || c->pub_head == 0) return 0;
  const size_t words = CYLGPU_MAX_SPECIES + PST_N;
  bool any_lazy = false;
  for (int i = 0; i < CYLGPU_MAX_SPECIES; ++i) any_lazy = any_lazy || c->species[i].lazy;
  for (uint64_t back = 0; back < 8 && back < c->pub_head; ++back) {
    const uint64_t id = c->pub_head - 1 - back;
    cylgpu_ctx::Publish& P = c->pub[id % 8];
    if (!P.pending) break;   // everything older has been consumed already
    if (back == 0 && block) CUDA_TRY(cudaEventSynchronize(P.ev));
    else if (cudaEventQuery(P.ev) != cudaSuccess) { cudaGetLastError(); continue; }
    const int64_t* h = c->h_pub + (size_t)(id % 8) * words;
    for (int i = 0; i < CYLGPU_MAX_SPECIES; ++i) {
      cylgpu::SpeciesState& S = c->species[i];
      if (!S.lazy) continue;
      const int64_t slack = P.bound_at[i] - h[i];
      S.n -= slack;
      // the copies still on their way were enqueued with bounds that contain this slack
      for (uint64_t b2 = 0; b2 < back; ++b2) c->pub[(c->pub_head - 1 - b2) % 8].bound_at[i] -= slack;
      if (back == 0) S.lazy = false;   // the newest publish: nothing has changed since
    }
    const int64_t* ps = h + CYLGPU_MAX_SPECIES;
    c->stats.n_sent_left = ps[PST_SENT_L];
    c->stats.n_sent_right = ps[PST_SENT_R];
    c->stats.n_removed = ps[PST_REMOVED];
    c->stats.n_recv = ps[PST_RECV];
    c->stats.n_window_removed = ps[PST_WINDOW_REMOVED];
    if (ps[PST_OVERFLOW]) c->overflowed = true;
    // this copy and all older ones are consumed
    for (uint64_t b2 = back; b2 < 8 && b2 < c->pub_head; ++b2) c->pub[(c->pub_head - 1 - b2) % 8].pending = false;
    break;
  }
  if (c->overflowed) {
    set_error("particle exchange overflow: more than %lld particles left a slab towards one neighbour in one step; "
              "raise cylgpu_set_exchange_capacity (or set it to 0 for the exact protocol)", (long long)c->xcap);
    return 7;
  }
  (void)any_lazy;
  return 0;
}

static int ensure_xbufs(cylgpu_ctx* c) {
  const int64_t need = 7 * c->xcap + XHDR;
  TRY(ensure_dbuf(&c->psend_l, &c->psend_l_cap, need, c->stream));
  TRY(ensure_dbuf(&c->psend_r, &c->psend_r_cap, need, c->stream));
  TRY(ensure_dbuf(&c->precv, &c->precv_cap, 2 * need, c->stream));
  return 0;
}

// compaction of species isp after its leavers have been listed (c->hole_list / c->flag / c->counters)
static int compact_dev(cylgpu_ctx* c, int isp, bool window, bool has_l, bool has_r) {
  cylgpu::SpeciesState& S = c->species[isp];
  unsigned long long* cnt = c->counters;
  unsigned long long* cnt2 = c->counters + 8;
  Soa s;
  for (int q = 0; q < 7; ++q) s.d[q] = S.d[q];
  int64_t* pstats = c->n_dev + CYLGPU_MAX_SPECIES;
  k_plan_compact<<<1, 1, 0, c->stream>>>(cnt, cnt2, c->n_dev + isp, c->d_plan, pstats, (long long)c->xcap,
                                         window ? 1 : 0, has_l ? c->psend_l : nullptr, has_r ? c->psend_r : nullptr);
  k_clear_tail<<<LEAVER_GRID, 256, 0, c->stream>>>(c->tailmark, c->d_plan);
  k_collect_dev<<<LEAVER_GRID, 256, 0, c->stream>>>(s, c->hole_list, c->flag, c->d_plan, c->psend_l + XHDR,
                                                    c->psend_r + XHDR, has_l || has_r ? (long long)c->xcap : 0LL,
                                                    c->lowhole, c->tailmark, cnt2);
  k_tail_keepers_dev<<<LEAVER_GRID, 256, 0, c->stream>>>(c->tailmark, c->d_plan, c->hightail, cnt2);
  k_fill_holes_dev<<<LEAVER_GRID, 256, 0, c->stream>>>(s, c->lowhole, c->hightail, c->d_plan, cnt2);
  c->stats.kernel_launches += 5;
  CUDA_TRY(cudaGetLastError());
  S.lazy = true;   // S.n stays as an upper bound
  return 0;
}

static int pbcs_species_fast(cylgpu_ctx* c, int isp, BcsConst B, bool classified_by_push) {
  cylgpu::SpeciesState& S = c->species[isp];
  const bool has_l = c->left >= 0, has_r = c->right >= 0;
  for (int k = 0; k < 4; ++k) B.bc[k] = S.sp.bc_particle[k];
  TRY(ensure_xbufs(c));
  if (has_l || has_r) TRY(reserve_particles(c, isp, S.n + 2 * c->xcap));
  if (!classified_by_push) {
    TRY(reserve_pscratch(c, S.n));
    CUDA_TRY(cudaMemsetAsync(c->counters, 0, 8 * sizeof(unsigned long long), c->stream));
    if (S.n > 0) {
      k_pbcs_classify_dev<<<nblk(S.n, 256), 256, 0, c->stream>>>(B, S.d[0], S.d[1], S.d[2], S.d[3], S.d[4], S.d[5],
                                                                c->hole_list, c->flag, c->counters, c->n_dev + isp);
      c->stats.kernel_launches += 1;
    }
  }
  TRY(compact_dev(c, isp, false, has_l, has_r));
  if (has_l || has_r) {
    const size_t bytes = (size_t)(7 * c->xcap + XHDR) * sizeof(double);
    double* rl = c->precv;
    double* rr = c->precv + (7 * c->xcap + XHDR);
    c->msg_counted = true;    // (a mailbox link moves the slots in use only)
    const int xr = transport_sendrecv(c, c->psend_l, has_l ? bytes : 0, rl, has_l ? bytes : 0, c->psend_r,
                                      has_r ? bytes : 0, rr, has_r ? bytes : 0);
    c->msg_counted = false;
    TRY(xr);
    Soa s;
    for (int q = 0; q < 7; ++q) s.d[q] = S.d[q];
    k_unpack_dev<<<LEAVER_GRID, 256, 0, c->stream>>>(s, c->n_dev + isp, has_r ? rr : nullptr, has_l ? rl : nullptr);
    k_bump_count<<<1, 1, 0, c->stream>>>(c->n_dev + isp, c->n_dev + CYLGPU_MAX_SPECIES, has_r ? rr : nullptr,
                                         has_l ? rl : nullptr);
    c->stats.kernel_launches += 2;
    CUDA_TRY(cudaGetLastError());
    S.n += ((has_l ? 1 : 0) + (has_r ? 1 : 0)) * c->xcap;   // upper bound
  }
  return 0;
}

static int zero_pstats(cylgpu_ctx* c) {
  if (!c->n_dev) return 0;
  // sent / removed / received of this particle_bcs (the window's removals and the overflow mark stay)
  CUDA_TRY(cudaMemsetAsync(c->n_dev + CYLGPU_MAX_SPECIES, 0, 4 * sizeof(int64_t), c->stream));
  return 0;
}

// particle_bcs of one species: (classification,) compaction, exchange, arrivals appended
static int pbcs_species(cylgpu_ctx* c, int isp, const BcsConst& B, bool classified_by_push) {
  if (c->xcap > 0) return pbcs_species_fast(c, isp, B, classified_by_push);
  cylgpu::SpeciesState& S = c->species[isp];
  int64_t nleft = 0, nright = 0, from_l = 0, from_r = 0;
  TRY(pbcs_classify_compact(c, isp, B, 0, 0, &nleft, &nright, classified_by_push));
  TRY(pbcs_exchange(c, nleft, nright, &from_l, &from_r));
  // the reference receives from the right neighbour first (ix = -1 iteration), then left
  const int64_t nrecv = from_l + from_r;
  if (nrecv > 0) {
    TRY(reserve_particles(c, isp, S.n + nrecv));
    Soa s;
    for (int q = 0; q < 7; ++q) s.d[q] = S.d[q];
    const double* rl = c->precv;
    const double* rr = c->precv + 7 * from_l;
    if (from_r > 0) k_unpack<<<nblk(from_r, 256), 256, 0, c->stream>>>(s, S.n, rr, from_r);
    if (from_l > 0) k_unpack<<<nblk(from_l, 256), 256, 0, c->stream>>>(s, S.n + from_r, rl, from_l);
    c->stats.kernel_launches += (from_r > 0) + (from_l > 0);
    S.n += nrecv;
    c->stats.n_recv += nrecv;
    CUDA_TRY(cudaGetLastError());
  }
  c->stats.n_particles[isp] = S.n;
  return set_count_exact(c, isp);
}

int do_particle_bcs(cylgpu_ctx* c) {
  TRY(presort_join(c, false));
  BcsConst B = make_bcs_const(c);
  c->stats.n_sent_left = c->stats.n_sent_right = c->stats.n_removed = c->stats.n_recv = 0;
  if (c->xcap > 0) {
    TRY(zero_pstats(c));
    // after a window shift (window.F90:364): the removal behind the window rides on this classification, and a
    // list the fused push left inside in r is only tested against the x faces that moved
    if (c->pending_remove) B.remove_x = c->pending_remove_x;
    B.x_only = c->r_clean ? 1 : 0;
  }
  for (int isp = 0; isp < c->cfg.n_species; ++isp)
    if (c->species[isp].set) TRY(pbcs_species(c, isp, B, false));
  if (c->xcap > 0) {
    c->pending_remove = false;
    c->r_clean = true;   // everything outside has been dealt with
    TRY(publish_counts(c));
  }
  return 0;
}

// remove_particles that a window shift left pending (device-resident counts): normally it rides on the
// particle_bcs that follows the shift; anything else that looks at the lists first makes it happen here
int flush_pending_remove(cylgpu_ctx* c) {
  if (!c->pending_remove) return 0;
  c->pending_remove = false;
  if (!c->cfg.x_min_boundary) return 0;
  for (int isp = 0; isp < c->cfg.n_species; ++isp) {
    cylgpu::SpeciesState& S = c->species[isp];
    if (!S.set || S.n == 0) continue;
    TRY(reserve_pscratch(c, S.n));
    TRY(ensure_xbufs(c));
    CUDA_TRY(cudaMemsetAsync(c->counters, 0, 8 * sizeof(unsigned long long), c->stream));
    k_flag_behind_dev<<<nblk(S.n, 256), 256, 0, c->stream>>>(S.d[0], c->pending_remove_x, c->hole_list, c->flag,
                                                            c->counters, c->n_dev + isp);
    c->stats.kernel_launches += 1;
    TRY(compact_dev(c, isp, true, false, false));
  }
  return publish_counts(c);
}

// The cell sort of the next push on the side stream (see ctx.cuh): everything enqueued on the library stream so far
// precedes it, nothing enqueued on the library stream afterwards touches the particle lists or their scratch until
// presort_join.  Only with device-resident counts (no host sync inside the sort) and the strip push.
static int ensure_side(cylgpu_ctx* c) {
  if (c->side) return 0;
  CUDA_TRY(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ev_pfork, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&c->ev_pdone, cudaEventDisableTiming));
  return 0;
}

int presort_fork(cylgpu_ctx* c) {
  if (c->presorted || c->xcap <= 0 || (c->push_variant != 2 && c->push_variant != 3) || c->sort_interval != 1 || c->pending_remove) return 0;
  // Worth it only where the field phase is latency: a slab of ~1 M cell-modes (C3 over 8 GPUs: 3.73 against 3.81 ms
  // per step).  On a big slab both are bandwidth and overlapping them gains nothing (measured on C3 at 1 and 2
  // GPUs: 24.9 / 13.0 ms per step either way; at 4 GPUs 6.44 against 6.48, profiles/r2j_bench_c3_n4_presort_forced.json).
  // CYLGPU_PRESORT=0/1 overrides.
  if (c->presort_policy < 0) {
    c->presort_policy = ((size_t)c->g.nx * c->g.ny * c->g.M <= ((size_t)3 << 19)) ? 1 : 0;
    if (const char* e = getenv("CYLGPU_PRESORT")) c->presort_policy = atoi(e) != 0;
  }
  if (!c->presort_policy) return 0;
  bool any = false;
  for (int isp = 0; isp < c->cfg.n_species; ++isp) {
    const cylgpu::SpeciesState& S = c->species[isp];
    any = any || (S.set && !S.sp.immobile && S.n > 0);
  }
  if (!any) return 0;
  TRY(ensure_side(c));
  if (c->side_pending) {   // (the side stream is in order: the sort follows the chain there; the library stream joins)
    CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_pdone, 0));
    c->side_pending = false;
  }
  TRY(poll_counts(c, false));
  // capacities first, on the library stream: the sort sizes the second buffer set by them
  for (int isp = 0; isp < c->cfg.n_species; ++isp) {
    cylgpu::SpeciesState& S = c->species[isp];
    if (S.set && !S.sp.immobile && S.n > 0) TRY(reserve_particles(c, isp, S.n + 2 * c->xcap));
  }
  CUDA_TRY(cudaEventRecord(c->ev_fork, c->stream));
  CUDA_TRY(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
  cudaStream_t lib = c->stream;
  c->stream = c->side;
  int rc = 0;
  {
    PhaseTimer sort_timer(c, &c->stats.ms_sort);
    for (int isp = 0; isp < c->cfg.n_species && rc == 0; ++isp) rc = do_sort_species(c, isp, /*physical=*/false);
  }
  c->stream = lib;
  if (rc != 0) return rc;
  CUDA_TRY(cudaEventRecord(c->ev_join, c->side));
  c->presorted = true;
  return 0;
}

int presort_join(cylgpu_ctx* c, bool still_valid) {
  if (c->side_pending) {   // the particle chain of the last push
    CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_pdone, 0));
    c->side_pending = false;
  }
  if (!c->presorted) return 0;
  CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  if (!still_valid) c->presorted = false;
  return 0;
}

// push_particles including its particle_bcs call (particles.F90:28-734): species by species,
// the strip kernel applying the boundary rules to what it stores, so that the list is read
// once per step
int do_push_bcs(cylgpu_ctx* c) {
  const BcsConst B = make_bcs_const(c);
  c->stats.n_sent_left = c->stats.n_sent_right = c->stats.n_removed = c->stats.n_recv = 0;
  // the library stream joins what the side stream holds: the particle chain of the last push and the pre-sort
  // (a removal still pending would change the lists: the pre-sort is then void -- never in the reference's order)
  TRY(presort_join(c, !c->pending_remove));
  if (c->xcap > 0) {
    TRY(flush_pending_remove(c));
    TRY(poll_counts(c, false));   // whatever count copies have arrived tighten the bounds; nobody waits
    TRY(zero_pstats(c));
  }
  TRY(push_prologue(c));
  const bool need_sort = c->sort_interval > 0 && (!c->sorted_valid || c->pushes_since_sort >= c->sort_interval);
  int last = -1;
  for (int isp = 0; isp < c->cfg.n_species; ++isp) if (c->species[isp].set) last = isp;
  // Opt-in (CYLGPU_SIDE_CHAIN=1): the particle chain of the last species runs on the side stream beside
  // current_finish / update_eb_fields_final.  Parity-tested, but measured SLOWER with neighbours (small LWFA slab on
  // 2 GPUs: 0.89 against 0.65 ms per step): the exchanges of the two communicators wait on each other across the
  // ranks' streams.  The pre-sort, which has no exchange in it, is what stays on by default.
  static const bool side_chain_env = [] { const char* e = getenv("CYLGPU_SIDE_CHAIN"); return e && atoi(e) != 0; }();
  const bool side_chain = side_chain_env && c->xcap > 0 && c->presort_policy == 1 && transport_two_streams(c);
  bool published = false;
  for (int isp = 0; isp < c->cfg.n_species; ++isp) {
    if (!c->species[isp].set) continue;
    bool fused = false;
    TRY(push_species(c, isp, need_sort, c->timing, true, &fused));
    if (isp == last && fused && side_chain) {
      TRY(ensure_side(c));
      CUDA_TRY(cudaEventRecord(c->ev_pfork, c->stream));
      CUDA_TRY(cudaStreamWaitEvent(c->side, c->ev_pfork, 0));
      cudaStream_t lib = c->stream;
      c->stream = c->side;
      int rc = pbcs_species(c, isp, B, fused);
      if (rc == 0) rc = publish_counts(c);
      c->stream = lib;
      if (rc != 0) return rc;
      CUDA_TRY(cudaEventRecord(c->ev_pdone, c->side));
      c->side_pending = true;
      published = true;
    } else {
      TRY(pbcs_species(c, isp, B, fused));
    }
  }
  c->presorted = false;
  if (c->xcap > 0) {
    c->r_clean = true;   // particle_bcs has seen every particle at its new position
    if (!published) TRY(publish_counts(c));
  }
  if (need_sort) {
    c->sorted_valid = true;
    c->pushes_since_sort = 0;
    c->stats.n_sorts += 1;
  }
  c->pushes_since_sort += 1;
  return do_r_min_final(c);
}

// ------------------------------------------------------------------------------------------
// push_particles + particle_bcs for particle lists that live in HOST memory (the reference's
// linked lists flattened in pack_particle order, partlist.F90:414-428): the list is streamed
// through the GPU in chunks on three streams -- upload chunk k+2 | sort + push + deposit +
// boundary conditions of chunk k | download chunk k-1 -- so that the step runs at the PCIe
// rate (both directions busy) instead of upload + compute + download in sequence, and a
// species never has to fit in HBM.  The survivors are written back in place (compacted; the
// output offset never overtakes the input offset), the migrants of all chunks are exchanged
// once at the end and appended.  J accumulates over the chunks exactly as over one list.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_all(Soa s, double* __restrict__ aos, int64_t n) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
#pragma unroll
  for (int q = 0; q < 7; ++q) aos[7 * k + q] = s.d[q][k];
}

static int host_stream_setup(cylgpu_ctx* c, int64_t chunk) {
  cylgpu::HostStream& H = c->hs;
  if (!H.up) {
    CUDA_TRY(cudaStreamCreateWithFlags(&H.up, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&H.down, cudaStreamNonBlocking));
    for (int k = 0; k < 3; ++k) {
      CUDA_TRY(cudaEventCreateWithFlags(&H.ev_up[k], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&H.ev_unpacked[k], cudaEventDisableTiming));
    }
    for (int k = 0; k < 2; ++k) {
      CUDA_TRY(cudaEventCreateWithFlags(&H.ev_packed[k], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&H.ev_down[k], cudaEventDisableTiming));
    }
  }
  if (H.cap < chunk) {
    CUDA_TRY(cudaDeviceSynchronize());
    for (int k = 0; k < 3; ++k) { if (H.in[k]) cudaFree(H.in[k]); H.in[k] = nullptr; }
    for (int k = 0; k < 2; ++k) { if (H.out[k]) cudaFree(H.out[k]); H.out[k] = nullptr; }
    for (int k = 0; k < 3; ++k) CUDA_TRY(cudaMalloc(&H.in[k], (size_t)chunk * 7 * sizeof(double)));
    for (int k = 0; k < 2; ++k) CUDA_TRY(cudaMalloc(&H.out[k], (size_t)chunk * 7 * sizeof(double)));
    H.cap = chunk;
  }
  return 0;
}

int do_push_host(cylgpu_ctx* c, const int64_t* n_in, double* const* host_aos, const int64_t* capacity,
                 int64_t* n_out) {
  TRY(presort_join(c, false));
  cylgpu::HostStream& H = c->hs;
  const int64_t chunk = c->host_chunk;
  TRY(host_stream_setup(c, chunk));
  const BcsConst B = make_bcs_const(c);
  c->stats.n_sent_left = c->stats.n_sent_right = c->stats.n_removed = c->stats.n_recv = 0;
  if (c->host_remove_active) c->stats.n_window_removed = 0;
  TRY(push_prologue(c));
  for (int isp = 0; isp < c->cfg.n_species; ++isp) {
    cylgpu::SpeciesState& S = c->species[isp];
    if (!S.set) continue;
    if (!host_aos[isp]) {   // device-resident species: the ordinary path
      const bool need_sort = c->sort_interval > 0;
      TRY(push_species(c, isp, need_sort, false));
      continue;
    }
    const int64_t n = n_in[isp];
    if (n < 0 || n > capacity[isp]) { set_error("push_host: bad particle count for species %d", isp); return 2; }
    double* host = host_aos[isp];
    const int64_t nchunks = (n + chunk - 1) / chunk;
    TRY(reserve_particles(c, isp, std::min<int64_t>(chunk, std::max<int64_t>(n, 1))));
    // everything queued on the library stream so far (fields, J prologue) precedes the chunks
    int64_t out_off = 0, acc_l = 0, acc_r = 0;
    auto chunk_len = [&](int64_t k) { return std::min<int64_t>(chunk, n - k * chunk); };
    auto enqueue_upload = [&](int64_t k) -> int {
      const int b = (int)(k % 3);
      if (k >= 3) CUDA_TRY(cudaStreamWaitEvent(H.up, H.ev_unpacked[b], 0));   // buffer consumed by chunk k-3
      CUDA_TRY(cudaMemcpyAsync(H.in[b], host + 7 * k * chunk, (size_t)chunk_len(k) * 7 * sizeof(double),
                               cudaMemcpyHostToDevice, H.up));
      CUDA_TRY(cudaEventRecord(H.ev_up[b], H.up));
      return 0;
    };
    for (int64_t k = 0; k < std::min<int64_t>(2, nchunks); ++k) TRY(enqueue_upload(k));
    for (int64_t k = 0; k < nchunks; ++k) {
      if (k + 2 < nchunks) TRY(enqueue_upload(k + 2));
      const int b = (int)(k % 3), ob = (int)(k % 2);
      const int64_t m = chunk_len(k);
      CUDA_TRY(cudaStreamWaitEvent(c->stream, H.ev_up[b], 0));
      Soa s;
      for (int q = 0; q < 7; ++q) s.d[q] = S.d[q];
      k_unpack<<<nblk(m, 256), 256, 0, c->stream>>>(s, 0, H.in[b], m);
      CUDA_TRY(cudaEventRecord(H.ev_unpacked[b], c->stream));
      c->stats.kernel_launches += 1;
      S.n = m;
      bool fused = false;
      c->sort_drops_behind = c->host_remove_active;
      const int prc = push_species(c, isp, c->sort_interval > 0, false, true, &fused);
      if (prc != 0) { c->sort_drops_behind = false; return prc; }
      int64_t nleft = 0, nright = 0;
      const bool dropping = c->sort_drops_behind;   // the window moved since the last push (cylgpu_window_shift)
      if (dropping && !fused) { set_error("push_host: the moving window needs the strip push (variant >= 2, sort every push)"); return 2; }
      TRY(pbcs_classify_compact(c, isp, B, acc_l, acc_r, &nleft, &nright, fused,
                                dropping ? S.cell_start + (S.cell_start_n - 1) : nullptr));   // host sync: counts
      c->sort_drops_behind = false;
      acc_l += nleft;
      acc_r += nright;
      const int64_t kept = S.n;
      if (kept > 0) {
        if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(c->stream, H.ev_down[ob], 0));   // staging buffer free again
        for (int q = 0; q < 7; ++q) s.d[q] = S.d[q];
        k_pack_all<<<nblk(kept, 256), 256, 0, c->stream>>>(s, H.out[ob], kept);
        c->stats.kernel_launches += 1;
        CUDA_TRY(cudaEventRecord(H.ev_packed[ob], c->stream));
        CUDA_TRY(cudaStreamWaitEvent(H.down, H.ev_packed[ob], 0));
        CUDA_TRY(cudaMemcpyAsync(host + 7 * out_off, H.out[ob], (size_t)kept * 7 * sizeof(double),
                                 cudaMemcpyDeviceToHost, H.down));
        CUDA_TRY(cudaEventRecord(H.ev_down[ob], H.down));
        out_off += kept;
      }
    }
    S.n = 0;
    // one exchange for the migrants of all chunks; arrivals are appended right-then-left
    int64_t from_l = 0, from_r = 0;
    TRY(pbcs_exchange(c, acc_l, acc_r, &from_l, &from_r));
    const int64_t nrecv = from_l + from_r;
    if (out_off + nrecv > capacity[isp]) {
      set_error("push_host: species %d needs room for %lld particles, capacity %lld", isp,
                (long long)(out_off + nrecv), (long long)capacity[isp]);
      return 2;
    }
    CUDA_TRY(cudaStreamSynchronize(H.down));
    if (from_r > 0)
      CUDA_TRY(cudaMemcpyAsync(host + 7 * out_off, c->precv + 7 * from_l, (size_t)from_r * 7 * sizeof(double),
                               cudaMemcpyDeviceToHost, c->stream));
    if (from_l > 0)
      CUDA_TRY(cudaMemcpyAsync(host + 7 * (out_off + from_r), c->precv, (size_t)from_l * 7 * sizeof(double),
                               cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->stats.n_recv += nrecv;
    n_out[isp] = out_off + nrecv;
    c->stats.n_particles[isp] = 0;
    TRY(set_count_exact(c, isp));
  }
  c->sorted_valid = false;
  c->host_remove_active = false;
  return do_r_min_final(c);
}

int do_remove_behind(cylgpu_ctx* c) {
  TRY(presort_join(c, false));
  c->stats.n_window_removed = 0;
  if (c->xcap > 0)   // device-resident counts: the removals of this shift are counted on the device
    CUDA_TRY(cudaMemsetAsync(c->n_dev + CYLGPU_MAX_SPECIES + PST_WINDOW_REMOVED, 0, sizeof(int64_t), c->stream));
  if (!c->cfg.x_min_boundary) return 0;
  unsigned long long* cnt = c->counters;
  unsigned long long* cnt2 = c->counters + 8;
  for (int isp = 0; isp < c->cfg.n_species; ++isp) {
    cylgpu::SpeciesState& S = c->species[isp];
    if (!S.set || S.n == 0) continue;
    TRY(reserve_pscratch(c, S.n));
    CUDA_TRY(cudaMemsetAsync(cnt, 0, 8 * sizeof(unsigned long long), c->stream));
    if (c->xcap > 0) {   // rides on the particle_bcs that follows the shift (do_particle_bcs / flush_pending_remove)
      c->pending_remove = true;
      c->pending_remove_x = c->x_min;
      break;
    }
    k_flag_behind<<<nblk(S.n, 256), 256, 0, c->stream>>>(S.d[0], c->x_min, c->hole_list, c->flag, cnt, S.n);
    c->stats.kernel_launches += 1;
    CUDA_TRY(cudaMemcpyAsync(c->h_counters, cnt, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                             c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const int64_t nholes = (int64_t)c->h_counters[CNT_HOLE];
    if (nholes > 0) TRY(compact(c, S, nholes, 0, 0, cnt2));
    c->stats.n_window_removed += nholes;
    c->stats.n_particles[isp] = S.n;
    TRY(set_count_exact(c, isp));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// cell sort (counting sort on the reference's cell index, split_particle.F90:62-63)
// ------------------------------------------------------------------------------------------
struct SortGeom {
  int ncx, ncy;   // nx + 2*CELL_PAD, ny + 2*CELL_PAD
  double x_grid_min_local, y_grid_min_local, dx, dy;
  // prediction of the upcoming half-step position (particles.F90:303-309)
  double ipart_mc, dtco2, idx, idy;
  // particles behind this x go to the extra bucket behind all cells: no strip owns it, so they are neither pushed
  // nor kept (remove_particles, window.F90:304-325, for lists that are streamed from host memory)
  double x_remove;
};

__device__ __forceinline__ void ref_cell(const SortGeom& G, double x, double y, double z, int& cx, int& cy) {
  const double r = sqrt(y * y + z * z);
#if CYL_SHAPE == 1   // split_particle.F90:57-63: the top-hat counts from the cell edge
  cx = (int)floor((x - G.x_grid_min_local) / G.dx) + 1;
  cy = (int)floor((r - G.y_grid_min_local) / G.dy) + 1;
#else
  cx = (int)floor((x - G.x_grid_min_local) / G.dx + 1.5);
  cy = (int)floor((r - G.y_grid_min_local) / G.dy + 1.5);
#endif
}

// The sort bucket is the STAGGERED cell (cell_x2, cell_y2) the particle will have after the
// half-step drift of the next push (particles.F90:303-309,369-374): every particle of a
// bucket gathers from the same field nodes and deposits into the same 5x5 node window,
// which is what the warp-window deposit needs.  Same arithmetic as push_one(); a rare
// last-bit disagreement only costs the fallback path, never correctness.
__device__ __forceinline__ void stag_cell(const SortGeom& G, double x, double y, double z, double px, double py,
                                          double pz, int& cx2, int& cy2) {
  const double ux = px * G.ipart_mc, uy = py * G.ipart_mc, uz = pz * G.ipart_mc;
  const double root = G.dtco2 / sqrt(ux * ux + uy * uy + uz * uz + 1.0);
  x = x + ux * root;
  y = y + uy * root;
  z = z + uz * root;
  const double r = sqrt(y * y + z * z);
  cx2 = (int)floor((x - G.x_grid_min_local) * G.idx) + 1;
  cy2 = (int)floor((r - G.y_grid_min_local) * G.idy) + 1;
}

__device__ __forceinline__ int cell_key(const SortGeom& G, int cx, int cy) {
  int kx = cx - 1 + CELL_PAD, ky = cy - 1 + CELL_PAD;
  kx = max(0, min(G.ncx - 1, kx));
  ky = max(0, min(G.ncy - 1, ky));
  return ky * G.ncx + kx;
}

__global__ void __launch_bounds__(256) k_sort_hist(SortGeom G, const double* __restrict__ x,
                                                   const double* __restrict__ y, const double* __restrict__ z,
                                                   const double* __restrict__ px, const double* __restrict__ py,
                                                   const double* __restrict__ pz, int* __restrict__ count,
                                                   uint32_t* __restrict__ key, uint32_t* __restrict__ rank,
                                                   int64_t n, const int64_t* __restrict__ n_dev) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = *n_dev;   // device-resident count: the launch covers an upper bound
  int k = -1;
  if (i < n) {
    int cx, cy;
    const double X = x[i];
    stag_cell(G, X, y[i], z[i], px[i], py[i], pz[i], cx, cy);
    k = (X < G.x_remove) ? G.ncx * G.ncy : cell_key(G, cx, cy);
  }
  // the list is almost sorted from the previous step, so the lanes of a warp share a few
  // buckets: one atomic per distinct bucket per warp instead of one per particle
  const unsigned act = __ballot_sync(0xffffffffu, i < n);
  if (i >= n) return;
  const unsigned peers = __match_any_sync(act, k);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(&count[k], __popc(peers));
  base = __shfl_sync(peers, base, leader);
  key[i] = (uint32_t)k;
  rank[i] = (uint32_t)(base + __popc(peers & ((1u << lane) - 1u)));
}

// exclusive scan, 3 kernels; SCAN_B elements per block
#define SCAN_B 1024
__global__ void __launch_bounds__(SCAN_B) k_scan_block(int* __restrict__ a, int* __restrict__ block_sum, int64_t n) {
  __shared__ int sh[SCAN_B];
  const int64_t i = (int64_t)blockIdx.x * SCAN_B + threadIdx.x;
  const int v = (i < n) ? a[i] : 0;
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int off = 1; off < SCAN_B; off <<= 1) {
    int t = 0;
    if ((int)threadIdx.x >= off) t = sh[threadIdx.x - off];
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  if (i < n) a[i] = sh[threadIdx.x] - v;   // exclusive
  if (threadIdx.x == SCAN_B - 1) block_sum[blockIdx.x] = sh[threadIdx.x];
}
__global__ void __launch_bounds__(SCAN_B) k_scan_sums(int* __restrict__ block_sum, int nb) {
  __shared__ int sh[SCAN_B];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += SCAN_B) {
    const int i = base + threadIdx.x;
    const int v = (i < nb) ? block_sum[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < SCAN_B; off <<= 1) {
      int t = 0;
      if ((int)threadIdx.x >= off) t = sh[threadIdx.x - off];
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nb) block_sum[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == SCAN_B - 1) carry += sh[threadIdx.x];
    __syncthreads();
  }
}
__global__ void __launch_bounds__(SCAN_B) k_scan_add(int* __restrict__ a, const int* __restrict__ block_sum,
                                                     int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * SCAN_B + threadIdx.x;
  if (i < n) a[i] += block_sum[blockIdx.x];
}

__global__ void __launch_bounds__(256) k_sort_dest(const int* __restrict__ start, const uint32_t* __restrict__ key,
                                                   uint32_t* __restrict__ rank_to_dest, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rank_to_dest[i] = (uint32_t)start[key[i]] + rank_to_dest[i];
}

// sorted slot -> particle index (the strip push gathers through it; no data moves here)
__global__ void __launch_bounds__(256) k_sort_src(const int* __restrict__ start, const uint32_t* __restrict__ key,
                                                  const uint32_t* __restrict__ rank, uint32_t* __restrict__ src,
                                                  int64_t n, const int64_t* __restrict__ n_dev) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = *n_dev;
  if (i >= n) return;
  src[(uint32_t)start[key[i]] + rank[i]] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) k_scatter(const double* __restrict__ src, double* __restrict__ dst,
                                                 const uint32_t* __restrict__ dest, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dst[dest[i]] = src[i];
}

// One species.  physical = true moves the particle data into bucket order (7 scatter passes
// through one rotating spare array); physical = false only builds the permutation `c->perm`
// (sorted slot -> particle) and the bucket starts, and makes sure the second buffer set
// exists: the strip push kernel gathers through the permutation and writes the sorted list.
int do_sort_species(cylgpu_ctx* c, int isp, bool physical) {
  SortGeom G;
  G.ncx = c->g.nx + 2 * CELL_PAD;
  G.ncy = c->g.ny + 2 * CELL_PAD;
  G.x_grid_min_local = c->x_grid_min_local;
  G.y_grid_min_local = c->cfg.y_grid_min_local;
  G.dx = c->cfg.dx; G.dy = c->cfg.dy;
  G.idx = 1.0 / c->cfg.dx; G.idy = 1.0 / c->cfg.dy;
  G.x_remove = (c->sort_drops_behind && !physical) ? c->host_remove_x : -1.0e300;   // do_push_host only
  const int64_t ncell = (int64_t)G.ncx * G.ncy;
  // one extra bucket behind the cells: empty unless x_remove is set; its exclusive-scan value is the number of
  // particles the strips own
  const int64_t nscan = ncell + 1;
  const int nb = (int)((nscan + SCAN_B - 1) / SCAN_B);
  if (!c->scan_blocks || c->ncell != ncell) {
    if (c->scan_blocks) cudaFree(c->scan_blocks);
    CUDA_TRY(cudaMalloc(&c->scan_blocks, (size_t)(nb + 1) * sizeof(int)));
    c->ncell = ncell;
  }
  cylgpu::SpeciesState& S = c->species[isp];
  if (physical && S.lazy) TRY(poll_counts(c, true));   // the scatter passes take the exact count
  if (!S.set || S.n == 0 || S.sp.immobile) return 0;
  if (S.n >= (int64_t)0x7FFFFFFFLL) { set_error("more than 2^31-1 particles per species per GPU"); return 3; }
  TRY(reserve_pscratch(c, S.cap));
  // per-species bucket starts: the strip push kernel reads them after the sort
  if (!S.cell_start || S.cell_start_n != nscan) {
    if (S.cell_start) cudaFree(S.cell_start);
    CUDA_TRY(cudaMalloc(&S.cell_start, (size_t)nscan * sizeof(int)));
    S.cell_start_n = nscan;
  }
  int* count = S.cell_start;
  uint32_t* key = c->hole_list;    // scratch reuse: the particle_bcs lists are idle during a sort
  uint32_t* rank = c->lowhole;
  if (!physical && S.perm_cap < S.cap) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (S.perm) CUDA_TRY(cudaFree(S.perm));
    S.perm = nullptr;
    CUDA_TRY(cudaMalloc(&S.perm, (size_t)S.cap * sizeof(uint32_t)));
    S.perm_cap = S.cap;
  }
  uint32_t* perm = S.perm;
  CUDA_TRY(cudaMemsetAsync(count, 0, (size_t)nscan * sizeof(int), c->stream));
  G.ipart_mc = 1.0 / (C_LIGHT * S.sp.mass);
  G.dtco2 = C_LIGHT * (c->dt / 2.0);
  const int64_t* n_dev = S.lazy ? c->n_dev + isp : nullptr;
  k_sort_hist<<<nblk(S.n, 256), 256, 0, c->stream>>>(G, S.d[0], S.d[1], S.d[2], S.d[3], S.d[4], S.d[5],
                                                    count, key, rank, S.n, n_dev);
  k_scan_block<<<nb, SCAN_B, 0, c->stream>>>(count, c->scan_blocks, nscan);
  k_scan_sums<<<1, SCAN_B, 0, c->stream>>>(c->scan_blocks, nb);
  k_scan_add<<<nb, SCAN_B, 0, c->stream>>>(count, c->scan_blocks, nscan);
  c->stats.kernel_launches += 4;
  if (!physical) {
    k_sort_src<<<nblk(S.n, 256), 256, 0, c->stream>>>(count, key, rank, perm, S.n, n_dev);
    c->stats.kernel_launches += 1;
    if (S.alt_cap < S.cap) {
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      for (int q = 0; q < 7; ++q) {
        if (S.alt[q]) CUDA_TRY(cudaFree(S.alt[q]));
        S.alt[q] = nullptr;
        CUDA_TRY(cudaMalloc(&S.alt[q], (size_t)S.cap * sizeof(double)));
      }
      S.alt_cap = S.cap;
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
  }
  k_sort_dest<<<nblk(S.n, 256), 256, 0, c->stream>>>(count, key, rank, S.n);   // rank -> destination
  c->stats.kernel_launches += 1;
  // scatter each component into the spare array, then rotate the pointers (the old
  // component array becomes the next spare): 8 B of scratch per particle, not 56
  if (c->ptmp_cap < S.cap) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->ptmp) cudaFree(c->ptmp);
    CUDA_TRY(cudaMalloc(&c->ptmp, (size_t)S.cap * sizeof(double)));
  }
  for (int q = 0; q < 7; ++q) {
    k_scatter<<<nblk(S.n, 256), 256, 0, c->stream>>>(S.d[q], c->ptmp, rank, S.n);
    c->stats.kernel_launches += 1;
    double* t = S.d[q];
    S.d[q] = c->ptmp;
    c->ptmp = t;
  }
  c->ptmp_cap = S.cap;   // the spare is now one of this species' old arrays
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int do_sort(cylgpu_ctx* c) {
  for (int isp = 0; isp < c->cfg.n_species; ++isp) TRY(do_sort_species(c, isp, true));
  c->sorted_valid = true;
  c->pushes_since_sort = 0;
  c->stats.n_sorts += 1;
  return 0;
}

// ------------------------------------------------------------------------------------------
// diagnostics
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cells(SortGeom G, const double* __restrict__ x, const double* __restrict__ y,
                                               const double* __restrict__ z, int32_t* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cx, cy;
  ref_cell(G, x[i], y[i], z[i], cx, cy);
  out[2 * i] = cx;
  out[2 * i + 1] = cy;
}

int do_cells(cylgpu_ctx* c, int isp, int64_t capn, int32_t* out) {
  cylgpu::SpeciesState& S = c->species[isp];
  if (capn < S.n) { set_error("cells_out too small"); return 2; }
  if (S.n == 0) return 0;
  SortGeom G;
  G.ncx = G.ncy = 0;
  G.x_grid_min_local = c->x_grid_min_local;
  G.y_grid_min_local = c->cfg.y_grid_min_local;
  G.dx = c->cfg.dx; G.dy = c->cfg.dy;
  G.x_remove = -1.0e300;
  int32_t* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, (size_t)S.n * 2 * sizeof(int32_t)));
  k_cells<<<nblk(S.n, 256), 256, 0, c->stream>>>(G, S.d[0], S.d[1], S.d[2], d, S.n);
  c->stats.kernel_launches += 1;
  CUDA_TRY(cudaMemcpyAsync(out, d, (size_t)S.n * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaFree(d));
  return 0;
}

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[8];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 8) t = sh[threadIdx.x];
  if (threadIdx.x < 32) for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __syncthreads();
  return t;
}

// field energy: theta-average of (Re sum_m F_m e^{-im theta})^2 = Re(F_0)^2 + 1/2 sum_{m>0} |F_m|^2,
// per cell i=1..nx, j=1..ny with volume 2 pi r dx dy, staggered components averaged to the centre
__global__ void __launch_bounds__(256) k_field_energy(Geom g, const cplx* __restrict__ exm, const cplx* __restrict__ erm,
                                                      const cplx* __restrict__ etm, const cplx* __restrict__ bxm,
                                                      const cplx* __restrict__ brm, const cplx* __restrict__ btm,
                                                      double dx, double dy, double y_grid_min_local,
                                                      double* __restrict__ out) {
  const int64_t ncell = (int64_t)g.nx * g.ny;
  double acc = 0.0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ncell * g.M;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int im = (int)(t / ncell);
    const int64_t q = t % ncell;
    const int j = (int)(q / g.nx) + 1, i = (int)(q % g.nx) + 1;
    const double r = y_grid_min_local + (double)(j - 1) * dy;
    auto sq = [&](cplx v) { return (im == 0) ? v.x * v.x : 0.5 * (v.x * v.x + v.y * v.y); };
    const size_t o = g.at(i, j, im);
    const size_t SX = g.SX;
    double e2 = 0.5 * (sq(exm[o]) + sq(exm[o - SX]))                      // Exm: r-staggered
              + 0.5 * (sq(erm[o]) + sq(erm[o - 1]))                       // Erm: x-staggered
              + 0.25 * (sq(etm[o]) + sq(etm[o - 1]) + sq(etm[o - SX]) + sq(etm[o - SX - 1]));
    double b2 = 0.5 * (sq(bxm[o]) + sq(bxm[o - 1]))
              + 0.5 * (sq(brm[o]) + sq(brm[o - SX]))
              + sq(btm[o]);
    acc += 0.5 * EPSILON0 * (e2 + C_LIGHT * C_LIGHT * b2) * (2.0 * PI * r * dx * dy);
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

__global__ void __launch_bounds__(256) k_kinetic_energy(const double* __restrict__ px, const double* __restrict__ py,
                                                        const double* __restrict__ pz, const double* __restrict__ w,
                                                        double mass, int64_t n, double* __restrict__ out) {
  double acc = 0.0;
  const double mc = mass * C_LIGHT;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ux = px[i] / mc, uy = py[i] / mc, uz = pz[i] / mc;
    const double u2 = ux * ux + uy * uy + uz * uz;
    const double gm1 = u2 / (sqrt(1.0 + u2) + 1.0);   // gamma - 1 without cancellation
    acc += w[i] * gm1 * mass * C_LIGHT * C_LIGHT;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

int do_energy(cylgpu_ctx* c, double* out2) {
  const Geom& g = c->g;
  CUDA_TRY(cudaMemsetAsync(c->d_energy, 0, 2 * sizeof(double), c->stream));
  k_field_energy<<<148 * 4, 256, 0, c->stream>>>(g, c->f[CYLGPU_EXM], c->f[CYLGPU_ERM], c->f[CYLGPU_ETM],
                                                 c->f[CYLGPU_BXM], c->f[CYLGPU_BRM], c->f[CYLGPU_BTM], c->cfg.dx,
                                                 c->cfg.dy, c->cfg.y_grid_min_local, c->d_energy);
  c->stats.kernel_launches += 1;
  for (int isp = 0; isp < c->cfg.n_species; ++isp) {
    cylgpu::SpeciesState& S = c->species[isp];
    if (!S.set || S.n == 0) continue;
    k_kinetic_energy<<<148 * 4, 256, 0, c->stream>>>(S.d[3], S.d[4], S.d[5], S.d[6], S.sp.mass, S.n,
                                                     c->d_energy + 1);
    c->stats.kernel_launches += 1;
  }
  CUDA_TRY(cudaMemcpyAsync(out2, c->d_energy, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}

}  // namespace cylgpu
