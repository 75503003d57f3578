// moments_kernels.cuh -- device kernels of the calc_df.F90 particle moments (host side: moments.cuh).
// Kernels only, included inside namespace cylgpu; also compiled for the CPU by the kernel-emulation
// tests (tests/emul/).  Product code: no oracle here.
#pragma once

constexpr double KB = 1.3806488e-23;   // constants.F90:176

struct MomentArgs {
  const double *x, *y, *z, *px, *py, *pz, *w;
  int64_t n;
  double x_grid_min_local, y_grid_min_local, dx, dy;
  double mass, charge;
  int kind, direction;
};

#if CYL_SHAPE == 0
#define GW0 1        // index of offset 0 in the weight arrays below
#else
#define GW0 WO
#endif
struct ToGrid {   // include/particle_to_grid.inc + <shape>/gxfac.inc
  int cell_x, cell_y;
#if CYL_SHAPE == 0
  double gx[3], gy[3];
#else
  double gx[NWT], gy[NWT];
#endif
  double part_r;
};

__device__ __forceinline__ ToGrid particle_to_grid(const MomentArgs& a, int64_t i) {
  ToGrid t;
  const double Y = a.y[i], Z = a.z[i];
  t.part_r = sqrt(Y * Y + Z * Z);
#if CYL_SHAPE != 0
  shape_particle_to_grid(a.x[i] - a.x_grid_min_local, t.part_r - a.y_grid_min_local, t.part_r, a.dx, a.dy, &t.cell_x,
                         &t.cell_y, t.gx, t.gy);
  return t;
#else
  const double cell_x_r = (a.x[i] - a.x_grid_min_local) / a.dx;
  const double cell_y_r = (t.part_r - a.y_grid_min_local) / a.dy;
  t.cell_x = (int)floor(cell_x_r + 0.5);
  t.cell_y = (int)floor(cell_y_r + 0.5);
  const double cell_frac_x = (double)t.cell_x - cell_x_r;
  const double cell_frac_y = (double)t.cell_y - cell_y_r;
  t.cell_x += 1;
  t.cell_y += 1;
  const double cx2 = cell_frac_x * cell_frac_x;
  t.gx[0] = 0.5 * (0.25 + cx2 + cell_frac_x);
  t.gx[1] = 0.75 - cx2;
  t.gx[2] = 0.5 * (0.25 + cx2 - cell_frac_x);
  const double cy2 = cell_frac_y * cell_frac_y;
  t.gy[0] = 0.5 * (0.25 + cy2 + cell_frac_y);
  t.gy[1] = 0.75 - cy2;
  t.gy[2] = 0.5 * (0.25 + cy2 - cell_frac_y);
  if (t.part_r < a.dy) {
    t.gy[1] = t.gy[1] + t.gy[0];
    t.gy[0] = 0.0;
  }
  return t;
#endif
}

// mass density, number density, per-species current (re only); ekbar, ekflux, average momentum
// (re = data, im = wt / part_count)
__global__ void __launch_bounds__(256) k_moment_deposit(Geom g, MomentArgs a, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const ToGrid t = particle_to_grid(a, i);
  const double part_w = a.w[i];
  const double c = C_LIGHT;
  const double macro_part_volume = 2.0 * PI * a.dx * a.dy * t.part_r;   // partlist.F90:999-1013
  double wdata = 0.0;
  bool averaged = false;
  switch (a.kind) {
    case CYLGPU_MOM_MASS_DENSITY:
      wdata = (a.mass * part_w) / macro_part_volume;
      break;
    case CYLGPU_MOM_NUMBER_DENSITY:
      wdata = part_w / macro_part_volume;
      break;
    case CYLGPU_MOM_SPECIES_CURRENT: {
      const double part_mc = c * a.mass;
      const double px = a.px[i], py = a.py[i], pz = a.pz[i];
      const double root = 1.0 / sqrt(part_mc * part_mc + px * px + py * py + pz * pz);
      const double pd = a.direction == 1 ? px : (a.direction == 2 ? py : pz);
      wdata = (a.charge * part_w) * pd * root;
      wdata = wdata * c / macro_part_volume;
    } break;
    case CYLGPU_MOM_EKBAR:
    case CYLGPU_MOM_EKFLUX: {
      averaged = true;
      const double part_mc = c * a.mass;
      const double fac = part_mc * part_w * c;
      const double part_ux = a.px[i] / part_mc, part_uy = a.py[i] / part_mc, part_uz = a.pz[i] / part_mc;
      const double part_u2 = part_ux * part_ux + part_uy * part_uy + part_uz * part_uz;
      const double gamma_rel = sqrt(part_u2 + 1.0);
      const double gamma_rel_m1 = part_u2 / (gamma_rel + 1.0);
      wdata = gamma_rel_m1 * fac;
      if (a.kind == CYLGPU_MOM_EKFLUX && a.direction != 0) {
        const int d = a.direction < 0 ? -a.direction : a.direction;
        const double f = d == 1 ? c * a.dy : (d == 2 ? c * a.dx : c * a.dx * a.dy);   // xfac, yfac, zfac :275-277
        const double u = d == 1 ? part_ux : (d == 2 ? part_uy : part_uz);
        const double part_flux = f * u / gamma_rel;
        wdata = a.direction < 0 ? -wdata * fmin(part_flux, 0.0) : wdata * fmax(part_flux, 0.0);
      }
    } break;
    case CYLGPU_MOM_AVERAGE_MOMENTUM: {
      averaged = true;
      const double pd = a.direction == 1 ? a.px[i] : (a.direction == 2 ? a.py[i] : a.pz[i]);
      wdata = part_w * pd;
    } break;
    default: return;
  }
#pragma unroll
  for (int iy = SF_MIN; iy <= SF_MAX; ++iy)
#pragma unroll
    for (int ix = SF_MIN; ix <= SF_MAX; ++ix) {
      const double gg = t.gx[ix + GW0] * t.gy[iy + GW0];
      if (gg == 0.0) continue;
      const size_t o = 2 * g.at(t.cell_x + ix, t.cell_y + iy, 0);
      atomicAdd(out + o, gg * wdata);
      if (averaged) atomicAdd(out + o + 1, gg * part_w);
    }
}

// calc_ppc / calc_average_weight: nearest cell, no shape function, no boundary pass
__global__ void __launch_bounds__(256) k_moment_count(Geom g, MomentArgs a, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const double Y = a.y[i], Z = a.z[i];
  const double part_r = sqrt(Y * Y + Z * Z);
  // (top-hat: without the half cell, calc_df.F90:696-702, 757-763)
  const double cell_x_r = (a.x[i] - a.x_grid_min_local) / a.dx + (0.5 - SHAPE_CELL_SHIFT);
  const double cell_y_r = (part_r - a.y_grid_min_local) / a.dy + (0.5 - SHAPE_CELL_SHIFT);
  const int cell_x = (int)floor(cell_x_r) + 1;
  const int cell_y = (int)floor(cell_y_r) + 1;
  const size_t o = 2 * g.at(cell_x, cell_y, 0);
  if (a.kind == CYLGPU_MOM_PPC) {
    atomicAdd(out + o, 1.0);
  } else {
    atomicAdd(out + o, a.w[i]);
    atomicAdd(out + o + 1, 1.0);
  }
}

// calc_temperature, first pass (:838-901): A = (meanx, meany), B = (meanz, part_count)
__global__ void __launch_bounds__(256) k_temperature_means(Geom g, MomentArgs a, double* __restrict__ A,
                                                           double* __restrict__ B) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const ToGrid t = particle_to_grid(a, i);
  const double sqrt_part_m = sqrt(a.mass);
  const double part_w = a.w[i];
  const double pmx = a.px[i] / sqrt_part_m, pmy = a.py[i] / sqrt_part_m, pmz = a.pz[i] / sqrt_part_m;
  const int dir = a.direction;
#pragma unroll
  for (int iy = SF_MIN; iy <= SF_MAX; ++iy)
#pragma unroll
    for (int ix = SF_MIN; ix <= SF_MAX; ++ix) {
      const double gf = t.gx[ix + GW0] * t.gy[iy + GW0] * part_w;
      if (gf == 0.0) continue;
      const size_t o = 2 * g.at(t.cell_x + ix, t.cell_y + iy, 0);
      if (dir <= 0 || dir == 1) atomicAdd(A + o, gf * pmx);
      if (dir <= 0 || dir == 2) atomicAdd(A + o + 1, gf * pmy);
      if (dir <= 0 || dir == 3) atomicAdd(B + o, gf * pmz);
      atomicAdd(B + o + 1, gf);
    }
}

// part_count = MAX(part_count, 1e-6); mean* = mean* / part_count on the whole array (:954-958)
__global__ void __launch_bounds__(256) k_temperature_normalise(cplx* __restrict__ A, cplx* __restrict__ B, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  cplx a = A[i], b = B[i];
  const double pc = fmax(b.y, 1.e-6);
  a.x = a.x / pc;
  a.y = a.y / pc;
  b.x = b.x / pc;
  b.y = pc;
  A[i] = a;
  B[i] = b;
}

// second pass (:977-1023): D = (sigma, part_count) with the unweighted shape factors
__global__ void __launch_bounds__(256) k_temperature_sigma(Geom g, MomentArgs a, const cplx* __restrict__ A,
                                                           const cplx* __restrict__ B, double* __restrict__ D) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const ToGrid t = particle_to_grid(a, i);
  const double sqrt_part_m = sqrt(a.mass);
  const double pmx = a.px[i] / sqrt_part_m, pmy = a.py[i] / sqrt_part_m, pmz = a.pz[i] / sqrt_part_m;
  const int dir = a.direction;
#pragma unroll
  for (int iy = SF_MIN; iy <= SF_MAX; ++iy)
#pragma unroll
    for (int ix = SF_MIN; ix <= SF_MAX; ++ix) {
      const double gf = t.gx[ix + GW0] * t.gy[iy + GW0];
      if (gf == 0.0) continue;
      const size_t o = g.at(t.cell_x + ix, t.cell_y + iy, 0);
      const cplx ma = A[o], mb = B[o];
      const double ddx = pmx - ma.x, ddy = pmy - ma.y, ddz = pmz - mb.x;
      double wdata;
      if (dir == 1) wdata = ddx * ddx;
      else if (dir == 2) wdata = ddy * ddy;
      else if (dir == 3) wdata = ddz * ddz;
      else wdata = ddx * ddx + ddy * ddy + ddz * ddz;
      atomicAdd(D + 2 * o, gf * wdata);
      atomicAdd(D + 2 * o + 1, gf);
    }
}

// mode 0: out = re; 1: out = re / MAX(im, c_tiny); 2: out = re / MAX(im, 1e-6) / kb / dof
__global__ void __launch_bounds__(256) k_moment_finish(const cplx* __restrict__ a, double* __restrict__ out, size_t n,
                                                       int mode, double dof) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const cplx v = a[i];
  double r = v.x;
  if (mode == 1) r = v.x / fmax(v.y, DBL_MIN);
  else if (mode == 2) r = v.x / fmax(v.y, 1.e-6) / KB / dof;
  out[i] = r;
}
