// bc_kernels.cuh -- the boundary kernels shared by the field / current exchange (bcs.cu) and the
// particle moments (moments.cuh): packed x halo, real-variant reflection, zero-gradient fill.
// Kernels only (no host code), included inside namespace cylgpu; also compiled for the CPU by the
// kernel-emulation tests (tests/emul/).  Product code: no oracle here.
#pragma once

// ---- field ghost fill at the domain edges: efield_bcs / bfield_bcs, boundary.F90:1355-1476 ----
enum { OP_NONE = 0, OP_CLAMP = 1, OP_ZEROGRAD = 2 };

struct Tri {
  cplx* f[3];
  int op[3];
  int stag[3];   // stagger of each array in the direction normal to the boundary
};

// x_min / x_max ghost fill: boundary.F90:772-829 (clamp) and :654-707 (zero gradient).
// One thread per (row j, mode, array).
__global__ void __launch_bounds__(128) k_edge_x(Geom g, Tri t, int side) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x + 1 - NG;
  if (j > g.ny + NG) return;
  const int im = blockIdx.y, k = blockIdx.z;
  const int op = t.op[k];
  if (op == OP_NONE) return;
  cplx* f = t.f[k];
  const double s = (op == OP_CLAMP) ? -1.0 : 1.0;
  if (side == CYLGPU_BD_X_MIN) {
    if (t.stag[k]) {
      for (int i = 1; i <= NG - 1; ++i) f[g.at(i - NG, j, im)] = s * f[g.at(NG - i, j, im)];
      if (op == OP_CLAMP) f[g.at(0, j, im)] = C(0.0, 0.0);
    } else {
      for (int i = 1; i <= NG; ++i) f[g.at(i - NG, j, im)] = s * f[g.at(NG + 1 - i, j, im)];
    }
  } else {
    const int nn = g.nx;
    if (t.stag[k]) {
      if (op == OP_CLAMP) f[g.at(nn, j, im)] = C(0.0, 0.0);
      for (int i = 1; i <= NG - 1; ++i) f[g.at(nn + i, j, im)] = s * f[g.at(nn - i, j, im)];
    } else {
      for (int i = 1; i <= NG; ++i) f[g.at(nn + i, j, im)] = s * f[g.at(nn + 1 - i, j, im)];
    }
  }
}

// r_max ghost fill; one thread per (column ix, mode, array)
__global__ void __launch_bounds__(128) k_edge_y(Geom g, Tri t) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x + 1 - NG;
  if (ix > g.nx + NG) return;
  const int im = blockIdx.y, k = blockIdx.z;
  const int op = t.op[k];
  if (op == OP_NONE) return;
  cplx* f = t.f[k];
  const double s = (op == OP_CLAMP) ? -1.0 : 1.0;
  const int nn = g.ny;
  if (t.stag[k]) {
    if (op == OP_CLAMP) f[g.at(ix, nn, im)] = C(0.0, 0.0);
    for (int i = 1; i <= NG - 1; ++i) f[g.at(ix, nn + i, im)] = s * f[g.at(ix, nn - i, im)];
  } else {
    for (int i = 1; i <= NG; ++i) f[g.at(ix, nn + i, im)] = s * f[g.at(ix, nn + 1 - i, im)];
  }
}

// ---- x halo: boundary.F90:158-169,500-553.  Buffer layout [comp][im][row][NG]. ----
struct Halo3 {
  cplx* f[3];
  int skip[3];   // rows excluded at the top (the one-row shift of the r-staggered arrays,
                 // boundary.F90:1362-1371,1428-1430: row ny+ng never takes part)
};

// mode 0: pack interior edge columns (send_l <- 1..ng, send_r <- nx+1-ng..nx)
// mode 1: pack ghost columns          (send_l <- 1-ng..0, send_r <- nx+1..nx+ng)  [J sums]
// mode 2: both, ghost block first then interior block (`half` elements apart)  [J sum + J halo in one message]
__global__ void __launch_bounds__(128) k_halo_pack(Geom g, Halo3 h, cplx* __restrict__ send_l,
                                                   cplx* __restrict__ send_r, int mode, size_t half) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = g.SY * NG;
  if (t >= per) return;
  const int im = blockIdx.y, k = blockIdx.z;
  const int row = t / NG, i = t % NG + 1;   // i = 1..ng
  const int j = row + 1 - NG;
  if (j > g.ny + NG - h.skip[k]) return;
  const size_t b = ((size_t)(k * g.M + im) * g.SY + row) * NG + (i - 1);
  const cplx* f = h.f[k];
  if (!f) return;   // exchanges of fewer than three arrays
  if (mode == 0) {
    if (send_l) send_l[b] = f[g.at(i, j, im)];
    if (send_r) send_r[b] = f[g.at(g.nx - NG + i, j, im)];
  } else {
    if (send_l) send_l[b] = f[g.at(i - NG, j, im)];
    if (send_r) send_r[b] = f[g.at(g.nx + i, j, im)];
    if (mode == 2) {
      if (send_l) send_l[half + b] = f[g.at(i, j, im)];
      if (send_r) send_r[half + b] = f[g.at(g.nx - NG + i, j, im)];
    }
  }
}

// mode 0: ghost <- received (recv_l -> 1-ng..0, recv_r -> nx+1..nx+ng)
// mode 1: interior += received (recv_l -> 1..ng, recv_r -> nx+1-ng..nx), boundary.F90:1192,1200
// mode 2: mode 1, and ghost <- the neighbour's interior edge AFTER its own sum, which is its
//         interior block + my ghost value (the very two operands the neighbour adds)
__global__ void __launch_bounds__(128) k_halo_unpack(Geom g, Halo3 h, const cplx* __restrict__ recv_l,
                                                     const cplx* __restrict__ recv_r, int mode, size_t half) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = g.SY * NG;
  if (t >= per) return;
  const int im = blockIdx.y, k = blockIdx.z;
  const int row = t / NG, i = t % NG + 1;
  const int j = row + 1 - NG;
  if (j > g.ny + NG - h.skip[k]) return;
  const size_t b = ((size_t)(k * g.M + im) * g.SY + row) * NG + (i - 1);
  cplx* f = h.f[k];
  if (!f) return;
  if (mode == 0) {
    if (recv_l) f[g.at(i - NG, j, im)] = recv_l[b];
    if (recv_r) f[g.at(g.nx + i, j, im)] = recv_r[b];
  } else {
    if (recv_l) { const size_t o = g.at(i, j, im); f[o] = f[o] + recv_l[b]; }
    if (recv_r) { const size_t o = g.at(g.nx - NG + i, j, im); f[o] = f[o] + recv_r[b]; }
    if (mode == 2) {
      if (recv_l) { const size_t o = g.at(i - NG, j, im); f[o] = recv_l[half + b] + f[o]; }
      if (recv_r) { const size_t o = g.at(g.nx + i, j, im); f[o] = recv_r[half + b] + f[o]; }
    }
  }
}

// particle_reflection_bcs, the real-valued variant (boundary.F90:833-914, flip_direction absent):
// bd 0 x_min (ng-1 ghost columns fold onto 1..ng-1), 1 x_max, 3 r_max (ng ghosts each)
__global__ void __launch_bounds__(128) k_density_reflect(Geom g, cplx* __restrict__ a, int bd) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int im = blockIdx.y;
  if (bd == CYLGPU_BD_Y_MAX) {
    if (t >= g.SX) return;
    const int ix = t + 1 - NG;
    for (int i = 1; i <= NG; ++i) {
      const size_t in = g.at(ix, g.ny + 1 - i, im), gh = g.at(ix, g.ny + i, im);
      a[in] = a[in] + a[gh];
      a[gh] = C(0.0, 0.0);
    }
    return;
  }
  if (t >= g.SY) return;
  const int j = t + 1 - NG;
  if (bd == CYLGPU_BD_X_MIN) {
    for (int i = 1; i <= NG - 1; ++i) {
      const size_t in = g.at(i, j, im), gh = g.at(1 - i, j, im);
      a[in] = a[in] + a[gh];
      a[gh] = C(0.0, 0.0);
    }
  } else {
    for (int i = 1; i <= NG; ++i) {
      const size_t in = g.at(g.nx + 1 - i, j, im), gh = g.at(g.nx + i, j, im);
      a[in] = a[in] + a[gh];
      a[gh] = C(0.0, 0.0);
    }
  }
}

// field_mode_zero_gradient for a cell-centred array (boundary.F90:654-707): ghost i <- interior mirror
__global__ void __launch_bounds__(128) k_density_zero_gradient(Geom g, cplx* __restrict__ a, int bd) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int im = blockIdx.y;
  if (bd == CYLGPU_BD_X_MIN || bd == CYLGPU_BD_X_MAX) {
    if (t >= g.SY) return;
    const int j = t + 1 - NG;
    for (int i = 1; i <= NG; ++i) {
      if (bd == CYLGPU_BD_X_MIN) a[g.at(i - NG, j, im)] = a[g.at(NG + 1 - i, j, im)];
      else a[g.at(g.nx + i, j, im)] = a[g.at(g.nx + 1 - i, j, im)];
    }
  } else {
    if (t >= g.SX) return;
    const int ix = t + 1 - NG;
    for (int i = 1; i <= NG; ++i) {
      if (bd == CYLGPU_BD_Y_MIN) a[g.at(ix, i - NG, im)] = a[g.at(ix, NG + 1 - i, im)];
      else a[g.at(ix, g.ny + i, im)] = a[g.at(ix, g.ny + 1 - i, im)];
    }
  }
}

__global__ void __launch_bounds__(256) k_real_part(const cplx* __restrict__ a, double* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i].x;
}

// ---- currents ----
// current_bcs_r_min_final, boundary.F90:1909-1959: fold the rows below the axis back and set
// the axis rows.  One thread per column; rows are independent between columns.
__global__ void __launch_bounds__(128) k_r_min_final(Geom g, cplx* __restrict__ jxm, cplx* __restrict__ jrm,
                                                     cplx* __restrict__ jtm) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x + 1 - NG;
  const int im = blockIdx.y;
  if (ix > g.nx + NG) return;
  const double mode_sign = (im & 1) ? -1.0 : 1.0;
  for (int j = 2; j <= JNG; ++j) {
    const size_t lo = g.at(ix, 1 - j, im);
    const size_t a = g.at(ix, j - 1, im), b = g.at(ix, j, im);
    jxm[a] = jxm[a] + mode_sign * jxm[lo];
    jrm[b] = jrm[b] - mode_sign * jrm[lo];
    jtm[a] = jtm[a] - mode_sign * jtm[lo];
    jxm[lo] = C(0.0, 0.0);
    jrm[lo] = C(0.0, 0.0);
    jtm[lo] = C(0.0, 0.0);
  }
  {
    const size_t a = g.at(ix, 1, im);
    jrm[a] = jrm[a] - mode_sign * jrm[g.at(ix, 0, im)];
  }
  const size_t a0 = g.at(ix, 0, im), a1 = g.at(ix, 1, im), a2 = g.at(ix, 2, im);
  if (im > 0) jxm[a0] = C(0.0, 0.0);
  else jxm[a0] = (4.0 * jxm[a1] - jxm[a2]) / 3.0;
  if (im == 1) {
    const cplx jt0 = (C(0.0, -1.0) * (9.0 * jrm[a1] - jrm[a2])) / 8.0;
    jtm[a0] = jt0;
    jrm[a0] = C(0.0, 2.0) * jt0 - jrm[a1];
  } else {
    jtm[a0] = C(0.0, 0.0);
    jrm[a0] = -jrm[a1];
  }
}

// particle_reflection_bcs_complex, boundary.F90:918-1015 (the COMPLEX variant's index pairing).
// x walls: one thread per (row, mode, comp); comp 0 = jx (flip_dir 1), 1 = jr, 2 = jt.
__global__ void __launch_bounds__(128) k_jreflect_x(Geom g, cplx* jx, cplx* jr, cplx* jt, int side) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x + 1 - NG;
  if (j > g.ny + NG) return;
  const int im = blockIdx.y, k = blockIdx.z;
  cplx* a = (k == 0) ? jx : (k == 1) ? jr : jt;
  const bool flip = (k == 0);
  if (side == 0) {
    for (int i = 1; i <= NG - 1; ++i) {
      if (flip) {
        const size_t o = g.at(i, j, im), s = g.at(1 - i, j, im);
        a[o] = a[o] - a[s];
        a[s] = C(0.0, 0.0);
      } else {
        const size_t o = g.at(i, j, im), s = g.at(-i, j, im);
        a[o] = a[o] + a[s];
        a[s] = C(0.0, 0.0);
      }
    }
  } else {
    const int nn = g.nx;
    for (int i = 1; i <= NG; ++i) {
      if (flip) {
        const size_t o = g.at(nn + 1 - i, j, im), s = g.at(nn + i, j, im);
        a[o] = a[o] - a[s];
        a[s] = C(0.0, 0.0);
      } else {
        const size_t o = g.at(nn - i, j, im), s = g.at(nn + i, j, im);
        a[o] = a[o] + a[s];
        a[s] = C(0.0, 0.0);
      }
    }
  }
}

// r_max wall with the face-radius ratios of boundary.F90:988-1005
__global__ void __launch_bounds__(128) k_jreflect_y(Geom g, cplx* jx, cplx* jr, cplx* jt, double dy,
                                                    double y_grid_min_local) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x + 1 - NG;
  if (ix > g.nx + NG) return;
  const int im = blockIdx.y, k = blockIdx.z;
  const int nn = g.ny;
  const double r_max = y_grid_min_local + ((double)g.ny - 0.5) * dy;
  if (k == 1) {          // jr: flip_dir == 2
    for (int i = 1; i <= NG; ++i) {
      const double num = r_max + ((double)i - 0.5) * dy, den = r_max - ((double)i - 0.5) * dy;
      const size_t o = g.at(ix, nn + 1 - i, im), s = g.at(ix, nn + i, im);
      jr[o] = jr[o] - (jr[s] * num) / den;
      jr[s] = C(0.0, 0.0);
    }
  } else if (k == 0) {   // jx: flip_dir == 1
    for (int i = 1; i <= NG; ++i) {
      const double num = r_max + (double)i * dy, den = r_max - (double)i * dy;
      const size_t o = g.at(ix, nn - i, im), s = g.at(ix, nn + i, im);
      jx[o] = jx[o] + (jx[s] * num) / den;
      jx[s] = C(0.0, 0.0);
    }
  } else {
    for (int i = 1; i <= NG; ++i) {
      const size_t o = g.at(ix, nn - i, im), s = g.at(ix, nn + i, im);
      jt[o] = jt[o] + jt[s];
      jt[s] = C(0.0, 0.0);
    }
  }
}

// ---- laser / outflow line updates: laser.f90:411-690 ----
struct FieldSet {
  cplx *exm, *erm, *etm, *bxm, *brm, *btm, *jxm, *jrm, *jtm;
  const cplx *bxo, *bro, *bto, *jxo, *jro, *jto;
};

// side 0: x_min (laser.f90:411-520), side 1: x_max (:524-633).  One thread per (ir, im).
// REFERENCE QUIRKS reproduced (quirks = 1, the default): `r_d_vals` is declared (0:ny) but used whole-array
// against (1:ny) sections, so element ir pairs with r_d_vals(ir-1); on x_max the same holds for source_t
// (laser.f90:604).  quirks = 0 (cylgpu_set_reference_quirks): element for element, r_d_vals(ir), source_t(ir).
__global__ void __launch_bounds__(128) k_outflow_x(Geom g, FieldSet F, const cplx* __restrict__ snap_er,
                                                   const cplx* __restrict__ snap_et,
                                                   const cplx* __restrict__ snap_bx,
                                                   const cplx* __restrict__ snap_br,
                                                   const cplx* __restrict__ snap_bt,
                                                   const double* __restrict__ s1, const double* __restrict__ s2,
                                                   int side, double dx, double dy, double dt,
                                                   double y_grid_min_local, int quirks) {
  const int qk = quirks ? 1 : 0;
  const int ir = blockIdx.x * blockDim.x + threadIdx.x;   // 0..ny
  const int im = blockIdx.y;
  if (ir > g.ny) return;
  const double c = C_LIGHT;
  const double dtc2 = dt * (c * c);
  const double lx = dtc2 / dx, lr = dtc2 / dy;
  const double sum = 1.0 / (lx + c), diff = lx - c, dt_eps = dt / EPSILON0;
  const size_t sn = (size_t)im * g.SY + (ir + NG - 1);   // snapshot (ir, im)
  const int nx = g.nx;
  if (side == 0) {
    F.bxm[g.at(0, ir, im)] = snap_bx[sn];
    if (ir >= 1) {
      const cplx source_t = (im == 1) ? C(s1[ir], s2[ir]) : C(0.0, 0.0);
      const double r_d_q = fabs((double)((ir - qk) - 1) * dy + y_grid_min_local);   // r_d_vals(ir-1) with the quirk
      F.btm[g.at(1, ir, im)] =
          sum * (4.0 * source_t + 2.0 * (snap_er[sn] + c * snap_bt[sn]) - 2.0 * F.erm[g.at(1, ir, im)]
                 + (((C(0.0, (double)im) * (c * c)) * dt) * F.bxm[g.at(1, ir, im)]) / r_d_q
                 + dt_eps * F.jrm[g.at(1, ir, im)] + diff * F.btm[g.at(2, ir, im)]);
    }
    if (ir >= 1 && ir <= g.ny - 1) {
      // source_r = -i*s1 + s2
      const cplx source_r = (im == 1) ? C(s2[ir], -s1[ir]) : C(0.0, 0.0);
      F.brm[g.at(1, ir, im)] =
          sum * (-4.0 * source_r - 2.0 * (snap_et[sn] + c * snap_br[sn]) + 2.0 * F.etm[g.at(1, ir, im)]
                 - lr * (F.bxm[g.at(1, ir + 1, im)] - F.bxm[g.at(1, ir, im)])
                 - dt_eps * F.jtm[g.at(1, ir, im)] + diff * F.brm[g.at(2, ir, im)]);
    }
  } else {
    F.bxm[g.at(nx, ir, im)] = snap_bx[sn];
    if (ir >= 1) {
      const cplx source_t = (im == 1) ? C(s1[ir - qk], s2[ir - qk]) : C(0.0, 0.0);
      const double r_d_q = fabs((double)((ir - qk) - 1) * dy + y_grid_min_local);
      F.btm[g.at(nx, ir, im)] =
          sum * (-4.0 * source_t - 2.0 * (snap_er[sn] + c * snap_bt[sn]) + 2.0 * F.erm[g.at(nx - 1, ir, im)]
                 - (((C(0.0, (double)im) * (c * c)) * dt) * F.bxm[g.at(nx - 1, ir, im)]) / r_d_q
                 - dt_eps * F.jrm[g.at(nx - 1, ir, im)] + diff * F.btm[g.at(nx - 1, ir, im)]);
    }
    if (ir >= 1 && ir <= g.ny - 1) {
      const cplx source_r = (im == 1) ? C(s2[ir], -s1[ir]) : C(0.0, 0.0);
      F.brm[g.at(nx, ir, im)] =
          sum * (4.0 * source_r + 2.0 * (snap_et[sn] + c * snap_br[sn]) - 2.0 * F.etm[g.at(nx - 1, ir, im)]
                 + lr * (F.bxm[g.at(nx - 1, ir + 1, im)] - F.bxm[g.at(nx - 1, ir, im)])
                 + dt_eps * F.jtm[g.at(nx - 1, ir, im)] + diff * F.brm[g.at(nx - 1, ir, im)]);
    }
  }
}

// laser.f90:637-690.  REFERENCE QUIRK reproduced: icdt_2r is declared REAL(num) but assigned
// a purely imaginary value, so it is 0 and the azimuthal coupling terms vanish.  quirks = 0
// (cylgpu_set_reference_quirks): the coefficient as the right-hand side spells it, 0.5 i c dt / r.
// Columns: Bx on ix_l..ix_h, Btheta on it_l..it_h (the reference: 0..nx without the domain-boundary column,
// and 1..nx; wider with field_ranges.cuh); the grid starts at column ix0.
__global__ void __launch_bounds__(128) k_outflow_r_max(Geom g, FieldSet F, int ix_l, int ix_h, int it_l, int it_h,
                                                       int ix0, double dx, double dy, double dt,
                                                       double y_grid_min_local, int quirks) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x + ix0;
  const int im = blockIdx.y;
  if (ix > ix_h && ix > it_h) return;
  const int ny = g.ny;
  const double c = C_LIGHT;
  const double dtc2 = dt * (c * c);
  const double inv_r = 1.0 / ((double)((float)ny - 1.5f) * dy + y_grid_min_local);
  const double dtc2_4r = 0.25 * dtc2 * inv_r;
  const double icdt_2r = 0.0;
  const double lx = dtc2 / dx, ly = dtc2 / dy;
  const double sum_x = 1.0 / (ly + c);
  const double sum_t = 1.0 / (ly + c + dtc2_4r);
  const double dt_2eps = 0.5 * dt / EPSILON0;
  if (ix >= ix_l && ix <= ix_h) {
    F.bxm[g.at(ix, ny, im)] =
        sum_x * ((-F.bxm[g.at(ix, ny - 1, im)]) * (c - ly) - F.bxo[g.at(ix, ny, im)] * (-c + ly)
                 - F.bxo[g.at(ix, ny - 1, im)] * (-c - ly) - ((c * dt) * inv_r) * F.etm[g.at(ix, ny - 1, im)]
                 + (0.5 * lx) * (F.brm[g.at(ix + 1, ny - 1, im)] - F.brm[g.at(ix, ny - 1, im)]
                                 + F.bro[g.at(ix + 1, ny - 1, im)] - F.bro[g.at(ix, ny - 1, im)])
                 - (icdt_2r * (double)im) * (F.erm[g.at(ix, ny, im)] + F.erm[g.at(ix, ny - 1, im)])
                 - dt_2eps * (F.jtm[g.at(ix, ny - 1, im)] + F.jto[g.at(ix, ny - 1, im)]));
    if (!quirks) {
      const cplx ic = C(0.0, ((0.5 * c) * dt) * inv_r) * (double)im;
      F.bxm[g.at(ix, ny, im)] = F.bxm[g.at(ix, ny, im)]
                                - sum_x * (ic * (F.erm[g.at(ix, ny, im)] + F.erm[g.at(ix, ny - 1, im)]));
    }
  }
  if (ix >= it_l && ix <= it_h) {
    F.btm[g.at(ix, ny, im)] =
        sum_t * ((-F.btm[g.at(ix, ny - 1, im)]) * (c - ly + dtc2_4r) - F.bto[g.at(ix, ny, im)] * (-c + ly + dtc2_4r)
                 - F.bto[g.at(ix, ny - 1, im)] * (-c - ly + dtc2_4r)
                 - ((0.5 * lx) / c) * (F.erm[g.at(ix, ny, im)] + F.erm[g.at(ix, ny - 1, im)]
                                       - F.erm[g.at(ix - 1, ny, im)] - F.erm[g.at(ix - 1, ny - 1, im)])
                 - ((icdt_2r * (double)im) * c) * (F.brm[g.at(ix, ny - 1, im)] + F.bro[g.at(ix, ny - 1, im)])
                 + dt_2eps * (F.jxm[g.at(ix, ny - 1, im)] + F.jxo[g.at(ix, ny - 1, im)]));
    if (!quirks) {
      const cplx ic = (C(0.0, ((0.5 * c) * dt) * inv_r) * (double)im) * c;
      F.btm[g.at(ix, ny, im)] = F.btm[g.at(ix, ny, im)]
                                - sum_t * (ic * (F.brm[g.at(ix, ny - 1, im)] + F.bro[g.at(ix, ny - 1, im)]));
    }
  }
}

// boundary.F90:1528-1531 zero_b on r_max
__global__ void __launch_bounds__(128) k_zero_b_rmax(Geom g, cplx* bxm, cplx* brm, cplx* btm) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x + 1 - NG;
  const int im = blockIdx.y;
  if (ix > g.nx + NG) return;
  const size_t o = g.at(ix, g.ny, im);
  bxm[o] = C(0.0, 0.0);
  brm[o] = C(0.0, 0.0);
  btm[o] = C(0.0, 0.0);
}

// ---- calc_number_density_modes, calc_df.F90:588-661 ----
// particle_to_grid.inc + triangle/gxfac.inc (with the r < dy fold onto the axis cell), number density =
// weight / macro-particle volume (partlist.F90:999-1013), mode factor 1 or 2 e^{i m theta}.  A
// diagnostic (dump steps only): per-particle REDs.
__global__ void __launch_bounds__(256) k_number_density(Geom g, const double* __restrict__ x, const double* __restrict__ y,
                                                        const double* __restrict__ z, const double* __restrict__ w,
                                                        int64_t n, double* __restrict__ out, double x_grid_min_local,
                                                        double y_grid_min_local, double dx, double dy, double wfac,
                                                        int nmodes) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double Y = y[i], Z = z[i];
  const double part_r = sqrt(Y * Y + Z * Z);
#if CYL_SHAPE != 0
  int cell_x, cell_y;
  double gx[NWT], gy[NWT];
  shape_particle_to_grid(x[i] - x_grid_min_local, part_r - y_grid_min_local, part_r, dx, dy, &cell_x, &cell_y, gx, gy);
  constexpr int G0 = WO;
#else
  constexpr int G0 = 1;
  const double cell_x_r = (x[i] - x_grid_min_local) / dx;
  const double cell_y_r = (part_r - y_grid_min_local) / dy;
  int cell_x = (int)floor(cell_x_r + 0.5);
  int cell_y = (int)floor(cell_y_r + 0.5);
  const double cell_frac_x = (double)cell_x - cell_x_r;
  const double cell_frac_y = (double)cell_y - cell_y_r;
  cell_x += 1;
  cell_y += 1;
  const double cx2 = cell_frac_x * cell_frac_x;
  const double gx[3] = {0.5 * (0.25 + cx2 + cell_frac_x), 0.75 - cx2, 0.5 * (0.25 + cx2 - cell_frac_x)};
  const double cy2 = cell_frac_y * cell_frac_y;
  double gy[3] = {0.5 * (0.25 + cy2 + cell_frac_y), 0.75 - cy2, 0.5 * (0.25 + cy2 - cell_frac_y)};
  if (part_r < dy) {
    gy[1] = gy[1] + gy[0];
    gy[0] = 0.0;
  }
#endif
  const double part_num_dens = (wfac * w[i]) / (2.0 * PI * dx * dy * part_r);   // wfac = 1, or the charge (calc_df.F90:479)
  const cplx exp_itheta = C(Y, Z) / part_r;
  cplx exp_imtheta = C(1.0, 0.0);
  for (int im = 0; im < nmodes; ++im) {
    cplx mode_fac = C(1.0, 0.0);
    if (im > 0) {
      exp_imtheta = exp_imtheta * exp_itheta;
      mode_fac = 2.0 * exp_imtheta;
    }
#pragma unroll
    for (int iy = SF_MIN; iy <= SF_MAX; ++iy)
#pragma unroll
      for (int ix = SF_MIN; ix <= SF_MAX; ++ix) {
        const double v = (gx[ix + G0] * gy[iy + G0]) * part_num_dens;
        if (v == 0.0) continue;
        const size_t o = 2 * g.at(cell_x + ix, cell_y + iy, im);
        atomicAdd(out + o, v * mode_fac.x);
        if (im > 0) atomicAdd(out + o + 1, v * mode_fac.y);
      }
  }
}

// ---- smooth_mode_array, current_smooth.F90:145-196: strided compensated binomial filter ----
struct Tri3 { cplx* f[3]; };
// dst(1:nx,1:ny) = alpha*src + (src(ix-s) + src(ix+s) + src(iy-s) + src(iy+s))*beta, three arrays
__global__ void __launch_bounds__(128) k_smooth(Geom g, Tri3 src, Tri3 dst, double alpha, double beta, int stride) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (ix > g.nx) return;
  const int iy = blockIdx.y + 1;
  const int im = blockIdx.z / 3, k = blockIdx.z % 3;
  const cplx* w = src.f[k];
  const size_t o = g.at(ix, iy, im);
  const size_t sy = (size_t)stride * g.SX;
  dst.f[k][o] = alpha * w[o] + (w[o - stride] + w[o + stride] + w[o - sy] + w[o + sy]) * beta;
}
// dst(1:nx,1:ny) = src(1:nx,1:ny)
__global__ void __launch_bounds__(128) k_copy_interior(Geom g, Tri3 src, Tri3 dst) {
  const int ix = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (ix > g.nx) return;
  const int iy = blockIdx.y + 1;
  const int im = blockIdx.z / 3, k = blockIdx.z % 3;
  const size_t o = g.at(ix, iy, im);
  dst.f[k][o] = src.f[k][o];
}

// ---- setup_field_boundaries, setup.F90:393-423 (no cpml: nx0 = 1, nx1 = nx) ----
struct Snaps { cplx* s[CYLGPU_NSNAPS]; };
__global__ void __launch_bounds__(128) k_snapshot(Geom g, FieldSet F, Snaps S) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const int im = blockIdx.y;
  if (row >= g.SY) return;
  const int j = row + 1 - NG;
  const size_t sn = (size_t)im * g.SY + row;
  const int nx0 = 1, nx1 = g.nx;
  S.s[0][sn] = 0.5 * (F.exm[g.at(nx0, j, im)] + F.exm[g.at(nx0 - 1, j, im)]);
  S.s[1][sn] = F.erm[g.at(nx0 - 1, j, im)];
  S.s[2][sn] = F.etm[g.at(nx0 - 1, j, im)];
  S.s[3][sn] = F.bxm[g.at(nx0 - 1, j, im)];
  S.s[4][sn] = 0.5 * (F.brm[g.at(nx0, j, im)] + F.brm[g.at(nx0 - 1, j, im)]);
  S.s[5][sn] = 0.5 * (F.btm[g.at(nx0, j, im)] + F.btm[g.at(nx0 - 1, j, im)]);
  S.s[6][sn] = 0.5 * (F.exm[g.at(nx1, j, im)] + F.exm[g.at(nx1 + 1, j, im)]);
  S.s[7][sn] = F.erm[g.at(nx1, j, im)];
  S.s[8][sn] = F.etm[g.at(nx1, j, im)];
  S.s[9][sn] = F.bxm[g.at(nx1, j, im)];
  S.s[10][sn] = 0.5 * (F.brm[g.at(nx1, j, im)] + F.brm[g.at(nx1 + 1, j, im)]);
  S.s[11][sn] = 0.5 * (F.btm[g.at(nx1, j, im)] + F.btm[g.at(nx1 + 1, j, im)]);
}

// ---- shift_fields of the moving window, window.F90:98-153 ----
// Out-of-place shift by one cell into a spare array, then the pointers are swapped: a
// pure streaming copy (1 read + 1 write per element) with no in-place hazard.
__global__ void __launch_bounds__(256) k_shift_x(Geom g, const cplx* __restrict__ src, cplx* __restrict__ dst) {
  const size_t n = g.plane * g.M;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const int col = (int)(t % g.SX);
    dst[t] = (col < g.SX - 1) ? src[t + 1] : src[t];   // last ghost column keeps its value
  }
}

__global__ void __launch_bounds__(128) k_window_fill_xmax(Geom g, FieldSet F, Snaps S) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const int im = blockIdx.y;
  if (row >= g.SY) return;
  const int j = row + 1 - NG;
  const int nx = g.nx;
  const size_t sn = (size_t)im * g.SY + row;
  // window.F90:114-131, statement order kept
  F.exm[g.at(nx + 1, j, im)] = S.s[6][sn];
  F.erm[g.at(nx, j, im)] = S.s[7][sn];
  F.etm[g.at(nx, j, im)] = S.s[8][sn];
  F.exm[g.at(nx, j, im)] = 0.5 * (F.exm[g.at(nx - 1, j, im)] + F.exm[g.at(nx + 1, j, im)]);
  F.erm[g.at(nx - 1, j, im)] = 0.5 * (F.erm[g.at(nx - 2, j, im)] + F.erm[g.at(nx, j, im)]);
  F.etm[g.at(nx - 1, j, im)] = 0.5 * (F.etm[g.at(nx - 2, j, im)] + F.etm[g.at(nx, j, im)]);
  F.bxm[g.at(nx, j, im)] = S.s[9][sn];
  F.brm[g.at(nx + 1, j, im)] = S.s[10][sn];
  F.btm[g.at(nx + 1, j, im)] = S.s[11][sn];
  F.bxm[g.at(nx - 1, j, im)] = 0.5 * (F.bxm[g.at(nx - 2, j, im)] + F.bxm[g.at(nx, j, im)]);
  F.brm[g.at(nx, j, im)] = 0.5 * (F.brm[g.at(nx - 1, j, im)] + F.brm[g.at(nx + 1, j, im)]);
  F.btm[g.at(nx, j, im)] = 0.5 * (F.btm[g.at(nx - 1, j, im)] + F.btm[g.at(nx + 1, j, im)]);
}
