// moments.cuh -- the real-valued particle moments of io/calc_df.F90 computed from the
// device-resident SoA lists (included at the end of bcs.cu, inside namespace cylgpu; it reuses
// that file's reflection / zero-gradient kernels and the packed x exchange).
//
//   calc_mass_density :59-136        calc_ekbar :140-245          calc_ekflux :249-391
//   calc_number_density :523-584     calc_ppc :665-712            calc_average_weight :716-778
//   calc_temperature :782-1033       calc_per_species_current :1037-1139
//   calc_average_momentum :1143-1221
//
// These run at dump steps only, so the deposit is one thread per particle with global REDs.  The
// work arrays are single-plane complex arrays carrying TWO real fields each (re = the data array,
// im = the weight / count array of the reference): reflection, the additive ghost exchange,
// the halo copy and the zero-gradient fill are linear and component-wise, so one pass serves
// both, and the final division commutes with the ghost copies.  Product code: no oracle here.
#pragma once

#include "moments_kernels.cuh"

// calc_boundary (calc_df.F90:24-31) on up to two single-plane work arrays: particle_reflection_bcs
// (real variant) then particle_periodic_bcs (additive ghost exchange with the x neighbours)
static int moment_summation_bcs(cylgpu_ctx* c, const Geom& g1, const int bca[4], cplx* a0, cplx* a1) {
  const dim3 gx_((g1.SY + 127) / 128, 1), gy_((g1.SX + 127) / 128, 1);
  cplx* arr[2] = {a0, a1};
  for (int k = 0; k < 2; ++k) {
    cplx* a = arr[k];
    if (!a) continue;
    if (c->cfg.x_min_boundary && bca[CYLGPU_BD_X_MIN] == CYLGPU_BC_REFLECT) {
      k_density_reflect<<<gx_, 128, 0, c->stream>>>(g1, a, CYLGPU_BD_X_MIN);
      c->stats.kernel_launches += 1;
    }
    if (c->cfg.x_max_boundary && bca[CYLGPU_BD_X_MAX] == CYLGPU_BC_REFLECT) {
      k_density_reflect<<<gx_, 128, 0, c->stream>>>(g1, a, CYLGPU_BD_X_MAX);
      c->stats.kernel_launches += 1;
    }
    if (bca[CYLGPU_BD_Y_MAX] == CYLGPU_BC_REFLECT) {
      k_density_reflect<<<gy_, 128, 0, c->stream>>>(g1, a, CYLGPU_BD_Y_MAX);
      c->stats.kernel_launches += 1;
    }
  }
  const bool to_l = c->left >= 0 && !(c->cfg.x_min_boundary && bca[CYLGPU_BD_X_MIN] != CYLGPU_BC_PERIODIC);
  const bool to_r = c->right >= 0 && !(c->cfg.x_max_boundary && bca[CYLGPU_BD_X_MAX] != CYLGPU_BC_PERIODIC);
  Halo3 h;
  h.f[0] = a0; h.f[1] = a1; h.f[2] = nullptr;
  h.skip[0] = h.skip[1] = h.skip[2] = 0;
  TRY(exchange3(c, h, 1, to_l, to_r, to_l, to_r, &g1));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// field_zero_gradient with c_stagger_centre on boundaries 1..4 (boundary.F90:597-650)
static int moment_zero_gradient(cylgpu_ctx* c, const Geom& g1, cplx* a) {
  const dim3 gx_((g1.SY + 127) / 128, 1), gy_((g1.SX + 127) / 128, 1);
  if (c->bc_field[CYLGPU_BD_X_MIN] != CYLGPU_BC_PERIODIC && c->cfg.x_min_boundary) {
    k_density_zero_gradient<<<gx_, 128, 0, c->stream>>>(g1, a, CYLGPU_BD_X_MIN);
    c->stats.kernel_launches += 1;
  }
  if (c->bc_field[CYLGPU_BD_X_MAX] != CYLGPU_BC_PERIODIC && c->cfg.x_max_boundary) {
    k_density_zero_gradient<<<gx_, 128, 0, c->stream>>>(g1, a, CYLGPU_BD_X_MAX);
    c->stats.kernel_launches += 1;
  }
  if (c->bc_field[CYLGPU_BD_Y_MIN] != CYLGPU_BC_PERIODIC) {
    k_density_zero_gradient<<<gy_, 128, 0, c->stream>>>(g1, a, CYLGPU_BD_Y_MIN);
    c->stats.kernel_launches += 1;
  }
  if (c->bc_field[CYLGPU_BD_Y_MAX] != CYLGPU_BC_PERIODIC) {
    k_density_zero_gradient<<<gy_, 128, 0, c->stream>>>(g1, a, CYLGPU_BD_Y_MAX);
    c->stats.kernel_launches += 1;
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// species < 0: the reference's `current_species <= 0` (all species that carry current).
// host_out: real array (1-ng:nx+ng, 1-ng:ny+ng).
int do_particle_moment(cylgpu_ctx* c, int kind, int species, int direction, double* host_out) {
  Geom g1 = c->g;
  g1.M = 1;
  const size_t n = g1.plane;
  int bca[4];
  for (int bd = 0; bd < 4; ++bd) {
    bca[bd] = bc_allspecies(c, bd);
    if (bd != CYLGPU_BD_Y_MIN && bca[bd] == -1) {
      set_error("mixed per-species particle boundary conditions are not supported");
      return 2;
    }
  }
  const int adir = direction < 0 ? -direction : direction;
  if (adir > 3) { set_error("particle_moment: direction must be 0 or +-1..3"); return 2; }
  if ((kind == CYLGPU_MOM_SPECIES_CURRENT || kind == CYLGPU_MOM_AVERAGE_MOMENTUM) && (direction < 1)) {
    // calc_df.F90:1053-1059,1155-1161: "No direction argument supplied"
    set_error("particle_moment: this moment needs a direction argument (1, 2 or 3)");
    return 2;
  }
  if (kind == CYLGPU_MOM_TEMPERATURE && direction < 0) { set_error("particle_moment: bad direction"); return 2; }

  // work arrays: A (+ B, D for the temperature) and the real output plane, freed on every path
  struct Work {
    cplx* a[3] = {nullptr, nullptr, nullptr};
    double* out = nullptr;
    ~Work() { for (cplx* p : a) if (p) cudaFree(p); if (out) cudaFree(out); }
  } W;
  const int narr = kind == CYLGPU_MOM_TEMPERATURE ? 3 : 1;
  for (int k = 0; k < narr; ++k) {
    CUDA_TRY(cudaMalloc(&W.a[k], n * sizeof(cplx)));
    CUDA_TRY(cudaMemsetAsync(W.a[k], 0, n * sizeof(cplx), c->stream));
  }
  CUDA_TRY(cudaMalloc(&W.out, n * sizeof(double)));

  auto for_each_species = [&](auto&& launch) {
    for (int isp = 0; isp < c->cfg.n_species; ++isp) {
      const SpeciesState& S = c->species[isp];
      if (!S.set || S.n == 0) continue;
      if (species >= 0 && isp != species) continue;
      if (species < 0 && S.sp.zero_current) continue;
      MomentArgs a;
      a.x = S.d[0]; a.y = S.d[1]; a.z = S.d[2]; a.px = S.d[3]; a.py = S.d[4]; a.pz = S.d[5]; a.w = S.d[6];
      a.n = S.n;
      a.x_grid_min_local = c->x_grid_min_local;
      a.y_grid_min_local = c->cfg.y_grid_min_local;
      a.dx = c->cfg.dx; a.dy = c->cfg.dy;
      a.mass = S.sp.mass; a.charge = S.sp.charge;
      a.kind = kind; a.direction = direction;
      launch(a, (unsigned)((S.n + 255) / 256));
      c->stats.kernel_launches += 1;
    }
  };
  const unsigned pb = (unsigned)((n + 255) / 256);
  int finish_mode = 0;
  double dof = 1.0;

  if (kind == CYLGPU_MOM_PPC || kind == CYLGPU_MOM_AVERAGE_WEIGHT) {
    for_each_species([&](const MomentArgs& a, unsigned blocks) {
      k_moment_count<<<blocks, 256, 0, c->stream>>>(g1, a, (double*)W.a[0]);
    });
    finish_mode = kind == CYLGPU_MOM_AVERAGE_WEIGHT ? 1 : 0;
  } else if (kind == CYLGPU_MOM_TEMPERATURE) {
    cplx *A = W.a[0], *B = W.a[1], *D = W.a[2];
    for_each_species([&](const MomentArgs& a, unsigned blocks) {
      k_temperature_means<<<blocks, 256, 0, c->stream>>>(g1, a, (double*)A, (double*)B);
    });
    TRY(moment_summation_bcs(c, g1, bca, A, B));
    k_temperature_normalise<<<pb, 256, 0, c->stream>>>(A, B, n);
    c->stats.kernel_launches += 1;
    {   // field_bc on the means (boundary.F90:146-153,306-358): halo copy where a neighbour fills my ghosts
      const bool has_l = c->left >= 0, has_r = c->right >= 0;
      const bool fill_r = has_r && (!c->cfg.x_max_boundary || c->bc_field[CYLGPU_BD_X_MAX] == CYLGPU_BC_PERIODIC);
      const bool fill_l = has_l && (!c->cfg.x_min_boundary || c->bc_field[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC);
      Halo3 h;
      h.f[0] = A; h.f[1] = B; h.f[2] = nullptr;
      h.skip[0] = h.skip[1] = h.skip[2] = 0;
      TRY(exchange3(c, h, 0, fill_l, fill_r, fill_l, fill_r, &g1));
    }
    for_each_species([&](const MomentArgs& a, unsigned blocks) {
      k_temperature_sigma<<<blocks, 256, 0, c->stream>>>(g1, a, A, B, (double*)D);
    });
    TRY(moment_summation_bcs(c, g1, bca, D, nullptr));
    finish_mode = 2;
    dof = direction > 0 ? 1.0 : 3.0;
    W.a[0] = D; W.a[2] = A;   // the finishing kernel reads a[0]
  } else if (kind == CYLGPU_MOM_MASS_DENSITY || kind == CYLGPU_MOM_NUMBER_DENSITY ||
             kind == CYLGPU_MOM_SPECIES_CURRENT || kind == CYLGPU_MOM_EKBAR || kind == CYLGPU_MOM_EKFLUX ||
             kind == CYLGPU_MOM_AVERAGE_MOMENTUM) {
    for_each_species([&](const MomentArgs& a, unsigned blocks) {
      k_moment_deposit<<<blocks, 256, 0, c->stream>>>(g1, a, (double*)W.a[0]);
    });
    TRY(moment_summation_bcs(c, g1, bca, W.a[0], nullptr));
    TRY(moment_zero_gradient(c, g1, W.a[0]));
    finish_mode = (kind == CYLGPU_MOM_EKBAR || kind == CYLGPU_MOM_EKFLUX || kind == CYLGPU_MOM_AVERAGE_MOMENTUM) ? 1 : 0;
  } else {
    set_error("particle_moment: unknown moment %d", kind);
    return 2;
  }
  k_moment_finish<<<pb, 256, 0, c->stream>>>(W.a[0], W.out, n, finish_mode, dof);
  c->stats.kernel_launches += 1;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(host_out, W.out, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}
