// cyl_moments.cpp -- CPU oracle, part 2: the particle moments of io/calc_df.F90.
//
// TEST INFRASTRUCTURE ONLY (see the header of cyl_oracle.hpp; PARITY UNPINNED applies here too).
// Restates, routine by routine and in the reference's operation order, the real-valued
// derived-field diagnostics the reference computes from its particle lists at dump steps:
//   calc_mass_density          calc_df.F90:59-136
//   calc_ekbar                 calc_df.F90:140-245
//   calc_ekflux                calc_df.F90:249-391
//   calc_number_density        calc_df.F90:523-584
//   calc_ppc                   calc_df.F90:665-712
//   calc_average_weight        calc_df.F90:716-778
//   calc_temperature           calc_df.F90:782-1033
//   calc_per_species_current   calc_df.F90:1037-1139
//   calc_average_momentum      calc_df.F90:1143-1221
// (calc_poynt_flux :395-438 reads the legacy Cartesian ex/ey/... arrays, which the cylindrical
// solver never fills, and is not restated.)  The weights are include/particle_to_grid.inc +
// include/triangle/gxfac.inc (r < dy fold onto the axis cell), densities divide by the
// macro-particle volume 2 pi dx dy r (partlist.F90:999-1013).  calc_boundary is the real-valued
// processor_summation_bcs (boundary.F90:1305-1326): particle_reflection_bcs :833-914 then
// particle_periodic_bcs :1019-1129; field_zero_gradient :597-650 with c_stagger_centre;
// field_bc :146-153 (x halo copy).  Real arrays live in the real part of mode 0 of an Arr3 so
// that the exchange helpers of cyl_oracle.cpp serve both.
#include <cassert>
#include <cfloat>
#include <cmath>

#include "cyl_oracle.hpp"

namespace cylo {

namespace {

struct ToGrid {   // include/particle_to_grid.inc, <shape>/gxfac.inc (cyl_oracle.hpp particle_to_grid)
  int cell_x, cell_y;
  double gx[NW], gy[NW];   // offsets -3..3 at [k + WO]
  double part_r;
};

inline ToGrid particle_to_grid(const Particle& p, double x_grid_min_local, double y_grid_min_local, double dx,
                               double dy) {
  ToGrid t;
  t.part_r = std::sqrt(p.pos[1] * p.pos[1] + p.pos[2] * p.pos[2]);
  cylo::particle_to_grid(p.pos[0] - x_grid_min_local, t.part_r - y_grid_min_local, t.part_r, dx, dy, &t.cell_x,
                         &t.cell_y, t.gx, t.gy);
  return t;
}

constexpr double C_TINY = DBL_MIN;   // constants.F90:29  TINY(1.0_num)

inline double& re(Arr3& a, int ix, int iy) { return a(ix, iy, 0).re; }

}  // namespace

// calc_boundary (calc_df.F90:24-31) without the species argument: with uniform per-species
// boundary conditions the per-species call returns at once (boundary.F90:1312-1314) and the
// final one does the work.
void World::moment_summation_bcs(Arr3 Rank::*f) {
  for (Rank& r : ranks) {   // particle_reflection_bcs, boundary.F90:833-914 (flip_direction absent)
    Arr3& a = r.*f;
    const int nx = r.nx, ny = r.ny;
    for (int im = 0; im < M; ++im) {
      if (r.x_min_boundary && bc_allspecies(BD_X_MIN) == BC_REFLECT)
        for (int i = 1; i <= NG - 1; ++i)
          for (int j = 1 - NG; j <= ny + NG; ++j) {
            a(i, j, im) = a(i, j, im) + a(1 - i, j, im);
            a(1 - i, j, im) = cplx(0.0);
          }
      if (r.x_max_boundary && bc_allspecies(BD_X_MAX) == BC_REFLECT)
        for (int i = 1; i <= NG; ++i)
          for (int j = 1 - NG; j <= ny + NG; ++j) {
            a(nx + 1 - i, j, im) = a(nx + 1 - i, j, im) + a(nx + i, j, im);
            a(nx + i, j, im) = cplx(0.0);
          }
      if (bc_allspecies(BD_Y_MAX) == BC_REFLECT)
        for (int i = 1; i <= NG; ++i)
          for (int ix = 1 - NG; ix <= nx + NG; ++ix) {
            a(ix, ny + 1 - i, im) = a(ix, ny + 1 - i, im) + a(ix, ny + i, im);
            a(ix, ny + i, im) = cplx(0.0);
          }
    }
  }
  periodic_sum_x(f);   // particle_periodic_bcs, x part (no r neighbours with x-slabs)
}

// field_zero_gradient / field_mode_zero_gradient with c_stagger_centre on boundaries 1..4
// (boundary.F90:597-650,654-707)
void World::centre_zero_gradient(Arr3 Rank::*f) {
  for (Rank& r : ranks) {
    Arr3& a = r.*f;
    const int nx = r.nx, ny = r.ny;
    for (int im = 0; im < M; ++im) {
      if (bc_field[BD_X_MIN] != BC_PERIODIC && r.x_min_boundary)
        for (int i = 1; i <= NG; ++i)
          for (int j = 1 - NG; j <= ny + NG; ++j) a(i - NG, j, im) = a(NG + 1 - i, j, im);
      if (bc_field[BD_X_MAX] != BC_PERIODIC && r.x_max_boundary)
        for (int i = 1; i <= NG; ++i)
          for (int j = 1 - NG; j <= ny + NG; ++j) a(nx + i, j, im) = a(nx + 1 - i, j, im);
      if (bc_field[BD_Y_MIN] != BC_PERIODIC)
        for (int i = 1; i <= NG; ++i)
          for (int ix = 1 - NG; ix <= nx + NG; ++ix) a(ix, i - NG, im) = a(ix, NG + 1 - i, im);
      if (bc_field[BD_Y_MAX] != BC_PERIODIC)
        for (int i = 1; i <= NG; ++i)
          for (int ix = 1 - NG; ix <= nx + NG; ++ix) a(ix, ny + i, im) = a(ix, ny + 1 - i, im);
    }
  }
}

// direction: c_dir_x/y/z = 1/2/3 (constants.F90:231-233), negative for the backward ekflux,
// 0 = argument absent (calc_temperature: all three degrees of freedom).  Result: Rank::m0, real
// part of mode 0.  species < 0 is the reference's `current_species <= 0` (sum over the species
// that carry current).
void World::calc_moment(int kind, int current_species, int direction) {
  for (int i = 0; i < 4; ++i) assert(bc_allspecies(i) != BC_MIXED);
  const bool spec_sum = current_species < 0;
  for (Rank& r : ranks) {
    r.m0.alloc(r.nx, r.ny, M);
    r.m1.alloc(r.nx, r.ny, M);
    r.m2.alloc(r.nx, r.ny, M);
    r.m3.alloc(r.nx, r.ny, M);
    r.m4.alloc(r.nx, r.ny, M);
  }
  auto skip = [&](size_t isp) {
    if (!spec_sum && (int)isp != current_species) return true;
    if (spec_sum && species[isp].zero_current) return true;
    return false;
  };
  const double c = C_LIGHT;

  if (kind == MOM_PPC || kind == MOM_AVERAGE_WEIGHT) {   // calc_df.F90:665-712, :716-778
    for (Rank& r : ranks) {
      for (size_t isp = 0; isp < species.size(); ++isp) {
        if (skip(isp)) continue;
        for (const Particle& p : r.parts[isp]) {
          const double part_r = std::sqrt(p.pos[1] * p.pos[1] + p.pos[2] * p.pos[2]);
          // (top-hat: without the half cell, calc_df.F90:696-702, 757-763)
          const double cell_x_r = (p.pos[0] - r.x_grid_min_local) / dx + (0.5 - SHAPE_CELL_SHIFT);
          const double cell_y_r = (part_r - y_grid_min_local) / dy + (0.5 - SHAPE_CELL_SHIFT);
          const int cell_x = (int)std::floor(cell_x_r) + 1;
          const int cell_y = (int)std::floor(cell_y_r) + 1;
          if (kind == MOM_PPC) {
            re(r.m0, cell_x, cell_y) = re(r.m0, cell_x, cell_y) + 1.0;
          } else {
            re(r.m0, cell_x, cell_y) = re(r.m0, cell_x, cell_y) + p.w;
            re(r.m1, cell_x, cell_y) = re(r.m1, cell_x, cell_y) + 1.0;
          }
        }
      }
      if (kind == MOM_AVERAGE_WEIGHT)
        for (size_t n = 0; n < r.m0.d.size(); ++n) r.m0.d[n].re = r.m0.d[n].re / std::max(r.m1.d[n].re, C_TINY);
    }
    return;
  }

  if (kind == MOM_TEMPERATURE) {   // calc_df.F90:782-1033
    const int dir = direction > 0 ? direction : -1;
    const double dof = direction > 0 ? 1.0 : 3.0;
    // m1..m3 = meanx, meany, meanz; m4 = part_count; m0 = sigma
    for (Rank& r : ranks)
      for (size_t isp = 0; isp < species.size(); ++isp) {
        if (skip(isp)) continue;
        const double sqrt_part_m = std::sqrt(species[isp].mass);
        for (const Particle& p : r.parts[isp]) {
          const double part_w = p.w;
          const double part_pmx = p.p[0] / sqrt_part_m;
          const double part_pmy = p.p[1] / sqrt_part_m;
          const double part_pmz = p.p[2] / sqrt_part_m;
          const ToGrid t = particle_to_grid(p, r.x_grid_min_local, y_grid_min_local, dx, dy);
          for (int iy = SF_MIN; iy <= SF_MAX; ++iy)
            for (int ix = SF_MIN; ix <= SF_MAX; ++ix) {
              const double gf = t.gx[ix + WO] * t.gy[iy + WO] * part_w;
              const int cx = t.cell_x + ix, cy = t.cell_y + iy;
              if (dir == 1 || dir == -1) re(r.m1, cx, cy) = re(r.m1, cx, cy) + gf * part_pmx;
              if (dir == 2 || dir == -1) re(r.m2, cx, cy) = re(r.m2, cx, cy) + gf * part_pmy;
              if (dir == 3 || dir == -1) re(r.m3, cx, cy) = re(r.m3, cx, cy) + gf * part_pmz;
              re(r.m4, cx, cy) = re(r.m4, cx, cy) + gf;
            }
        }
      }
    if (dir == 1 || dir == -1) moment_summation_bcs(&Rank::m1);
    if (dir == 2 || dir == -1) moment_summation_bcs(&Rank::m2);
    if (dir == 3 || dir == -1) moment_summation_bcs(&Rank::m3);
    moment_summation_bcs(&Rank::m4);
    for (Rank& r : ranks)
      for (size_t n = 0; n < r.m4.d.size(); ++n) {
        r.m4.d[n].re = std::max(r.m4.d[n].re, 1.e-6);
        r.m1.d[n].re = r.m1.d[n].re / r.m4.d[n].re;
        r.m2.d[n].re = r.m2.d[n].re / r.m4.d[n].re;
        r.m3.d[n].re = r.m3.d[n].re / r.m4.d[n].re;
      }
    // "Restore ghost cell values for means": field_bc, boundary.F90:146-153
    if (dir == 1 || dir == -1) halo_x(&Rank::m1, 0, 0);
    if (dir == 2 || dir == -1) halo_x(&Rank::m2, 0, 0);
    if (dir == 3 || dir == -1) halo_x(&Rank::m3, 0, 0);
    for (Rank& r : ranks) r.m4.zero();
    for (Rank& r : ranks)
      for (size_t isp = 0; isp < species.size(); ++isp) {
        if (skip(isp)) continue;
        const double sqrt_part_m = std::sqrt(species[isp].mass);
        for (const Particle& p : r.parts[isp]) {
          const double part_pmx = p.p[0] / sqrt_part_m;
          const double part_pmy = p.p[1] / sqrt_part_m;
          const double part_pmz = p.p[2] / sqrt_part_m;
          const ToGrid t = particle_to_grid(p, r.x_grid_min_local, y_grid_min_local, dx, dy);
          for (int iy = SF_MIN; iy <= SF_MAX; ++iy)
            for (int ix = SF_MIN; ix <= SF_MAX; ++ix) {
              const double gf = t.gx[ix + WO] * t.gy[iy + WO];
              const int cx = t.cell_x + ix, cy = t.cell_y + iy;
              double wdata;
              const double ddx = part_pmx - re(r.m1, cx, cy);
              const double ddy = part_pmy - re(r.m2, cx, cy);
              const double ddz = part_pmz - re(r.m3, cx, cy);
              if (dir == 1) wdata = ddx * ddx;
              else if (dir == 2) wdata = ddy * ddy;
              else if (dir == 3) wdata = ddz * ddz;
              else wdata = ddx * ddx + ddy * ddy + ddz * ddz;
              re(r.m0, cx, cy) = re(r.m0, cx, cy) + gf * wdata;
              re(r.m4, cx, cy) = re(r.m4, cx, cy) + gf;
            }
        }
      }
    moment_summation_bcs(&Rank::m0);
    moment_summation_bcs(&Rank::m4);
    // 3/2 kT = <p^2>/(2m)
    for (Rank& r : ranks)
      for (size_t n = 0; n < r.m0.d.size(); ++n)
        r.m0.d[n].re = r.m0.d[n].re / std::max(r.m4.d[n].re, 1.e-6) / KB / dof;
    return;
  }

  // the moments that share one deposit loop: data (m0) and, for the averages, a weight array (m1)
  const bool averaged = (kind == MOM_EKBAR || kind == MOM_EKFLUX || kind == MOM_AVERAGE_MOMENTUM);
  const double xfac = c * dy, yfac = c * dx, zfac = c * dx * dy;   // calc_df.F90:275-277
  for (Rank& r : ranks)
    for (size_t isp = 0; isp < species.size(); ++isp) {
      if (skip(isp)) continue;
      const Species& sp = species[isp];
      for (const Particle& p : r.parts[isp]) {
        double wdata = 0.0, part_w = p.w;
        const ToGrid t = particle_to_grid(p, r.x_grid_min_local, y_grid_min_local, dx, dy);
        const double macro_part_volume = 2.0 * PI * dx * dy * t.part_r;
        switch (kind) {
          case MOM_MASS_DENSITY: {   // :59-136
            wdata = sp.mass * p.w;
            wdata = wdata / macro_part_volume;
          } break;
          case MOM_NUMBER_DENSITY: {   // :523-584
            wdata = p.w;
            wdata = wdata / macro_part_volume;
          } break;
          case MOM_SPECIES_CURRENT: {   // :1037-1139
            const double part_mc = c * sp.mass;
            wdata = sp.charge * p.w;
            const double root =
                1.0 / std::sqrt(part_mc * part_mc + p.p[0] * p.p[0] + p.p[1] * p.p[1] + p.p[2] * p.p[2]);
            assert(direction >= 1 && direction <= 3);
            wdata = wdata * p.p[direction - 1] * root;
            wdata = wdata * c / macro_part_volume;
          } break;
          case MOM_EKBAR:
          case MOM_EKFLUX: {   // :140-245, :249-391
            const double part_mc = c * sp.mass;
            const double fac = part_mc * part_w * c;
            const double part_ux = p.p[0] / part_mc;
            const double part_uy = p.p[1] / part_mc;
            const double part_uz = p.p[2] / part_mc;
            const double part_u2 = part_ux * part_ux + part_uy * part_uy + part_uz * part_uz;
            const double gamma_rel = std::sqrt(part_u2 + 1.0);
            const double gamma_rel_m1 = part_u2 / (gamma_rel + 1.0);
            wdata = gamma_rel_m1 * fac;
            if (kind == MOM_EKFLUX) {
              double part_flux;
              switch (direction) {
                case -1: part_flux = xfac * part_ux / gamma_rel; wdata = -wdata * std::min(part_flux, 0.0); break;
                case 1: part_flux = xfac * part_ux / gamma_rel; wdata = wdata * std::max(part_flux, 0.0); break;
                case -2: part_flux = yfac * part_uy / gamma_rel; wdata = -wdata * std::min(part_flux, 0.0); break;
                case 2: part_flux = yfac * part_uy / gamma_rel; wdata = wdata * std::max(part_flux, 0.0); break;
                case -3: part_flux = zfac * part_uz / gamma_rel; wdata = -wdata * std::min(part_flux, 0.0); break;
                case 3: part_flux = zfac * part_uz / gamma_rel; wdata = wdata * std::max(part_flux, 0.0); break;
                default: break;   // SELECT CASE without a matching case: wdata unchanged
              }
            }
          } break;
          case MOM_AVERAGE_MOMENTUM: {   // :1143-1221
            assert(direction >= 1 && direction <= 3);
            wdata = p.w * p.p[direction - 1];
          } break;
          default: assert(!"unknown moment");
        }
        for (int iy = SF_MIN; iy <= SF_MAX; ++iy)
          for (int ix = SF_MIN; ix <= SF_MAX; ++ix) {
            const int cx = t.cell_x + ix, cy = t.cell_y + iy;
            re(r.m0, cx, cy) = re(r.m0, cx, cy) + t.gx[ix + WO] * t.gy[iy + WO] * wdata;
            if (averaged) re(r.m1, cx, cy) = re(r.m1, cx, cy) + t.gx[ix + WO] * t.gy[iy + WO] * part_w;
          }
      }
    }
  moment_summation_bcs(&Rank::m0);
  if (averaged) {
    moment_summation_bcs(&Rank::m1);
    for (Rank& r : ranks)
      for (size_t n = 0; n < r.m0.d.size(); ++n) r.m0.d[n].re = r.m0.d[n].re / std::max(r.m1.d[n].re, C_TINY);
  }
  centre_zero_gradient(&Rank::m0);
}

}  // namespace cylo
