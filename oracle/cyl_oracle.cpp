// cyl_oracle.cpp -- CPU oracle (TEST INFRASTRUCTURE ONLY; see cyl_oracle.hpp header).
// Every routine cites the reference file:line it restates (paths relative to
// /root/reference/epoch_axial/src).  Build with -ffp-contract=off.
#include "cyl_oracle.hpp"

#include <cassert>
#include <cstdio>
#include <cstdlib>

namespace cylo {

// ---------------------------------------------------------------------------------------
// RNG: random_generator.f90:45-108 (KISS), :112-173 (polar Box-Muller with cached spare)
// ---------------------------------------------------------------------------------------
void Rng::init(int seed) {
  x = (uint32_t)(123456789 + seed);
  y = (uint32_t)(362436069 + seed);
  z = (uint32_t)(521288629 + seed);
  w = (uint32_t)(916191069 + seed);
  cached = false;
  cached_value = 0.0;
  for (int i = 0; i < 1000; ++i) (void)uniform();
}

double Rng::uniform() {
  x = 69069u * x + 1327217885u;
  uint32_t a = y;
  a ^= a << 13;
  a ^= a >> 17;
  a ^= a << 5;
  y = a;
  z = 18000u * (z & 65535u) + (z >> 16);
  w = 30903u * (w & 65535u) + (w >> 16);
  uint32_t k = x + y + (z << 16) + w;
  int32_t kiss = (int32_t)k;
  return ((double)kiss + 2147483648.0) / 4294967296.0;
}

double Rng::box_muller(double stdev, double mu) {
  if (cached) {
    cached = false;
    return cached_value * stdev + mu;
  }
  cached = true;
  double r1, r2, ww;
  const double tiny = 2.2250738585072014e-308;
  for (;;) {
    r1 = uniform();
    r2 = uniform();
    r1 = 2.0 * r1 - 1.0;
    r2 = 2.0 * r2 - 1.0;
    ww = r1 * r1 + r2 * r2;
    if (ww > tiny && ww < 1.0) break;
  }
  ww = std::sqrt((-2.0 * std::log(ww)) / ww);
  cached_value = r2 * ww;
  return r1 * ww * stdev + mu;
}

// ---------------------------------------------------------------------------------------
// Set-up: setup.F90:164-206 (grid), :629-646 (dt), mpi_routines.F90:312-337 (slabs),
// :385-400 (allocation)
// ---------------------------------------------------------------------------------------
World::World(const Config& c) : cfg(c) {
  M = c.n_mode;
  x_min = c.x_min;
  x_max = c.x_max;
  y_max = c.y_max;
  length_x = x_max - x_min;
  dx = length_x / (double)c.nx_global;
  x_grid_min = x_min;
  double length_y = y_max - 0.0;
  dy = length_y / (double)c.ny_global;
  double y_grid_min = 0.0;
  xb_min = x_grid_min;
  x_grid_min = x_grid_min + dx / 2.0;
  y_grid_min = y_grid_min + dy / 2.0;
  y_grid_min_local = y_grid_min;

  dt = 0.9 * std::min(dx, dy) / std::sqrt(2.0) / C_LIGHT;
  dt = c.dt_multiplier * dt;
  time = 0.0;
  step = 0;
  for (int i = 0; i < 4; ++i) bc_field[i] = c.bc_field[i];

  int P = c.nranks;
  ranks.resize(P);
  int nx0 = c.nx_global / P;
  int nxp = (nx0 * P != c.nx_global) ? (nx0 + 1) * P - c.nx_global : P;
  for (int k = 0; k < P; ++k) {
    Rank& r = ranks[k];
    int idim = k + 1;
    if (idim <= nxp) {
      r.cell_x_min = (idim - 1) * nx0 + 1;
      r.cell_x_max = idim * nx0;
    } else {
      r.cell_x_min = nxp * nx0 + (idim - nxp - 1) * (nx0 + 1) + 1;
      r.cell_x_max = nxp * nx0 + (idim - nxp) * (nx0 + 1);
    }
    r.nx = r.cell_x_max - r.cell_x_min + 1;
    r.ny = c.ny_global;
    r.M = M;
    r.x_coord = k;
    r.nprocx = P;
    r.x_min_boundary = (k == 0);
    r.x_max_boundary = (k == P - 1);
    Arr3* all[] = {&r.exm, &r.erm, &r.etm, &r.bxm, &r.brm, &r.btm, &r.jxm, &r.jrm, &r.jtm,
                   &r.bxm_old, &r.brm_old, &r.btm_old, &r.jxm_old, &r.jrm_old, &r.jtm_old};
    for (Arr3* a : all) a->alloc(r.nx, r.ny, M);
    Arr2* snaps[] = {&r.exm_x_min, &r.erm_x_min, &r.etm_x_min, &r.bxm_x_min, &r.brm_x_min, &r.btm_x_min,
                     &r.exm_x_max, &r.erm_x_max, &r.etm_x_max, &r.bxm_x_max, &r.brm_x_max, &r.btm_x_max};
    for (Arr2* a : snaps) a->alloc(r.ny, M);
    r.rng.init(7842432 + k);   // setup.F90:563-567
  }
  setup_grid_x();
  setup_boundaries();
}

void World::setup_grid_x() {   // utilities.f90:343-372 with cpml offsets = 0
  for (Rank& r : ranks) {
    r.x_grid_min_local = x_grid_min + (double)(r.cell_x_min - 1) * dx;
    r.x_grid_max_local = x_grid_min + (double)(r.cell_x_max - 1) * dx;
    r.x_min_local = r.x_grid_min_local + (0 - 0.5) * dx;
    r.x_max_local = r.x_grid_max_local - (0 - 0.5) * dx;
  }
}

int World::add_species(const Species& s) {
  species.push_back(s);
  for (Rank& r : ranks) r.parts.emplace_back();
  setup_boundaries();
  return (int)species.size() - 1;
}

void World::setup_boundaries() {   // boundary.F90:30-75, :100-140
  for (int i = 0; i < 4; ++i) {
    add_laser[i] = false;
    if (bc_field[i] == BC_OTHER) bc_field[i] = BC_CLAMP;
    if (bc_field[i] == BC_SIMPLE_LASER) add_laser[i] = true;
    if (bc_field[i] == BC_REFLECT) bc_field[i] = BC_CLAMP;
    if (bc_field[i] == BC_OPEN) bc_field[i] = BC_SIMPLE_OUTFLOW;
  }
  for (Species& s : species) {
    for (int i = 0; i < 4; ++i) {
      if (i == BD_Y_MIN) continue;
      int& b = s.bc_particle[i];
      if (b == BC_OTHER || b == BC_CONDUCT) b = BC_REFLECT;
      if (b == BC_SIMPLE_LASER || b == BC_SIMPLE_OUTFLOW || b == BC_CPML_LASER || b == BC_CPML_OUTFLOW)
        b = BC_OPEN;
    }
  }
}

int World::bc_allspecies(int bd) const {
  if (species.empty()) return BC_OPEN;
  int b = species[0].bc_particle[bd];
  for (const Species& s : species)
    if (s.bc_particle[bd] != b) return BC_MIXED;
  return b;
}

// ---------------------------------------------------------------------------------------
// Loader (init only; stays on the host in the product).  helper.F90:552-583 positions,
// :713-782 weights (include/particle_to_grid.inc, triangle/gxfac.inc),
// particle_temperature.F90:30-83,388-398 momenta, partlist.F90:999-1011 volume.
// Uniform density / temperature / drift, integer particles per cell.
// ---------------------------------------------------------------------------------------
void World::load_uniform(int isp) {
  const Species& s = species[isp];
  int64_t ppc = (int64_t)std::floor(s.npart_per_cell);
  for (Rank& r : ranks) {
    std::vector<Particle>& pl = r.parts[isp];
    pl.clear();
    pl.reserve((size_t)ppc * r.nx * r.ny);
    for (int iy = 1; iy <= r.ny; ++iy) {
      for (int ix = 1; ix <= r.nx; ++ix) {
        double xc = r.x_grid_min_local + (double)(ix - 1) * dx;     // x(ix)
        double yc = y_grid_min_local + (double)(iy - 1) * dy;       // y(iy)
        for (int64_t ip = 0; ip < ppc; ++ip) {
          Particle p;
          p.pos[0] = xc + (r.rng.uniform() - 0.5) * dx;
          double part_r = yc + (r.rng.uniform() - 0.5) * dy;
          double part_th = 2.0 * PI * r.rng.uniform();
          p.pos[1] = part_r * std::cos(part_th);
          p.pos[2] = part_r * std::sin(part_th);
          p.p[0] = p.p[1] = p.p[2] = 0.0;
          p.w = 0.0;
          pl.push_back(p);
        }
      }
    }
    // weights, first pass: density at the particle through the normalised triangle shape
    std::vector<int> npart_in_cell((size_t)(r.nx + 2 * NG) * (r.ny + 2 * NG), 0);
    auto cidx = [&](int cx, int cy) { return (size_t)(cy + NG - 1) * (r.nx + 2 * NG) + (cx + NG - 1); };
    for (Particle& p : pl) {
      double part_r = std::sqrt(p.pos[1] * p.pos[1] + p.pos[2] * p.pos[2]);
      int cell_x, cell_y;
      double gx[NW], gy[NW];
      particle_to_grid(p.pos[0] - r.x_grid_min_local, part_r - y_grid_min_local, part_r, dx, dy, &cell_x, &cell_y, gx, gy);
      double wdata = 0.0;
      for (int isuby = SF_MIN; isuby <= SF_MAX; ++isuby)
        for (int isubx = SF_MIN; isubx <= SF_MAX; ++isubx)
          wdata = wdata + gx[isubx + WO] * gy[isuby + WO] * s.density;   // uniform density map
      p.w = wdata;
#if CYLO_SHAPE == 1
      // top-hat: (cell_x, cell_y) may not be the cell that holds the particle (helper.F90:752-757)
      if (gx[1 + WO] > gx[0 + WO]) cell_x = cell_x + 1;
      if (gy[1 + WO] > gy[0 + WO]) cell_y = cell_y + 1;
#endif
      npart_in_cell[cidx(cell_x, cell_y)] += 1;
    }
    // second pass: macro-particle volume / particles in cell
    for (Particle& p : pl) {
      double part_r = std::sqrt(p.pos[1] * p.pos[1] + p.pos[2] * p.pos[2]);
      int cell_x = (int)std::floor((p.pos[0] - r.x_grid_min_local) / dx + 1.5);
      int cell_y = (int)std::floor((part_r - y_grid_min_local) / dy + 1.5);
      double vol = 2.0 * PI * dx * dy * part_r;
      p.w = p.w * vol / (double)npart_in_cell[cidx(cell_x, cell_y)];
    }
    // momenta, direction by direction (helper.F90:139-142)
    for (int n = 0; n < 3; ++n) {
      for (Particle& p : pl) {
        // uniform temperature/drift: the 3x3 normalised interpolation is restated so the
        // rounding matches (particle_temperature.F90:54-63)
        double part_r = std::sqrt(p.pos[1] * p.pos[1] + p.pos[2] * p.pos[2]);
        int cell_x, cell_y;
        double gx[NW], gy[NW];
        particle_to_grid(p.pos[0] - r.x_grid_min_local, part_r - y_grid_min_local, part_r, dx, dy, &cell_x, &cell_y, gx, gy);
        double temp_local = 0.0, drift_local = 0.0;
        for (int iy = SF_MIN; iy <= SF_MAX; ++iy)
          for (int ix = SF_MIN; ix <= SF_MAX; ++ix) {
            temp_local = temp_local + gx[ix + WO] * gy[iy + WO] * s.temp[n];
            drift_local = drift_local + gx[ix + WO] * gy[iy + WO] * s.drift[n];
          }
        double stdev = std::sqrt(temp_local * KB * s.mass);
        p.p[n] = r.rng.box_muller(stdev, drift_local);
      }
    }
  }
}

void World::snapshot_field_boundaries() {   // setup.F90:393-423 (no cpml: nx0 = 1, nx1 = nx)
  for (Rank& r : ranks) {
    int nx0 = 1, nx1 = r.nx;
    for (int im = 0; im < M; ++im)
      for (int j = 1 - NG; j <= r.ny + NG; ++j) {
        r.exm_x_min(j, im) = 0.5 * (r.exm(nx0, j, im) + r.exm(nx0 - 1, j, im));
        r.erm_x_min(j, im) = r.erm(nx0 - 1, j, im);
        r.etm_x_min(j, im) = r.etm(nx0 - 1, j, im);
        r.exm_x_max(j, im) = 0.5 * (r.exm(nx1, j, im) + r.exm(nx1 + 1, j, im));
        r.erm_x_max(j, im) = r.erm(nx1, j, im);
        r.etm_x_max(j, im) = r.etm(nx1, j, im);
        r.bxm_x_min(j, im) = r.bxm(nx0 - 1, j, im);
        r.brm_x_min(j, im) = 0.5 * (r.brm(nx0, j, im) + r.brm(nx0 - 1, j, im));
        r.btm_x_min(j, im) = 0.5 * (r.btm(nx0, j, im) + r.btm(nx0 - 1, j, im));
        r.bxm_x_max(j, im) = r.bxm(nx1, j, im);
        r.brm_x_max(j, im) = 0.5 * (r.brm(nx1, j, im) + r.brm(nx1 + 1, j, im));
        r.btm_x_max(j, im) = 0.5 * (r.btm(nx1, j, im) + r.btm(nx1 + 1, j, im));
      }
  }
}

void World::init_half_step() {   // epoch2d.F90:143-161
  particle_bcs();
  efield_bcs();
  double dt_store = dt;
  dt = dt / 2.0;
  time = time + dt;
  bfield_final_bcs();
  dt = dt_store;
}

// ---------------------------------------------------------------------------------------
// Field solver: fields.f90:53-182 (E), :186-312 (B)
// ---------------------------------------------------------------------------------------
void World::update_e_field(Rank& r) {
  const int nx = r.nx, ny = r.ny;
  const double c = C_LIGHT;
  const int ir_min = 1;   // y_min_boundary is always true (nprocy = 1)
  for (int im = 0; im < M; ++im) {
    for (int ir = ir_min; ir <= ny; ++ir) {
      double r_d = std::abs((double)(ir - 1) * dy + y_grid_min_local);
      double r_p = r_d + 0.5 * dy;
      double fac_x = c * c / r_p;
      cplx im_fac_x = (IMAGI * (double)im) * fac_x;
      cplx im_fac_r = ((IMAGI * (double)im) * (c * c)) / r_d;
      for (int ix = 0; ix <= nx; ++ix) {
        r.exm(ix, ir, im) = r.exm(ix, ir, im) +
            (((fac_x * 0.5) * (r.btm(ix, ir + 1, im) + r.btm(ix, ir, im))
              + im_fac_x * r.brm(ix, ir, im)
              + ((c * c) * (r.btm(ix, ir + 1, im) - r.btm(ix, ir, im))) / dy
              - r.jxm(ix, ir, im) / EPSILON0) * 0.5) * dt;
        r.erm(ix, ir, im) = r.erm(ix, ir, im) +
            (((-im_fac_r) * r.bxm(ix, ir, im)
              - ((c * c) * (r.btm(ix + 1, ir, im) - r.btm(ix, ir, im))) / dx
              - r.jrm(ix, ir, im) / EPSILON0) * 0.5) * dt;
        r.etm(ix, ir, im) = r.etm(ix, ir, im) +
            ((((c * c) * (r.brm(ix + 1, ir, im) - r.brm(ix, ir, im))) / dx
              - ((c * c) * (r.bxm(ix, ir + 1, im) - r.bxm(ix, ir, im))) / dy
              - r.jtm(ix, ir, im) / EPSILON0) * 0.5) * dt;
      }
    }
  }
  // axis (fields.f90:116-180)
  const int lo = 1 - NG, hi = nx + NG;
  for (int ix = lo; ix <= hi; ++ix) {
    // m = 0
    r.exm(ix, 0, 0) = r.exm(ix, 0, 0) +
        ((((4.0 * (c * c)) / dy) * r.btm(ix, 1, 0) - r.jxm(ix, 0, 0) / EPSILON0) * 0.5) * dt;
    r.etm(ix, 0, 0) = cplx(0.0);
    r.erm(ix, 0, 0) = -r.erm(ix, 1, 0);
    for (int ir = 1 - NG; ir <= -1; ++ir) {
      r.etm(ix, ir, 0) = -r.etm(ix, -ir, 0);
      r.erm(ix, ir, 0) = -r.erm(ix, -ir + 1, 0);
      r.exm(ix, ir, 0) = r.exm(ix, -ir, 0);
    }
    if (M > 1) {
      r.exm(ix, 0, 1) = cplx(0.0);
      r.erm(ix, 0, 1) = (2.0 * IMAGI) * r.etm(ix, 0, 1) - r.erm(ix, 1, 1);
      r.etm(ix, 0, 1) = ((-IMAGI) / 8.0) * (9.0 * r.erm(ix, 1, 1) - r.erm(ix, 2, 1));
      for (int ir = 1 - NG; ir <= -1; ++ir) {
        r.etm(ix, ir, 1) = r.etm(ix, -ir, 1);
        r.erm(ix, ir, 1) = r.erm(ix, -ir + 1, 1);
        r.exm(ix, ir, 1) = -r.exm(ix, -ir, 1);
      }
    }
    double mode_sign = 1.0;
    for (int im = 2; im < M; ++im) {
      r.exm(ix, 0, im) = cplx(0.0);
      r.etm(ix, 0, im) = cplx(0.0);
      r.erm(ix, 0, im) = -r.erm(ix, 1, im);
      r.erm(ix, 1, im) = r.erm(ix, 2, im) / 9.0;
      for (int ir = 1 - NG; ir <= -1; ++ir) {
        r.etm(ix, ir, im) = (-mode_sign) * r.etm(ix, -ir, im);
        r.erm(ix, ir, im) = (-mode_sign) * r.erm(ix, -ir + 1, im);
        r.exm(ix, ir, im) = mode_sign * r.exm(ix, -ir, im);
      }
      mode_sign = -mode_sign;
    }
  }
}

void World::update_b_field(Rank& r) {
  const int nx = r.nx, ny = r.ny;
  const int ir_min = 1, ir_max = ny - 1;   // y_min_boundary, y_max_boundary both true
  for (int im = 0; im < M; ++im) {
    for (int ir = ir_min; ir <= ir_max; ++ir) {
      double r_d = std::abs((double)(ir - 1) * dy + y_grid_min_local);
      double r_p = r_d + 0.5 * dy;
      cplx im_fac_x = (IMAGI * (double)im) / r_d;
      cplx im_fac_r = (IMAGI * (double)im) / r_p;
      for (int ix = 0; ix <= nx; ++ix) {
        r.bxm(ix, ir, im) = r.bxm(ix, ir, im) -
            ((im_fac_x * r.erm(ix, ir, im)
              + (0.5 * (r.etm(ix, ir, im) + r.etm(ix, ir - 1, im))) / r_d
              + (r.etm(ix, ir, im) - r.etm(ix, ir - 1, im)) / dy) * 0.5) * dt;
        r.brm(ix, ir, im) = r.brm(ix, ir, im) +
            ((im_fac_r * r.exm(ix, ir, im)
              + (r.etm(ix, ir, im) - r.etm(ix - 1, ir, im)) / dx) * 0.5) * dt;
        r.btm(ix, ir, im) = r.btm(ix, ir, im) +
            (((-(r.erm(ix, ir, im) - r.erm(ix - 1, ir, im))) / dx
              + (r.exm(ix, ir, im) - r.exm(ix, ir - 1, im)) / dy) * 0.5) * dt;
      }
    }
  }
  // axis (fields.f90:249-310).  The m=1 Brm(0) update reads etm(ix-1) so it is done for
  // all ix before the per-column mirror fills (it only touches row 0 of brm/exm/etm).
  const int lo = 1 - NG, hi = nx + NG;
  if (M > 1) {
    for (int ix = lo + 1; ix <= hi; ++ix) {
      r.brm(ix, 0, 1) = r.brm(ix, 0, 1) +
          (((IMAGI / dy) * r.exm(ix, 0, 1) + (r.etm(ix, 0, 1) - r.etm(ix - 1, 0, 1)) / dx) * 0.5) * dt;
    }
  }
  for (int ix = lo; ix <= hi; ++ix) {
    r.brm(ix, 0, 0) = cplx(0.0);
    r.bxm(ix, 0, 0) = r.bxm(ix, 1, 0);
    r.btm(ix, 0, 0) = -r.btm(ix, 1, 0);
    for (int ir = 1 - NG; ir <= -1; ++ir) {
      r.btm(ix, ir, 0) = -r.btm(ix, -ir + 1, 0);
      r.brm(ix, ir, 0) = -r.brm(ix, -ir, 0);
      r.bxm(ix, ir, 0) = r.bxm(ix, -ir + 1, 0);
    }
    if (M > 1) {
      r.bxm(ix, 0, 1) = -r.bxm(ix, 1, 1);
      r.btm(ix, 0, 1) = (-2.0 * IMAGI) * r.brm(ix, 0, 1) - r.btm(ix, 1, 1);
      for (int ir = 1 - NG; ir <= -1; ++ir) {
        r.btm(ix, ir, 1) = r.btm(ix, -ir + 1, 1);
        r.brm(ix, ir, 1) = r.brm(ix, -ir, 1);
        r.bxm(ix, ir, 1) = -r.bxm(ix, -ir + 1, 1);
      }
    }
    double mode_sign = 1.0;
    for (int im = 2; im < M; ++im) {
      r.bxm(ix, 0, im) = -r.bxm(ix, 1, im);
      r.brm(ix, 0, im) = cplx(0.0);
      r.btm(ix, 0, im) = -r.btm(ix, 1, im);
      for (int ir = 1 - NG; ir <= -1; ++ir) {
        r.btm(ix, ir, im) = (-mode_sign) * r.btm(ix, -ir + 1, im);
        r.brm(ix, ir, im) = (-mode_sign) * r.brm(ix, -ir, im);
        r.bxm(ix, ir, im) = mode_sign * r.bxm(ix, -ir + 1, im);
      }
      mode_sign = -mode_sign;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Halo and edge boundary conditions
// ---------------------------------------------------------------------------------------
// boundary.F90:158-169,500-553: x-direction exchange of ng interior columns into the
// neighbour's ghost columns, every mode.  With nprocy = 1 the r-direction sendrecvs go to
// MPI_PROC_NULL and the copies are guarded off (:566,:579), so nothing happens in r.
// `top_skip` = 1 reproduces the one-row shift of the r-staggered arrays (boundary.F90:
// 1362-1371,1428-1430): row ny+ng of Exm, Etm, Brm never takes part in the exchange.
void World::halo_x(Arr3 Rank::*f, int /*row_lo_off*/, int top_skip) {
  const int P = (int)ranks.size();
  for (int k = 0; k < P; ++k) {
    Rank& r = ranks[k];
    Arr3& a = r.*f;
    const int jlo = 1 - NG, jhi = r.ny + NG - top_skip;
    if (!r.x_max_boundary || bc_field[BD_X_MAX] == BC_PERIODIC) {
      const Rank& s = ranks[(k + 1) % P];
      const Arr3& b = s.*f;
      for (int im = 0; im < M; ++im)
        for (int j = jlo; j <= jhi; ++j)
          for (int i = 1; i <= NG; ++i) a(r.nx + i, j, im) = b(i, j, im);
    }
    if (!r.x_min_boundary || bc_field[BD_X_MIN] == BC_PERIODIC) {
      const Rank& s = ranks[(k - 1 + P) % P];
      const Arr3& b = s.*f;
      for (int im = 0; im < M; ++im)
        for (int j = jlo; j <= jhi; ++j)
          for (int i = 1; i <= NG; ++i) a(i - NG, j, im) = b(s.nx - NG + i, j, im);
    }
  }
}

void World::clamp_zero(Rank& r, Arr3& f, bool stag_x, bool stag_y, int bd) {   // boundary.F90:772-829
  if (bc_field[bd] == BC_PERIODIC) return;
  const int nx = r.nx, ny = r.ny;
  if (bd == BD_X_MIN && r.x_min_boundary) {
    for (int im = 0; im < M; ++im)
      for (int j = 1 - NG; j <= ny + NG; ++j) {
        if (stag_x) {
          for (int i = 1; i <= NG - 1; ++i) f(i - NG, j, im) = -f(NG - i, j, im);
          f(0, j, im) = cplx(0.0);
        } else {
          for (int i = 1; i <= NG; ++i) f(i - NG, j, im) = -f(NG + 1 - i, j, im);
        }
      }
  } else if (bd == BD_X_MAX && r.x_max_boundary) {
    const int nn = nx;
    for (int im = 0; im < M; ++im)
      for (int j = 1 - NG; j <= ny + NG; ++j) {
        if (stag_x) {
          f(nn, j, im) = cplx(0.0);
          for (int i = 1; i <= NG - 1; ++i) f(nn + i, j, im) = -f(nn - i, j, im);
        } else {
          for (int i = 1; i <= NG; ++i) f(nn + i, j, im) = -f(nn + 1 - i, j, im);
        }
      }
  } else if (bd == BD_Y_MAX) {
    const int nn = ny;
    for (int im = 0; im < M; ++im) {
      if (stag_y) {
        for (int ix = 1 - NG; ix <= nx + NG; ++ix) f(ix, nn, im) = cplx(0.0);
        for (int i = 1; i <= NG - 1; ++i)
          for (int ix = 1 - NG; ix <= nx + NG; ++ix) f(ix, nn + i, im) = -f(ix, nn - i, im);
      } else {
        for (int i = 1; i <= NG; ++i)
          for (int ix = 1 - NG; ix <= nx + NG; ++ix) f(ix, nn + i, im) = -f(ix, nn + 1 - i, im);
      }
    }
  }
}

void World::zero_gradient(Rank& r, Arr3& f, bool stag_x, bool stag_y, int bd) {   // boundary.F90:654-707
  if (bc_field[bd] == BC_PERIODIC) return;
  const int nx = r.nx, ny = r.ny;
  if (bd == BD_X_MIN && r.x_min_boundary) {
    for (int im = 0; im < M; ++im)
      for (int j = 1 - NG; j <= ny + NG; ++j) {
        if (stag_x) {
          for (int i = 1; i <= NG - 1; ++i) f(i - NG, j, im) = f(NG - i, j, im);
        } else {
          for (int i = 1; i <= NG; ++i) f(i - NG, j, im) = f(NG + 1 - i, j, im);
        }
      }
  } else if (bd == BD_X_MAX && r.x_max_boundary) {
    const int nn = nx;
    for (int im = 0; im < M; ++im)
      for (int j = 1 - NG; j <= ny + NG; ++j) {
        if (stag_x) {
          for (int i = 1; i <= NG - 1; ++i) f(nn + i, j, im) = f(nn - i, j, im);
        } else {
          for (int i = 1; i <= NG; ++i) f(nn + i, j, im) = f(nn + 1 - i, j, im);
        }
      }
  } else if (bd == BD_Y_MAX) {
    const int nn = ny;
    for (int im = 0; im < M; ++im) {
      if (stag_y) {
        for (int i = 1; i <= NG - 1; ++i)
          for (int ix = 1 - NG; ix <= nx + NG; ++ix) f(ix, nn + i, im) = f(ix, nn - i, im);
      } else {
        for (int i = 1; i <= NG; ++i)
          for (int ix = 1 - NG; ix <= nx + NG; ++ix) f(ix, nn + i, im) = f(ix, nn + 1 - i, im);
      }
    }
  }
}

// stagger flags (setup.F90:126-136, constants.F90:291-299): Exm (x:F,r:T), Erm (T,F),
// Etm (T,T), Bxm (T,F), Brm (F,T), Btm (F,F)
void World::efield_bcs() {   // boundary.F90:1355-1413
  halo_x(&Rank::exm, 0, 1);
  halo_x(&Rank::erm, 0, 0);
  halo_x(&Rank::etm, 0, 1);
  for (Rank& r : ranks) {
    for (int i : {BD_X_MIN, BD_X_MAX}) {
      if (bc_field[i] == BC_CONDUCT) {
        clamp_zero(r, r.exm, false, true, i);
        zero_gradient(r, r.erm, true, false, i);
        zero_gradient(r, r.etm, true, true, i);
      }
    }
    if (bc_field[BD_Y_MAX] == BC_CONDUCT) {
      zero_gradient(r, r.exm, false, true, BD_Y_MAX);
      clamp_zero(r, r.erm, true, false, BD_Y_MAX);
      zero_gradient(r, r.etm, true, true, BD_Y_MAX);
    }
    for (int i = 0; i < 4; ++i) {
      if (i == BD_Y_MIN) continue;
      int b = bc_field[i];
      if (b == BC_CLAMP || b == BC_SIMPLE_LASER || b == BC_SIMPLE_OUTFLOW) {
        clamp_zero(r, r.exm, false, true, i);
        clamp_zero(r, r.erm, true, false, i);
        clamp_zero(r, r.etm, true, true, i);
      }
      if (b == BC_ZERO_GRADIENT || b == BC_CPML_LASER || b == BC_CPML_OUTFLOW) {
        zero_gradient(r, r.exm, false, true, i);
        zero_gradient(r, r.erm, true, false, i);
        zero_gradient(r, r.etm, true, true, i);
      }
    }
  }
}

void World::bfield_bcs(bool mpi_only) {   // boundary.F90:1417-1476
  halo_x(&Rank::bxm, 0, 0);
  halo_x(&Rank::brm, 0, 1);
  halo_x(&Rank::btm, 0, 0);
  if (mpi_only) return;
  for (Rank& r : ranks) {
    for (int i : {BD_X_MIN, BD_X_MAX}) {
      if (bc_field[i] == BC_CONDUCT) {
        zero_gradient(r, r.bxm, true, false, i);
        clamp_zero(r, r.brm, false, true, i);
        clamp_zero(r, r.btm, false, false, i);
      }
    }
    if (bc_field[BD_Y_MAX] == BC_CONDUCT) {
      clamp_zero(r, r.bxm, true, false, BD_Y_MAX);
      zero_gradient(r, r.brm, false, true, BD_Y_MAX);
      clamp_zero(r, r.btm, false, false, BD_Y_MAX);
    }
    for (int i = 0; i < 4; ++i) {
      if (i == BD_Y_MIN) continue;
      int b = bc_field[i];
      if (b == BC_CLAMP || b == BC_SIMPLE_LASER || b == BC_SIMPLE_OUTFLOW) {
        clamp_zero(r, r.bxm, true, false, i);
        clamp_zero(r, r.brm, false, true, i);
        clamp_zero(r, r.btm, false, false, i);
      }
      if (b == BC_ZERO_GRADIENT || b == BC_CPML_LASER || b == BC_CPML_OUTFLOW) {
        zero_gradient(r, r.bxm, true, false, i);
        zero_gradient(r, r.brm, false, true, i);
        zero_gradient(r, r.btm, false, false, i);
      }
    }
  }
}

// laser.f90:276-328 (phase = omega*time), :443-461 (source1/source2 on 0:ny).
// This is the host-side (deck/parser) part that stays in Fortran in the product; the
// product receives source1/source2 through cylgpu_fields_final().
void World::laser_sources(int bd, const Rank& r, std::vector<double>& s1, std::vector<double>& s2) {
  s1.assign(r.ny + 1, 0.0);
  s2.assign(r.ny + 1, 0.0);
  if (!add_laser[bd]) return;
  for (const Laser& L : lasers) {
    if (L.boundary != bd) continue;
    if (time >= L.t_start && time <= L.t_end) {
      double phase_int = L.omega * time;
      double tprof = 1.0;
      if (L.t_width > 0.0) {
        double a = (time - L.t_centre) / L.t_width;
        tprof = std::exp(-(a * a));   // evaluator_blocks.F90:988-991 gauss()
      }
      double t_env = tprof * L.amp;
      for (int ir = 0; ir <= r.ny; ++ir) {
        double yv = y_grid_min_local + (double)(ir - 1) * dy;   // y(ir)
        double prof = 1.0;
        if (L.r_width > 0.0) {
          double a = (yv - 0.0) / L.r_width;
          prof = std::exp(-(a * a));
        }
        double base = t_env * prof * std::sin(phase_int + (L.phase + L.phase_curv * (yv * yv)));
        s1[ir] = s1[ir] + base * std::cos(L.pol_angle);
        s2[ir] = s2[ir] + base * std::sin(L.pol_angle);
      }
    }
  }
}

// laser.f90:411-520.  NOTE (reference quirk, reproduced): `r_d_vals` is declared (0:ny)
// but used whole-array against (1:ny) sections, so element ir pairs with r_d_vals(ir-1).
void World::outflow_bcs_x_min(Rank& r) {
  const int ny = r.ny;
  const double c = C_LIGHT;
  double dtc2 = dt * (c * c);
  double lx = dtc2 / dx, lr = dtc2 / dy;
  double sum = 1.0 / (lx + c), diff = lx - c, dt_eps = dt / EPSILON0;
  std::vector<double> s1, s2;
  laser_sources(BD_X_MIN, r, s1, s2);
  for (int im = 0; im < M; ++im)
    for (int ir = 0; ir <= ny; ++ir) r.bxm(0, ir, im) = r.bxm_x_min(ir, im);
  std::vector<double> r_d_vals(ny + 1);
  for (int ir = 0; ir <= ny; ++ir) r_d_vals[ir] = std::abs((double)(ir - 1) * dy + y_grid_min_local);
  const int qk = reference_quirks ? 1 : 0;   // the reference's (0:ny)-against-(1:ny) pairing, or element for element
  for (int im = 0; im < M; ++im) {
    std::vector<cplx> bt_new(ny + 1), br_new(ny + 1);
    for (int ir = 1; ir <= ny; ++ir) {
      cplx source_t = (im == 1) ? (cplx(s1[ir]) + IMAGI * s2[ir]) : cplx(0.0);
      bt_new[ir] = sum * (4.0 * source_t
                          + 2.0 * (r.erm_x_min(ir, im) + c * r.btm_x_min(ir, im))
                          - 2.0 * r.erm(1, ir, im)
                          + ((((IMAGI * (double)im) * (c * c)) * dt) * r.bxm(1, ir, im)) / r_d_vals[ir - qk]
                          + dt_eps * r.jrm(1, ir, im)
                          + diff * r.btm(2, ir, im));
    }
    for (int ir = 1; ir <= ny; ++ir) r.btm(1, ir, im) = bt_new[ir];
    const int ir_l = 1, ir_h = ny - 1;
    for (int ir = ir_l; ir <= ir_h; ++ir) {
      cplx source_r = (im == 1) ? ((-IMAGI) * s1[ir] + cplx(s2[ir])) : cplx(0.0);
      br_new[ir] = sum * (-4.0 * source_r
                          - 2.0 * (r.etm_x_min(ir, im) + c * r.brm_x_min(ir, im))
                          + 2.0 * r.etm(1, ir, im)
                          - lr * (r.bxm(1, ir + 1, im) - r.bxm(1, ir, im))
                          - dt_eps * r.jtm(1, ir, im)
                          + diff * r.brm(2, ir, im));
    }
    for (int ir = ir_l; ir <= ir_h; ++ir) r.brm(1, ir, im) = br_new[ir];
  }
}

// laser.f90:524-633.  Same (0:ny)-vs-(1:ny) pairing quirk for r_d_vals, and here also for
// source_t (used whole-array at :604).
void World::outflow_bcs_x_max(Rank& r) {
  const int nx = r.nx, ny = r.ny;
  const double c = C_LIGHT;
  double dtc2 = dt * (c * c);
  double lx = dtc2 / dx, lr = dtc2 / dy;
  double sum = 1.0 / (lx + c), diff = lx - c, dt_eps = dt / EPSILON0;
  std::vector<double> s1, s2;
  laser_sources(BD_X_MAX, r, s1, s2);
  for (int im = 0; im < M; ++im)
    for (int ir = 0; ir <= ny; ++ir) r.bxm(nx, ir, im) = r.bxm_x_max(ir, im);
  std::vector<double> r_d_vals(ny + 1);
  for (int ir = 0; ir <= ny; ++ir) r_d_vals[ir] = std::abs((double)(ir - 1) * dy + y_grid_min_local);
  const int qk = reference_quirks ? 1 : 0;   // the reference's (0:ny)-against-(1:ny) pairing, or element for element
  for (int im = 0; im < M; ++im) {
    std::vector<cplx> bt_new(ny + 1), br_new(ny + 1);
    for (int ir = 1; ir <= ny; ++ir) {
      cplx source_t = (im == 1) ? (cplx(s1[ir - qk]) + IMAGI * s2[ir - qk]) : cplx(0.0);
      bt_new[ir] = sum * (-4.0 * source_t
                          - 2.0 * (r.erm_x_max(ir, im) + c * r.btm_x_max(ir, im))
                          + 2.0 * r.erm(nx - 1, ir, im)
                          - ((((IMAGI * (double)im) * (c * c)) * dt) * r.bxm(nx - 1, ir, im)) / r_d_vals[ir - qk]
                          - dt_eps * r.jrm(nx - 1, ir, im)
                          + diff * r.btm(nx - 1, ir, im));
    }
    for (int ir = 1; ir <= ny; ++ir) r.btm(nx, ir, im) = bt_new[ir];
    const int ir_l = 1, ir_h = ny - 1;
    for (int ir = ir_l; ir <= ir_h; ++ir) {
      cplx source_r = (im == 1) ? ((-IMAGI) * s1[ir] + cplx(s2[ir])) : cplx(0.0);
      br_new[ir] = sum * (4.0 * source_r
                          + 2.0 * (r.etm_x_max(ir, im) + c * r.brm_x_max(ir, im))
                          - 2.0 * r.etm(nx - 1, ir, im)
                          + lr * (r.bxm(nx - 1, ir + 1, im) - r.bxm(nx - 1, ir, im))
                          + dt_eps * r.jtm(nx - 1, ir, im)
                          + diff * r.brm(nx - 1, ir, im));
    }
    for (int ir = ir_l; ir <= ir_h; ++ir) r.brm(nx, ir, im) = br_new[ir];
  }
}

void World::outflow_bcs_r_max(Rank& r) {   // laser.f90:637-690
  const int nx = r.nx, ny = r.ny;
  const double c = C_LIGHT;
  double dtc2 = dt * (c * c);
  double inv_r = 1.0 / ((double)((float)ny - 1.5f) * dy + y_grid_min_local);
  double dtc2_4r = 0.25 * dtc2 * inv_r;
  // REFERENCE QUIRK (reproduced): icdt_2r is declared REAL(num) (laser.f90:640) but assigned
  // the purely imaginary 0.5*imagi*c*dt*inv_r, so Fortran keeps only its real part = 0 and
  // the two azimuthal (im) coupling terms below vanish identically.
  double icdt_2r = ((((0.5 * IMAGI) * c) * dt) * inv_r).re;
  // reference_quirks off: the coefficient as the right-hand side spells it, purely imaginary
  const cplx icdt_2r_c = reference_quirks ? cplx(0.0) : (((0.5 * IMAGI) * c) * dt) * inv_r;
  double lx = dtc2 / dx, ly = dtc2 / dy;
  double sum_x = 1.0 / (ly + c);
  double sum_t = 1.0 / (ly + c + dtc2_4r);
  double dt_2eps = 0.5 * dt / EPSILON0;
  int ix_l = r.x_min_boundary ? 1 : 0;
  int ix_h = r.x_max_boundary ? nx - 1 : nx;
  for (int im = 0; im < M; ++im) {
    for (int ix = ix_l; ix <= ix_h; ++ix) {
      r.bxm(ix, ny, im) = sum_x * ((-r.bxm(ix, ny - 1, im)) * (c - ly)
          - r.bxm_old(ix, ny, im) * (-c + ly)
          - r.bxm_old(ix, ny - 1, im) * (-c - ly)
          - ((c * dt) * inv_r) * r.etm(ix, ny - 1, im)
          + (0.5 * lx) * (r.brm(ix + 1, ny - 1, im) - r.brm(ix, ny - 1, im)
                          + r.brm_old(ix + 1, ny - 1, im) - r.brm_old(ix, ny - 1, im))
          - (icdt_2r * (double)im) * (r.erm(ix, ny, im) + r.erm(ix, ny - 1, im))
          - dt_2eps * (r.jtm(ix, ny - 1, im) + r.jtm_old(ix, ny - 1, im)));
      if (!reference_quirks)
        r.bxm(ix, ny, im) = r.bxm(ix, ny, im)
            - sum_x * ((icdt_2r_c * (double)im) * (r.erm(ix, ny, im) + r.erm(ix, ny - 1, im)));
    }
    // RHS uses only rows ny-1 / *_old of btm, so in-place is safe
    for (int ix = 1; ix <= nx; ++ix) {
      r.btm(ix, ny, im) = sum_t * ((-r.btm(ix, ny - 1, im)) * (c - ly + dtc2_4r)
          - r.btm_old(ix, ny, im) * (-c + ly + dtc2_4r)
          - r.btm_old(ix, ny - 1, im) * (-c - ly + dtc2_4r)
          - ((0.5 * lx) / c) * (r.erm(ix, ny, im) + r.erm(ix, ny - 1, im)
                                - r.erm(ix - 1, ny, im) - r.erm(ix - 1, ny - 1, im))
          - ((icdt_2r * (double)im) * c) * (r.brm(ix, ny - 1, im) + r.brm_old(ix, ny - 1, im))
          + dt_2eps * (r.jxm(ix, ny - 1, im) + r.jxm_old(ix, ny - 1, im)));
      if (!reference_quirks)
        r.btm(ix, ny, im) = r.btm(ix, ny, im)
            - sum_t * (((icdt_2r_c * (double)im) * c) * (r.brm(ix, ny - 1, im) + r.brm_old(ix, ny - 1, im)));
    }
  }
}

void World::bfield_final_bcs() {   // boundary.F90:1505-1537
  bfield_bcs(false);
  for (Rank& r : ranks) {
    if (r.x_min_boundary) {
      if (add_laser[BD_X_MIN] || bc_field[BD_X_MIN] == BC_SIMPLE_OUTFLOW) outflow_bcs_x_min(r);
    }
    if (r.x_max_boundary) {
      if (add_laser[BD_X_MAX] || bc_field[BD_X_MAX] == BC_SIMPLE_OUTFLOW) outflow_bcs_x_max(r);
    }
    if (bc_field[BD_Y_MAX] == BC_SIMPLE_OUTFLOW) {
      outflow_bcs_r_max(r);
    } else if (bc_field[BD_Y_MAX] == BC_ZERO_B) {
      for (int im = 0; im < M; ++im)
        for (int ix = 1 - NG; ix <= r.nx + NG; ++ix) {
          r.bxm(ix, r.ny, im) = cplx(0.0);
          r.brm(ix, r.ny, im) = cplx(0.0);
          r.btm(ix, r.ny, im) = cplx(0.0);
        }
    }
  }
  bfield_bcs(true);
}

// (the `omp parallel for` over ranks plays the role of the reference's MPI ranks running
// concurrently; it is only compiled in for the CPU-baseline timing build)
void World::update_eb_fields_half() {   // fields.f90:316-337
  const int P = (int)ranks.size();
#pragma omp parallel for schedule(static) if (P > 1)
  for (int k = 0; k < P; ++k) update_e_field(ranks[k]);
  efield_bcs();
#pragma omp parallel for schedule(static) if (P > 1)
  for (int k = 0; k < P; ++k) {
    Rank& r = ranks[k];
    r.bxm_old.d = r.bxm.d;
    r.brm_old.d = r.brm.d;
    r.btm_old.d = r.btm.d;
    update_b_field(r);
  }
  bfield_bcs(true);
}

void World::update_eb_fields_final() {   // fields.f90:341-353
  const int P = (int)ranks.size();
#pragma omp parallel for schedule(static) if (P > 1)
  for (int k = 0; k < P; ++k) update_b_field(ranks[k]);
  bfield_final_bcs();
#pragma omp parallel for schedule(static) if (P > 1)
  for (int k = 0; k < P; ++k) update_e_field(ranks[k]);
  efield_bcs();
}

// ---------------------------------------------------------------------------------------
// Particle push + gather + deposit: particles.F90:163-217 (prologue), :296-479 (push),
// :515-665 (deposit); shapes include/triangle/{gx,hx_dcell,e_part,b_part}.inc
// ---------------------------------------------------------------------------------------
namespace {
inline int ifloor(double v) { return (int)std::floor(v); }
}

void World::push_rank(Rank& r) {
  const int nx = r.nx, ny = r.ny;
  const double c = C_LIGHT;
  (void)nx;
  r.jxm_old.d = r.jxm.d;
  r.jrm_old.d = r.jrm.d;
  r.jtm_old.d = r.jtm.d;
  r.jxm.zero();
  r.jrm.zero();
  r.jtm.zero();

  const double fac = SHAPE_FAC;   // particles.F90:145-153: (0.5)**c_ndims for the triangle, c_ndims = 2
  double idx = 1.0 / dx, idy = 1.0 / dy, idt = 1.0 / dt;
  double dto2 = dt / 2.0, dtco2 = c * dto2, dtfac = 0.5 * dt * fac;
  double third = 1.0 / 3.0, sixth = 0.5 * third;

  // per-radius tables (particles.F90:190-217); index iy in [1-jng, ny+jng], xt from 0-jng
  const int tlo = 0 - JNG;
  std::vector<double> inv_area_rt_v(ny + 2 * JNG + 1), inv_area_xt_v(ny + 2 * JNG + 1),
      inv_volume_v(ny + 2 * JNG + 1), ratio_v(ny + 2 * JNG + 1);
  auto T = [&](std::vector<double>& v, int iy) -> double& { return v[iy - tlo]; };
  double r_low = y_grid_min_local - (double)JNG * dy;
  T(inv_area_xt_v, 0 - JNG) = 1.0 / (2.0 * PI * std::abs(r_low) * dx);
  for (int iy = 1 - JNG; iy <= ny + JNG; ++iy) {
    if (std::lround(2.0 * r_low / dy) == -1) {
      T(inv_area_rt_v, iy) = 1.0 / (PI * ((0.5 * dy) * (0.5 * dy)));
    } else {
      T(inv_area_rt_v, iy) = 1.0 / (PI * std::abs((r_low + dy) * (r_low + dy) - r_low * r_low));
    }
    T(inv_area_xt_v, iy) = 1.0 / (2.0 * PI * std::abs(r_low + dy) * dx);
    r_low = r_low + dy;
  }
  for (int iy = 1 - JNG; iy <= ny + JNG; ++iy) {
    T(inv_volume_v, iy) = T(inv_area_rt_v, iy) / dx;
    T(ratio_v, iy) = T(inv_area_xt_v, iy) / T(inv_area_xt_v, iy - 1);
  }

  for (size_t isp = 0; isp < species.size(); ++isp) {
    const Species& sp = species[isp];
    if (sp.immobile) continue;
    const bool deposit = !sp.zero_current;
    double part_q = sp.charge, part_mc = c * sp.mass;
    double ipart_mc = 1.0 / part_mc;
    double cmratio = part_q * dtfac * ipart_mc;
    double ccmratio = c * cmratio;

    for (Particle& cur : r.parts[isp]) {
      double part_weight = cur.w;
      double part_x = cur.pos[0], part_y = cur.pos[1], part_z = cur.pos[2];
      double part_ux = cur.p[0] * ipart_mc, part_uy = cur.p[1] * ipart_mc, part_uz = cur.p[2] * ipart_mc;

      double gamma_rel = std::sqrt(part_ux * part_ux + part_uy * part_uy + part_uz * part_uz + 1.0);
      double root = dtco2 / gamma_rel;
      part_x = part_x + part_ux * root;
      part_y = part_y + part_uy * root;
      part_z = part_z + part_uz * root;

      double part_x_local = part_x - r.x_grid_min_local;
      double part_r = std::sqrt(part_y * part_y + part_z * part_z);
      double part_r_local = part_r - y_grid_min_local;

      cplx exp_min_itheta = (cplx(part_y) - IMAGI * part_z) / part_r;
      double theta_05 = std::atan2(part_z, part_y);
      cplx exp_min_itheta_05 = exp_min_itheta;
      cplx exp_itheta_05 = recip(exp_min_itheta_05);

      double cell_x_r = part_x_local * idx - SHAPE_CELL_SHIFT;   // (top-hat: - 0.5, particles.F90:336-342)
      double cell_y_r = part_r_local * idy - SHAPE_CELL_SHIFT;
      int cell_x1 = ifloor(cell_x_r + 0.5);
      double cell_frac_x = (double)cell_x1 - cell_x_r;
      cell_x1 = cell_x1 + 1;
      int cell_y1 = ifloor(cell_y_r + 0.5);
      double cell_frac_y = (double)cell_y1 - cell_y_r;
      cell_y1 = cell_y1 + 1;

      // arrays indexed -3..3 -> [k + WO] (sf_min-1 : sf_max+1 of the widest shape)
      double gx[NW] = {0, 0, 0, 0, 0, 0, 0}, gy[NW] = {0, 0, 0, 0, 0, 0, 0};
      double hx[NW] = {0, 0, 0, 0, 0, 0, 0}, hy[NW] = {0, 0, 0, 0, 0, 0, 0};
      shape_weights(cell_frac_x, 0, gx);   // <shape>/gx.inc
      shape_weights(cell_frac_y, 0, gy);

      int cell_x2 = ifloor(cell_x_r);
      cell_frac_x = (double)cell_x2 - cell_x_r + 0.5;
      cell_x2 = cell_x2 + 1;
      int cell_y2 = ifloor(cell_y_r);
      cell_frac_y = (double)cell_y2 - cell_y_r + 0.5;
      cell_y2 = cell_y2 + 1;

      int dcellx = 0, dcelly = 0;
      shape_weights(cell_frac_x, dcellx, hx);   // <shape>/hx_dcell.inc
      shape_weights(cell_frac_y, dcelly, hy);

      // gather (<shape>/e_part.inc, b_part.inc): rows in r weighted by wy, columns in x by wx, summed in the
      // includes' order (left to right)
      auto gather = [&](const Arr3& F, const double* wy, const double* wx, int cx, int cy, int im) -> cplx {
        cplx total;
        for (int iy = SF_MIN; iy <= SF_MAX; ++iy) {
          cplx sacc = wx[SF_MIN + WO] * F(cx + SF_MIN, cy + iy, im);
          for (int ix = SF_MIN + 1; ix <= SF_MAX; ++ix) sacc = sacc + wx[ix + WO] * F(cx + ix, cy + iy, im);
          const cplx row = wy[iy + WO] * sacc;
          total = (iy == SF_MIN) ? row : total + row;
        }
        return total;
      };
      // which cell pair each B component is read at: the top-hat include pairs them differently from the other
      // two (tophat/b_part.inc: bxm at (cell_x1, cell_y2), brm at (cell_x2, cell_y1), btm at (cell_x2, cell_y2))
#if CYLO_SHAPE == 1
      const int bx_cx = cell_x1, bx_cy = cell_y2, br_cx = cell_x2, br_cy = cell_y1, bt_cx = cell_x2, bt_cy = cell_y2;
#else
      const int bx_cx = cell_x2, bx_cy = cell_y1, br_cx = cell_x1, br_cy = cell_y2, bt_cx = cell_x1, bt_cy = cell_y1;
#endif
      cplx exp_min_imtheta(1.0);
      double ex_part = 0, er_part = 0, et_part = 0;
      for (int im = 0; im < M; ++im) {
        ex_part = ex_part + (exp_min_imtheta * gather(r.exm, hy, gx, cell_x1, cell_y2, im)).re;
        er_part = er_part + (exp_min_imtheta * gather(r.erm, gy, hx, cell_x2, cell_y1, im)).re;
        et_part = et_part + (exp_min_imtheta * gather(r.etm, hy, hx, cell_x2, cell_y2, im)).re;
        exp_min_imtheta = exp_min_imtheta * exp_min_itheta;
      }
      double ey_part = er_part * exp_min_itheta.re + et_part * exp_min_itheta.im;
      double ez_part = -er_part * exp_min_itheta.im + et_part * exp_min_itheta.re;

      exp_min_imtheta = cplx(1.0);
      double bx_part = 0, br_part = 0, bt_part = 0;
      for (int im = 0; im < M; ++im) {
        bx_part = bx_part + (exp_min_imtheta * gather(r.bxm, gy, hx, bx_cx, bx_cy, im)).re;
        br_part = br_part + (exp_min_imtheta * gather(r.brm, hy, gx, br_cx, br_cy, im)).re;
        bt_part = bt_part + (exp_min_imtheta * gather(r.btm, gy, gx, bt_cx, bt_cy, im)).re;
        exp_min_imtheta = exp_min_imtheta * exp_min_itheta;
      }
      double by_part = br_part * exp_min_itheta.re + bt_part * exp_min_itheta.im;
      double bz_part = -br_part * exp_min_itheta.im + bt_part * exp_min_itheta.re;

      // Boris (particles.F90:405-451)
      double uxm = part_ux + cmratio * ex_part;
      double uym = part_uy + cmratio * ey_part;
      double uzm = part_uz + cmratio * ez_part;
      if (hc_push) {
        // Higuera-Cary (particles.F90:409-421, -DHC_PUSH)
        gamma_rel = uxm * uxm + uym * uym + uzm * uzm + 1.0;
        const double alpha = 0.5 * part_q * dt / sp.mass;
        const double beta_x = alpha * bx_part, beta_y = alpha * by_part, beta_z = alpha * bz_part;
        const double beta2 = beta_x * beta_x + beta_y * beta_y + beta_z * beta_z;
        const double sigma = gamma_rel - beta2;
        const double beta_dot_u = beta_x * uxm + beta_y * uym + beta_z * uzm;
        gamma_rel = sigma + std::sqrt(sigma * sigma + 4.0 * (beta2 + beta_dot_u * beta_dot_u));
        gamma_rel = std::sqrt(0.5 * gamma_rel);
      } else {
        gamma_rel = std::sqrt(uxm * uxm + uym * uym + uzm * uzm + 1.0);
      }
      root = ccmratio / gamma_rel;
      double taux = bx_part * root, tauy = by_part * root, tauz = bz_part * root;
      double taux2 = taux * taux, tauy2 = tauy * tauy, tauz2 = tauz * tauz;
      double tau = 1.0 / (1.0 + taux2 + tauy2 + tauz2);
      double uxp = ((1.0 + taux2 - tauy2 - tauz2) * uxm
                    + 2.0 * ((taux * tauy + tauz) * uym + (taux * tauz - tauy) * uzm)) * tau;
      double uyp = ((1.0 - taux2 + tauy2 - tauz2) * uym
                    + 2.0 * ((tauy * tauz + taux) * uzm + (tauy * taux - tauz) * uxm)) * tau;
      double uzp = ((1.0 - taux2 - tauy2 + tauz2) * uzm
                    + 2.0 * ((tauz * taux + tauy) * uxm + (tauz * tauy - taux) * uym)) * tau;
      part_ux = uxp + cmratio * ex_part;
      part_uy = uyp + cmratio * ey_part;
      part_uz = uzp + cmratio * ez_part;

      double part_u2 = part_ux * part_ux + part_uy * part_uy + part_uz * part_uz;
      gamma_rel = std::sqrt(part_u2 + 1.0);
      double igamma = 1.0 / gamma_rel;
      root = dtco2 * igamma;
      double delta_x = part_ux * root, delta_y = part_uy * root, delta_z = part_uz * root;
      part_x = part_x + delta_x;
      part_y = part_y + delta_y;
      part_z = part_z + delta_z;

      cur.pos[0] = part_x; cur.pos[1] = part_y; cur.pos[2] = part_z;
      cur.p[0] = part_mc * part_ux; cur.p[1] = part_mc * part_uy; cur.p[2] = part_mc * part_uz;

      double part_vy = part_uy * c * igamma;
      double part_vz = part_uz * c * igamma;
      part_r = std::sqrt(part_y * part_y + part_z * part_z);
      cplx exp_itheta_10 = (cplx(part_y) + IMAGI * part_z) / part_r;
      double part_vt = -part_vy * exp_itheta_10.im + part_vz * exp_itheta_10.re;

      if (!deposit) continue;

      // deposit (particles.F90:515-665)
      part_x_local = part_x + delta_x - r.x_grid_min_local;
      part_y = part_y + delta_y;
      part_z = part_z + delta_z;
      part_r = std::sqrt(part_y * part_y + part_z * part_z);
      part_r_local = part_r - y_grid_min_local;
      cplx exp_itheta_15 = (cplx(part_y) + IMAGI * part_z) / part_r;
      double theta_15 = std::atan2(part_z, part_y);
      cplx exp_idtheta = exp_itheta_15 * exp_min_itheta_05;
      double dtheta = theta_15 - theta_05;

      for (int k = 0; k < NW; ++k) { gx[k] = hx[k]; gy[k] = hy[k]; }

      cell_x_r = part_x_local * idx - SHAPE_CELL_SHIFT;
      cell_y_r = part_r_local * idy - SHAPE_CELL_SHIFT;
      int cell_x3 = ifloor(cell_x_r);
      cell_frac_x = (double)cell_x3 - cell_x_r + 0.5;
      cell_x3 = cell_x3 + 1;
      int cell_y3 = ifloor(cell_y_r);
      cell_frac_y = (double)cell_y3 - cell_y_r + 0.5;
      cell_y3 = cell_y3 + 1;

      for (int k = 0; k < NW; ++k) { hx[k] = 0.0; hy[k] = 0.0; }
      dcellx = cell_x3 - cell_x2;
      dcelly = cell_y3 - cell_y2;
      shape_weights(cell_frac_x, dcellx, hx);
      shape_weights(cell_frac_y, dcelly, hy);
      for (int k = 0; k < NW; ++k) { hx[k] = hx[k] - gx[k]; hy[k] = hy[k] - gy[k]; }

      int xmin = SF_MIN + (dcellx - 1) / 2, xmax = SF_MAX + (dcellx + 1) / 2;   // truncating division
      int ymin = SF_MIN + (dcelly - 1) / 2, ymax = SF_MAX + (dcelly + 1) / 2;

      double q_weight_fac = part_q * part_weight * fac;
      double fcx = q_weight_fac * idt;
      double fcy = q_weight_fac * idt;
      double fcz = q_weight_fac * part_vt;

      cplx exp_imtheta0(1.0), exp_imdtheta(1.0);
      cplx m_fac_1, m_fac_2, m_fac_3, m_fac_4;
      for (int im = 0; im < M; ++im) {
        double mdth = (double)im * dtheta;
        double m2dth2 = mdth * mdth;
        bool small = std::abs(mdth) < taylor_switch;   // particles.F90:593 (1.0e-4; a test knob moves it)
        double inv_mdth = 0.0, inv_m2dth2 = 0.0;
        if (!small && im > 0) {
          inv_mdth = 1.0 / mdth;
          inv_m2dth2 = inv_mdth * inv_mdth;
        }
        if (im == 0) {
          exp_imtheta0 = cplx(1.0);
          exp_imdtheta = cplx(1.0);
        } else {
          exp_imtheta0 = exp_imtheta0 * exp_itheta_05;
          exp_imdtheta = exp_imdtheta * exp_idtheta;
          if (small) {
            m_fac_1 = 2.0 * exp_imtheta0;
            m_fac_2 = m_fac_1 * ((cplx(1.0) + (0.5 * IMAGI) * mdth) - cplx(sixth * m2dth2));
            m_fac_3 = m_fac_1 * ((cplx(0.5) + (third * IMAGI) * mdth) - cplx(0.125 * m2dth2));
            m_fac_4 = m_fac_1 * ((cplx(third) + (0.25 * IMAGI) * mdth) - cplx(0.1 * m2dth2));
          } else {
            m_fac_1 = (2.0 * inv_mdth) * exp_imtheta0;
            m_fac_2 = m_fac_1 * ((-IMAGI) * (exp_imdtheta - cplx(1.0)));
            m_fac_3 = m_fac_1 * (inv_mdth * (exp_imdtheta * (cplx(1.0) - IMAGI * mdth) - cplx(1.0)));
            m_fac_4 = m_fac_1 * ((IMAGI * inv_m2dth2) *
                                 (exp_imdtheta * ((cplx(-m2dth2) - (2.0 * IMAGI) * mdth) + cplx(2.0)) - cplx(2.0)));
          }
        }
        cplx jyh[NW];
        for (int iy = ymin; iy <= ymax; ++iy) {
          int cy = cell_y2 + iy;
          cplx w_rt, ym_fac_1;
          if (im == 0) {
            w_rt = cplx(gy[iy + WO] + 0.5 * hy[iy + WO]);
            ym_fac_1 = cplx(0.5 * gy[iy + WO] + third * hy[iy + WO]);
          } else {
            w_rt = m_fac_2 * gy[iy + WO] + m_fac_3 * hy[iy + WO];
            ym_fac_1 = m_fac_3 * gy[iy + WO] + m_fac_4 * hy[iy + WO];
          }
          double fjx = fcx * T(inv_area_rt_v, cy);
          double fjy = fcy * hy[iy + WO] * T(inv_area_xt_v, cy);
          double fjz = fcz * T(inv_volume_v, cy);
          cplx jxh(0.0);
          for (int ix = xmin; ix <= xmax; ++ix) {
            int cx = cell_x2 + ix;
            cplx w_xt;
            if (im == 0) w_xt = cplx(gx[ix + WO] + 0.5 * hx[ix + WO]);
            else w_xt = m_fac_2 * gx[ix + WO] + m_fac_3 * hx[ix + WO];
            cplx w_xr = gx[ix + WO] * w_rt + hx[ix + WO] * ym_fac_1;
            jxh = jxh - (fjx * hx[ix + WO]) * w_rt;
            jyh[ix + WO] = jyh[ix + WO] * T(ratio_v, cy) - fjy * w_xt;
            cplx jzh = fjz * w_xr;
            r.jxm(cx + 1, cy, im) += jxh;
            r.jrm(cx, cy + 1, im) += jyh[ix + WO];
            r.jtm(cx, cy, im) += jzh;
          }
        }
      }
    }
    // current_bcs(species) is a no-op unless species BCs are mixed (boundary.F90:1337-1339)
  }
  current_bcs_r_min_final(r);
}

void World::push_particles() {
  const int P = (int)ranks.size();
#pragma omp parallel for schedule(static) if (P > 1)
  for (int k = 0; k < P; ++k) push_rank(ranks[k]);
  particle_bcs();
}

void World::current_bcs_r_min_final(Rank& r) {   // boundary.F90:1909-1959
  const int lo = 1 - NG, hi = r.nx + NG;
  double mode_sign = 1.0;
  for (int im = 0; im < M; ++im) {
    for (int j = 2; j <= JNG; ++j) {
      for (int ix = lo; ix <= hi; ++ix) {
        r.jxm(ix, j - 1, im) = r.jxm(ix, j - 1, im) + mode_sign * r.jxm(ix, 1 - j, im);
        r.jrm(ix, j, im) = r.jrm(ix, j, im) - mode_sign * r.jrm(ix, 1 - j, im);
        r.jtm(ix, j - 1, im) = r.jtm(ix, j - 1, im) - mode_sign * r.jtm(ix, 1 - j, im);
        r.jxm(ix, 1 - j, im) = cplx(0.0);
        r.jrm(ix, 1 - j, im) = cplx(0.0);
        r.jtm(ix, 1 - j, im) = cplx(0.0);
      }
    }
    for (int ix = lo; ix <= hi; ++ix) r.jrm(ix, 1, im) = r.jrm(ix, 1, im) - mode_sign * r.jrm(ix, 0, im);
    mode_sign = -mode_sign;
  }
  for (int im = 0; im < M; ++im) {
    for (int ix = lo; ix <= hi; ++ix) {
      if (im > 0) r.jxm(ix, 0, im) = cplx(0.0);
      else r.jxm(ix, 0, im) = (4.0 * r.jxm(ix, 1, im) - r.jxm(ix, 2, im)) / 3.0;
      if (im == 1) {
        r.jtm(ix, 0, im) = ((-IMAGI) * (9.0 * r.jrm(ix, 1, im) - r.jrm(ix, 2, im))) / 8.0;
        r.jrm(ix, 0, im) = (2.0 * IMAGI) * r.jtm(ix, 0, im) - r.jrm(ix, 1, im);
      } else {
        r.jtm(ix, 0, im) = cplx(0.0);
        r.jrm(ix, 0, im) = -r.jrm(ix, 1, im);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// particle_bcs: boundary.F90:1541-1889 (non-cpml, non-thermal branches), exchange order
// partlist.F90:822-876 via the (ix,iy) neighbour loop at boundary.F90:1867-1877.
// ---------------------------------------------------------------------------------------
void World::particle_bcs() {
  const int P = (int)ranks.size();
  double boundary_shift = dx * (double)((1 + PNG + 0) / 2);
  double x_min_outer = x_min - boundary_shift;
  double x_max_outer = x_max + boundary_shift;
  double x_shift = length_x;
  boundary_shift = dy * (double)((1 + PNG + 0) / 2);
  double y_max_outer = y_max + boundary_shift;
  double y_max_local = y_max;   // nprocy = 1

  for (Rank& r : ranks) { r.n_sent_left = r.n_sent_right = r.n_removed = r.n_recv = 0; }

  for (size_t isp = 0; isp < species.size(); ++isp) {
    const int* bc_species = species[isp].bc_particle;
    std::vector<std::vector<Particle>> send_l(P), send_r(P);
#pragma omp parallel for schedule(static) if (P > 1)
    for (int k = 0; k < P; ++k) {
      Rank& r = ranks[k];
      std::vector<Particle>& pl = r.parts[isp];
      std::vector<Particle> keep;
      keep.reserve(pl.size());
      int bc = -1;   // stale across particles, as in the reference
      for (Particle& cur : pl) {
        int xbd = 0;
        bool out_of_bounds = false;
        double part_pos = cur.pos[0];
        int sgn = -1;
        if (part_pos < r.x_min_local) {
          xbd = sgn;
          if (r.x_min_boundary) {
            xbd = 0;
            bc = bc_species[BD_X_MIN];
            if (bc == BC_REFLECT) {
              cur.pos[0] = 2.0 * x_min - part_pos;
              cur.p[0] = -cur.p[0];
            } else if (bc == BC_PERIODIC) {
              xbd = sgn;
              cur.pos[0] = part_pos - (double)sgn * x_shift;
            }
          }
          if (part_pos < x_min_outer && bc != BC_PERIODIC) out_of_bounds = true;
        }
        sgn = 1;
        if (part_pos >= r.x_max_local) {
          xbd = sgn;
          if (r.x_max_boundary) {
            xbd = 0;
            bc = bc_species[BD_X_MAX];
            if (bc == BC_REFLECT) {
              cur.pos[0] = 2.0 * x_max - part_pos;
              cur.p[0] = -cur.p[0];
            } else if (bc == BC_PERIODIC) {
              xbd = sgn;
              cur.pos[0] = part_pos - (double)sgn * x_shift;
            }
          }
          if (part_pos >= x_max_outer && bc != BC_PERIODIC) out_of_bounds = true;
        }
        part_pos = std::sqrt(cur.pos[1] * cur.pos[1] + cur.pos[2] * cur.pos[2]);
        if (part_pos >= y_max_local) {
          bc = bc_species[BD_Y_MAX];
          if (bc == BC_REFLECT) {
            double radial_reduction = 2.0 * y_max / part_pos - 1.0;
            cur.pos[1] = cur.pos[1] * radial_reduction;
            cur.pos[2] = cur.pos[2] * radial_reduction;
            double inv_final_r = 1.0 / std::sqrt(cur.pos[1] * cur.pos[1] + cur.pos[2] * cur.pos[2]);
            double cos_theta = cur.pos[1] * inv_final_r;
            double sin_theta = cur.pos[2] * inv_final_r;
            double part_pr = cur.p[1] * cos_theta + cur.p[2] * sin_theta;
            double part_pt = -cur.p[1] * sin_theta + cur.p[2] * cos_theta;
            cur.p[1] = -part_pr * cos_theta - part_pt * sin_theta;
            cur.p[2] = -part_pr * sin_theta + part_pt * cos_theta;
          }
          if (part_pos >= y_max_outer && bc != BC_PERIODIC) out_of_bounds = true;
        }
        if (out_of_bounds) {
          r.n_removed++;
        } else if (xbd == -1) {
          send_l[k].push_back(cur);
          r.n_sent_left++;
        } else if (xbd == 1) {
          send_r[k].push_back(cur);
          r.n_sent_right++;
        } else {
          keep.push_back(cur);
        }
      }
      pl.swap(keep);
    }
    // exchange: (ix=-1) send left, receive from the right neighbour; then (ix=+1)
    const bool per_min = (bc_field[BD_X_MIN] == BC_PERIODIC);   // Cartesian comm periodicity
    const bool per_max = (bc_field[BD_X_MAX] == BC_PERIODIC);
    for (int k = 0; k < P; ++k) {
      Rank& r = ranks[k];
      int right = (k + 1 < P) ? k + 1 : (per_max ? 0 : -1);
      int left = (k - 1 >= 0) ? k - 1 : (per_min ? P - 1 : -1);
      if (right >= 0) {
        auto& v = send_l[right];
        // what `right` sent to ITS left neighbour arrives here only if that neighbour is me
        int rl = (right - 1 >= 0) ? right - 1 : (per_min ? P - 1 : -1);
        if (rl == k) {
          r.parts[isp].insert(r.parts[isp].end(), v.begin(), v.end());
          r.n_recv += (int64_t)v.size();
        }
      }
      if (left >= 0) {
        auto& v = send_r[left];
        int lr = (left + 1 < P) ? left + 1 : (per_max ? 0 : -1);
        if (lr == k) {
          r.parts[isp].insert(r.parts[isp].end(), v.begin(), v.end());
          r.n_recv += (int64_t)v.size();
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Current boundary conditions: boundary.F90:1893-1905 -> :1330-1351 -> :918-1015, :1133-1245
// ---------------------------------------------------------------------------------------
void World::reflection_bcs(Rank& r, Arr3& a, int im, int flip_dir) {
  const int nx = r.nx, ny = r.ny;
  const int jlo = 1 - NG, jhi = ny + NG;
  int bc = bc_allspecies(BD_X_MIN);
  if (r.x_min_boundary && bc == BC_REFLECT) {
    if (flip_dir == 1) {
      for (int i = 1; i <= NG - 1; ++i)
        for (int j = jlo; j <= jhi; ++j) {
          a(i, j, im) = a(i, j, im) - a(1 - i, j, im);
          a(1 - i, j, im) = cplx(0.0);
        }
    } else {
      for (int i = 1; i <= NG - 1; ++i)
        for (int j = jlo; j <= jhi; ++j) {
          a(i, j, im) = a(i, j, im) + a(-i, j, im);
          a(-i, j, im) = cplx(0.0);
        }
    }
  }
  bc = bc_allspecies(BD_X_MAX);
  int nn = nx;
  if (r.x_max_boundary && bc == BC_REFLECT) {
    if (flip_dir == 1) {
      for (int i = 1; i <= NG; ++i)
        for (int j = jlo; j <= jhi; ++j) {
          a(nn + 1 - i, j, im) = a(nn + 1 - i, j, im) - a(nn + i, j, im);
          a(nn + i, j, im) = cplx(0.0);
        }
    } else {
      for (int i = 1; i <= NG; ++i)
        for (int j = jlo; j <= jhi; ++j) {
          a(nn - i, j, im) = a(nn - i, j, im) + a(nn + i, j, im);
          a(nn + i, j, im) = cplx(0.0);
        }
    }
  }
  nn = ny;
  bc = bc_allspecies(BD_Y_MAX);
  if (bc == BC_REFLECT) {
    const int ilo = 1 - NG, ihi = nx + NG;
    if (flip_dir == 2) {
      double r_max = y_grid_min_local + ((double)ny - 0.5) * dy;
      for (int i = 1; i <= NG; ++i) {
        double ratio_num = r_max + ((double)i - 0.5) * dy, ratio_den = r_max - ((double)i - 0.5) * dy;
        for (int ix = ilo; ix <= ihi; ++ix) {
          a(ix, nn + 1 - i, im) = a(ix, nn + 1 - i, im) - (a(ix, nn + i, im) * ratio_num) / ratio_den;
          a(ix, nn + i, im) = cplx(0.0);
        }
      }
    } else if (flip_dir == 1) {
      double r_max = y_grid_min_local + ((double)ny - 0.5) * dy;
      for (int i = 1; i <= NG; ++i) {
        double ratio_num = r_max + (double)i * dy, ratio_den = r_max - (double)i * dy;
        for (int ix = ilo; ix <= ihi; ++ix) {
          a(ix, nn - i, im) = a(ix, nn - i, im) + (a(ix, nn + i, im) * ratio_num) / ratio_den;
          a(ix, nn + i, im) = cplx(0.0);
        }
      }
    } else {
      for (int i = 1; i <= NG; ++i)
        for (int ix = ilo; ix <= ihi; ++ix) {
          a(ix, nn - i, im) = a(ix, nn - i, im) + a(ix, nn + i, im);
          a(ix, nn + i, im) = cplx(0.0);
        }
    }
  }
}

void World::periodic_sum_x(Arr3 Rank::*f) {   // boundary.F90:1133-1203, x part, all modes
  const int P = (int)ranks.size();
  const bool per_min = (bc_field[BD_X_MIN] == BC_PERIODIC);
  const bool per_max = (bc_field[BD_X_MAX] == BC_PERIODIC);
  const bool sum_min = (bc_allspecies(BD_X_MIN) == BC_PERIODIC);
  const bool sum_max = (bc_allspecies(BD_X_MAX) == BC_PERIODIC);
  // step 1: my columns 1..ng += left neighbour's ghost columns nx+1..nx+ng
  std::vector<std::vector<cplx>> temp(P);
  for (int k = 0; k < P; ++k) {
    Rank& r = ranks[k];
    Arr3& a = r.*f;
    int left = (k - 1 >= 0) ? k - 1 : (per_min ? P - 1 : -1);
    // the sender (left) sends to its x_max neighbour unless it is on the boundary with a
    // non-periodic particle bc; the receiver receives from its x_min neighbour likewise
    bool recv_ok = !(r.x_min_boundary && !sum_min);
    temp[k].assign((size_t)NG * (r.ny + 2 * NG) * M, cplx());
    if (left >= 0 && recv_ok) {
      Rank& s = ranks[left];
      bool send_ok = !(s.x_max_boundary && !sum_max);
      if (send_ok) {
        Arr3& b = s.*f;
        size_t n = 0;
        for (int im = 0; im < M; ++im)
          for (int j = 1 - NG; j <= r.ny + NG; ++j)
            for (int i = 1; i <= NG; ++i) temp[k][n++] = b(s.nx + i, j, im);
      }
    }
    (void)a;
  }
  for (int k = 0; k < P; ++k) {
    Rank& r = ranks[k];
    Arr3& a = r.*f;
    size_t n = 0;
    for (int im = 0; im < M; ++im)
      for (int j = 1 - NG; j <= r.ny + NG; ++j)
        for (int i = 1; i <= NG; ++i) { a(i, j, im) = a(i, j, im) + temp[k][n++]; }
  }
  // step 2: my columns nx+1-ng..nx += right neighbour's ghost columns 1-ng..0
  for (int k = 0; k < P; ++k) {
    Rank& r = ranks[k];
    int right = (k + 1 < P) ? k + 1 : (per_max ? 0 : -1);
    bool recv_ok = !(r.x_max_boundary && !sum_max);
    temp[k].assign((size_t)NG * (r.ny + 2 * NG) * M, cplx());
    if (right >= 0 && recv_ok) {
      Rank& s = ranks[right];
      bool send_ok = !(s.x_min_boundary && !sum_min);
      if (send_ok) {
        Arr3& b = s.*f;
        size_t n = 0;
        for (int im = 0; im < M; ++im)
          for (int j = 1 - NG; j <= r.ny + NG; ++j)
            for (int i = 1; i <= NG; ++i) temp[k][n++] = b(i - NG, j, im);
      }
    }
  }
  for (int k = 0; k < P; ++k) {
    Rank& r = ranks[k];
    Arr3& a = r.*f;
    size_t n = 0;
    for (int im = 0; im < M; ++im)
      for (int j = 1 - NG; j <= r.ny + NG; ++j)
        for (int i = 1; i <= NG; ++i) { a(r.nx - NG + i, j, im) = a(r.nx - NG + i, j, im) + temp[k][n++]; }
  }
}

void World::current_bcs() {
  // mixed species boundary conditions are not supported by the oracle (none of the
  // configurations in scope use them)
  for (int i = 0; i < 4; ++i) assert(bc_allspecies(i) != BC_MIXED);
  for (Rank& r : ranks)
    for (int im = 0; im < M; ++im) {
      reflection_bcs(r, r.jxm, im, 1);
      reflection_bcs(r, r.jrm, im, 2);
      reflection_bcs(r, r.jtm, im, 3);
    }
  periodic_sum_x(&Rank::jxm);
  periodic_sum_x(&Rank::jrm);
  periodic_sum_x(&Rank::jtm);
}

void World::current_finish() {   // current_smooth.F90:29-45
  current_bcs();
  halo_x(&Rank::jxm, 0, 0);
  halo_x(&Rank::jrm, 0, 0);
  halo_x(&Rank::jtm, 0, 0);
  if (smooth_currents) {   // smooth_current, current_smooth.F90:49-57
    smooth_mode_array(&Rank::jxm);
    smooth_mode_array(&Rank::jrm);
    smooth_mode_array(&Rank::jtm);
  }
}

// current_smooth.F90:145-196, strided compensated binomial filter.  Restated with its two
// oddities: beta is computed once from the initial alpha = 0.5 and kept when alpha changes, and
// alpha changes at the END of iteration its+1, so a single compensation pass (comp_its = 1)
// still runs with alpha = 0.5.  Strides up to ng (sng <= jng, so ng_l = jng) are supported.
void World::smooth_mode_array(Arr3 Rank::*f) {
  std::vector<int> stride_inner = smooth_strides.empty() ? std::vector<int>{1} : smooth_strides;
  double alpha = 0.5;
  const double beta = (1.0 - alpha) * 0.25;
  for (Rank& r : ranks) {
    r.wk.alloc(r.nx, r.ny, M);
    r.wk.d = (r.*f).d;
  }
  for (int it = 1; it <= smooth_its + smooth_comp_its; ++it) {
    for (int cstride : stride_inner) {
      halo_x(&Rank::wk, 0, 0);   // field_mode_bc(wk_array, ng_l)
      for (Rank& r : ranks) {
        Arr3& a = r.*f;
        const Arr3& w = r.wk;
        for (int im = 0; im < M; ++im)
          for (int iy = 1; iy <= r.ny; ++iy)
            for (int ix = 1; ix <= r.nx; ++ix)
              a(ix, iy, im) = alpha * w(ix, iy, im) +
                              (w(ix - cstride, iy, im) + w(ix + cstride, iy, im) + w(ix, iy - cstride, im) +
                               w(ix, iy + cstride, im)) * beta;
        for (int im = 0; im < M; ++im)
          for (int iy = 1; iy <= r.ny; ++iy)
            for (int ix = 1; ix <= r.nx; ++ix) r.wk(ix, iy, im) = a(ix, iy, im);
      }
    }
    if (it > smooth_its) alpha = (double)smooth_its * 0.5 + 1.0;
  }
  for (Rank& r : ranks) {
    Arr3& a = r.*f;
    for (int im = 0; im < M; ++im)
      for (int iy = 1; iy <= r.ny; ++iy)
        for (int ix = 1; ix <= r.nx; ++ix) a(ix, iy, im) = r.wk(ix, iy, im);
  }
}

// ---------------------------------------------------------------------------------------
// calc_number_density_modes, calc_df.F90:588-661: per-mode number density on the cell centres.
// particle_to_grid.inc + triangle/gxfac.inc weights (with the r < dy fold onto the axis cell),
// number density = weight / macro-particle volume 2 pi dx dy r (partlist.F90:999-1013), mode
// factor 1 (m = 0) or 2 e^{i m theta}; then calc_boundary_modes (real-valued
// processor_summation_bcs on the real and imaginary parts: particle_reflection_bcs
// boundary.F90:833-914, particle_periodic_bcs :1019-1129) and field_mode_zero_gradient
// (:654-707, centred stagger) on all four boundaries.  Result in Rank::wk.
// ---------------------------------------------------------------------------------------
void World::calc_number_density_modes(int current_species) { density_deposit_and_bcs(current_species, false); }

// calc_charge_density, calc_df.F90:442-519: the same deposit of wdata = charge * weight into a real
// array (no azimuthal factors), calc_boundary + field_zero_gradient.  Result: real part of mode 0.
void World::calc_charge_density(int current_species) { density_deposit_and_bcs(current_species, true); }

void World::density_deposit_and_bcs(int current_species, bool charge) {
  for (int i = 0; i < 4; ++i) assert(bc_allspecies(i) != BC_MIXED);
  const bool spec_sum = current_species < 0;
  for (Rank& r : ranks) {
    r.wk.alloc(r.nx, r.ny, M);
    for (size_t isp = 0; isp < species.size(); ++isp) {
      if (!spec_sum && (int)isp != current_species) continue;
      if (spec_sum && species[isp].zero_current) continue;
      for (const Particle& p : r.parts[isp]) {
        const double part_r = std::sqrt(p.pos[1] * p.pos[1] + p.pos[2] * p.pos[2]);
        int cell_x, cell_y;
        double gx[NW], gy[NW];
        particle_to_grid(p.pos[0] - r.x_grid_min_local, part_r - y_grid_min_local, part_r, dx, dy, &cell_x, &cell_y, gx, gy);
        const double macro_part_volume = 2.0 * PI * dx * dy * part_r;
        const double wdata = charge ? species[isp].charge * p.w : p.w;
        const double part_num_dens = wdata / macro_part_volume;
        const cplx exp_itheta = cplx(p.pos[1], p.pos[2]) / part_r;
        cplx exp_imtheta = cplx(1.0);
        for (int im = 0; im < (charge ? 1 : M); ++im) {
          cplx mode_fac;
          if (im == 0) {
            mode_fac = cplx(1.0);
          } else {
            exp_imtheta = exp_imtheta * exp_itheta;
            mode_fac = 2.0 * exp_imtheta;
          }
          for (int iy = SF_MIN; iy <= SF_MAX; ++iy)
            for (int ix = SF_MIN; ix <= SF_MAX; ++ix)
              r.wk(cell_x + ix, cell_y + iy, im) =
                  r.wk(cell_x + ix, cell_y + iy, im) + ((gx[ix + WO] * gy[iy + WO]) * part_num_dens) * mode_fac;
        }
      }
    }
  }
  moment_summation_bcs(&Rank::wk);   // calc_boundary_modes (cyl_moments.cpp)
  centre_zero_gradient(&Rank::wk);   // field_mode_zero_gradient, c_stagger_centre, boundaries 1..4
}

// ---------------------------------------------------------------------------------------
// Moving window: window.F90:62-153, :157-300, :304-376
// ---------------------------------------------------------------------------------------
void World::insert_particles(Rank& r) {
  if (!r.x_max_boundary) return;
  double x_grid_max = x_grid_min + (double)(cfg.nx_global - 1) * dx;
  for (size_t isp = 0; isp < species.size(); ++isp) {
    const Species& s = species[isp];
    int64_t npart_per_cell = (int64_t)std::floor(s.npart_per_cell);
    double npart_frac = s.npart_per_cell - (double)npart_per_cell;
    double x0 = x_grid_max + 0.5 * dx;
    for (int iy = 1; iy <= r.ny; ++iy) {
      int64_t n_frac = 0;
      if (npart_frac > 0.0) {
        if (r.rng.uniform() < npart_frac) n_frac = 1;
      }
      for (int64_t ip = 1; ip <= npart_per_cell + n_frac; ++ip) {
        Particle p;
        double cell_frac_y = 0.5 - r.rng.uniform();
        double yc = y_grid_min_local + (double)(iy - 1) * dy;
        double part_r = yc - cell_frac_y * dy;
        double part_theta = 2.0 * PI * r.rng.uniform();
        p.pos[0] = x0 + r.rng.uniform() * dx;
        p.pos[1] = part_r * std::cos(part_theta);
        p.pos[2] = part_r * std::sin(part_theta);
        double wdata = (2.0 * PI * dx * dy * part_r) / (double)(npart_per_cell + n_frac);
        double cy2 = cell_frac_y * cell_frac_y;
        double gy[3];
        gy[0] = 0.5 * (0.25 + cy2 + cell_frac_y);
        gy[1] = 0.75 - cy2;
        gy[2] = 0.5 * (0.25 + cy2 - cell_frac_y);
        for (int i = 0; i < 3; ++i) {
          double temp_local = 0.0, drift_local = 0.0;
          for (int k = 0; k < 3; ++k) {
            temp_local = temp_local + gy[k] * s.temp[i];
            drift_local = drift_local + gy[k] * s.drift[i];
          }
          double stdev = std::sqrt(temp_local * KB * s.mass);
          p.p[i] = r.rng.box_muller(stdev, drift_local);
        }
        double weight_local = 0.0;
        for (int k = 0; k < 3; ++k) weight_local = weight_local + gy[k] * s.density;
        p.w = weight_local * wdata;
        r.parts[isp].push_back(p);
      }
    }
  }
}

void World::shift_fields() {   // window.F90:98-153
  Arr3 Rank::*all[] = {&Rank::exm, &Rank::erm, &Rank::etm, &Rank::bxm, &Rank::brm, &Rank::btm,
                       &Rank::jxm, &Rank::jrm, &Rank::jtm};
  for (auto f : all) {
    for (Rank& r : ranks) {
      Arr3& a = r.*f;
      for (int im = 0; im < M; ++im)
        for (int j = 1 - NG; j <= r.ny + NG; ++j)
          for (int i = 1 - NG; i <= r.nx + NG - 1; ++i) a(i, j, im) = a(i + 1, j, im);
    }
    halo_x(f, 0, 0);
  }
  for (Rank& r : ranks) {
    if (!r.x_max_boundary) continue;
    const int nx = r.nx;
    for (int im = 0; im < M; ++im)
      for (int j = 1 - NG; j <= r.ny + NG; ++j) {
        r.exm(nx + 1, j, im) = r.exm_x_max(j, im);
        r.erm(nx, j, im) = r.erm_x_max(j, im);
        r.etm(nx, j, im) = r.etm_x_max(j, im);
        r.exm(nx, j, im) = 0.5 * (r.exm(nx - 1, j, im) + r.exm(nx + 1, j, im));
        r.erm(nx - 1, j, im) = 0.5 * (r.erm(nx - 2, j, im) + r.erm(nx, j, im));
        r.etm(nx - 1, j, im) = 0.5 * (r.etm(nx - 2, j, im) + r.etm(nx, j, im));
        r.bxm(nx, j, im) = r.bxm_x_max(j, im);
        r.brm(nx + 1, j, im) = r.brm_x_max(j, im);
        r.btm(nx + 1, j, im) = r.btm_x_max(j, im);
        r.bxm(nx - 1, j, im) = 0.5 * (r.bxm(nx - 2, j, im) + r.bxm(nx, j, im));
        r.brm(nx, j, im) = 0.5 * (r.brm(nx - 1, j, im) + r.brm(nx + 1, j, im));
        r.btm(nx, j, im) = 0.5 * (r.btm(nx - 1, j, im) + r.btm(nx + 1, j, im));
      }
  }
}

void World::moving_window() {
  if (!cfg.move_window) return;
  if (!window_started) {
    if (time >= cfg.window_start_time && time < cfg.window_stop_time) {
      bc_field[BD_X_MIN] = cfg.bc_x_min_after_move;
      bc_field[BD_X_MAX] = cfg.bc_x_max_after_move;
      setup_boundaries();
      window_shift_fraction = 0.0;
      window_started = true;
    }
  }
  if (!window_started) return;
  if (time >= cfg.window_stop_time) return;
  if (cfg.window_v_x <= 0.0) return;
  window_shift_fraction = window_shift_fraction + dt * cfg.window_v_x / dx;
  int window_shift_cells = (int)std::floor(window_shift_fraction);
  if (window_shift_cells > 0) {
    double window_shift_real = (double)window_shift_cells;
    for (int iw = 0; iw < window_shift_cells; ++iw) {   // shift_window, window.F90:62-94
      for (Rank& r : ranks) {
        if (counter_insert) insert_particles_counter(r);
        else insert_particles(r);
      }
      x_grid_min = x_grid_min + dx;       // x_global(1) + dx
      xb_min = xb_min + dx;               // xb_global(1) + dx
      x_min = xb_min;
      x_max = xb_min + (double)cfg.nx_global * dx;   // xb_global(nx_global+1)
      setup_grid_x();
      Rank& r0 = ranks[0];                // remove_particles, window.F90:304-325
      for (auto& pl : r0.parts) {
        std::vector<Particle> keep;
        keep.reserve(pl.size());
        for (Particle& p : pl)
          if (!(p.pos[0] < x_min)) keep.push_back(p);
        pl.swap(keep);
      }
      shift_fields();
      window_shifts_total++;
    }
    particle_bcs();
    window_shift_fraction = window_shift_fraction - window_shift_real;
  }
}

void World::step_once() {   // epoch2d.F90:189-266 loop body with all optional physics off
  update_eb_fields_half();
  push_particles();
  current_finish();
  step = step + 1;
  time = time + dt / 2.0;
  for (Rank& r : ranks) r.rng.flush_cache();   // output_routines -> diagnostics.F90:235
  time = time + dt / 2.0;
  update_eb_fields_final();
  moving_window();
}

}  // namespace cylo
