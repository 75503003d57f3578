// driver.cu -- the body of the reference's main loop (epoch2d.F90:189-266) and the host-side pieces it calls
// on the way, natively: time / step bookkeeping, the laser source evaluation of outflow_bcs_x_min / x_max
// (laser.f90:276-328,442-461,556-575), the moving-window trigger (window.F90:330-376), shift_window's grid
// update (window.F90:62-94, utilities.f90:343-372) and the set-up half step (epoch2d.F90:143-161).  These are
// host work in the reference too; they live here so that a step costs the host a handful of launches and no
// interpreter time -- with x-slabs over 8 GPUs a step is a few milliseconds and the host must stay ahead of the
// device.  Everything goes through the public entry points of api.cu, in the reference's order.
// Product code: never includes, links or calls anything under oracle/.
#include <cmath>
#include <cstring>
#include <vector>

#include "ctx.cuh"

namespace cylgpu {

struct Driver {
  cylgpu_driver_config cfg;
  std::vector<cylgpu_laser> lasers;
  // grid scalars of this rank (decomp.SlabGrid)
  double dx = 0.0, x_grid_min = 0.0, xb_min = 0.0, x_min = 0.0, x_max = 0.0;
  double x_grid_min_local = 0.0, x_min_local = 0.0, x_max_local = 0.0;
  int nx_global = 0, cell_x_min = 1, cell_x_max = 1;
  // run state
  double time = 0.0, dt = 0.0;
  int64_t step = 0;
  int raw_bc_field[4] = {0, 0, 0, 0};
  int bc_field[4] = {0, 0, 0, 0};
  bool add_laser[4] = {false, false, false, false};
  bool window_started = false;
  double window_shift_fraction = 0.0;
  int64_t window_shifts_total = 0;
  std::vector<double> src[4];   // source1 / source2 on x_min, then x_max, ir = 0..ny
  std::vector<double> prof_density, prof_temp, prof_drift;
};

static Driver* drv(cylgpu_ctx* c) { return static_cast<Driver*>(c->driver); }

// setup_boundaries, boundary.F90:30-75 (field part): other / reflect -> clamp, open -> simple_outflow
static void normalise_bc_field(Driver& D) {
  for (int i = 0; i < 4; ++i) {
    int b = D.raw_bc_field[i];
    D.add_laser[i] = false;
    if (b == 2 /* c_bc_other */) b = CYLGPU_BC_CLAMP;
    if (b == CYLGPU_BC_SIMPLE_LASER) D.add_laser[i] = true;
    if (b == CYLGPU_BC_REFLECT) b = CYLGPU_BC_CLAMP;
    if (b == CYLGPU_BC_OPEN) b = CYLGPU_BC_SIMPLE_OUTFLOW;
    D.bc_field[i] = b;
  }
}

// setup_grid_x, utilities.f90:343-372 with cpml offsets = 0
static void setup_grid_x(Driver& D) {
  D.x_grid_min_local = D.x_grid_min + (double)(D.cell_x_min - 1) * D.dx;
  const double x_grid_max_local = D.x_grid_min + (double)(D.cell_x_max - 1) * D.dx;
  D.x_min_local = D.x_grid_min_local + (0 - 0.5) * D.dx;
  D.x_max_local = x_grid_max_local - (0 - 0.5) * D.dx;
}

// source1 / source2 of one x boundary on ir = 0..ny (laser.f90:442-461 and :556-575): the temporal profile
// gauss(time, t_centre, t_width) (evaluator_blocks.F90:988-991), the radial profile gauss(y, 0, r_width) on
// y = y_grid_min_local + (ir - 1) dy, phase = omega * time (laser.f90:287) + phase(y); libm's sin / exp / cos
// element by element, as the Fortran intrinsics
static void laser_sources(const cylgpu_ctx* c, Driver& D, int bd, std::vector<double>& s1, std::vector<double>& s2) {
  const int ny = c->g.ny;
  s1.assign((size_t)ny + 1, 0.0);
  s2.assign((size_t)ny + 1, 0.0);
  if (!D.add_laser[bd]) return;
  for (const cylgpu_laser& L : D.lasers) {
    if (L.boundary != bd || !(L.t_start <= D.time && D.time <= L.t_end)) continue;
    double tprof = 1.0;
    if (L.t_width > 0.0) {
      const double a = (D.time - L.t_centre) / L.t_width;
      tprof = std::exp(-(a * a));
    }
    const double t_env = tprof * L.amp;
    const double cpol = std::cos(L.pol_angle), spol = std::sin(L.pol_angle);
    for (int ir = 0; ir <= ny; ++ir) {
      const double y = c->cfg.y_grid_min_local + ((double)ir - 1.0) * c->cfg.dy;
      double prof = 1.0;
      if (L.r_width > 0.0) {
        const double a = (y - 0.0) / L.r_width;
        prof = std::exp(-(a * a));
      }
      const double base = t_env * prof * std::sin(L.omega * D.time + (L.phase + L.phase_curv * (y * y)));
      s1[(size_t)ir] = s1[(size_t)ir] + base * cpol;
      s2[(size_t)ir] = s2[(size_t)ir] + base * spol;
    }
  }
}

static void all_sources(cylgpu_ctx* c, Driver& D) {
  laser_sources(c, D, CYLGPU_BD_X_MIN, D.src[0], D.src[1]);
  laser_sources(c, D, CYLGPU_BD_X_MAX, D.src[2], D.src[3]);
}

// shift_window for one cell, window.F90:62-94
static int shift_window_once(cylgpu_ctx* c, Driver& D) {
  const int ny = c->g.ny, nrow = ny + 2;
  const double x_grid_max = D.x_grid_min + (double)(D.nx_global - 1) * D.dx;
  for (int isp = 0; isp < c->cfg.n_species; ++isp) {   // insert_particles, species in deck order (window.F90:191)
    const cylgpu_insert_profile& P = D.cfg.insert[isp];
    if (!(P.npart_per_cell > 0.0 && P.density > 0.0)) continue;
    D.prof_density.assign((size_t)nrow, P.density);
    D.prof_temp.resize((size_t)3 * nrow);
    D.prof_drift.resize((size_t)3 * nrow);
    for (int k = 0; k < 3; ++k)
      for (int iy = 0; iy < nrow; ++iy) {
        D.prof_temp[(size_t)k * nrow + iy] = P.temp[k];
        D.prof_drift[(size_t)k * nrow + iy] = P.drift[k];
      }
    int64_t n = 0;
    if (D.cfg.insert_mode == 1)
      TRY(cylgpu_insert_particles_device(c, isp, x_grid_max, P.npart_per_cell, D.prof_density.data(), D.prof_temp.data(),
                                         D.prof_drift.data(), P.density_min, P.density_max, D.cfg.insert_seed,
                                         (uint64_t)D.window_shifts_total, &n));
    else
      TRY(cylgpu_insert_particles(c, isp, x_grid_max, P.npart_per_cell, D.prof_density.data(), D.prof_temp.data(),
                                  D.prof_drift.data(), P.density_min, P.density_max, &n));
  }
  // the grid moves by one cell (window.F90:76-86)
  D.x_grid_min = D.x_grid_min + D.dx;
  D.xb_min = D.xb_min + D.dx;
  D.x_min = D.xb_min;
  D.x_max = D.xb_min + (double)D.nx_global * D.dx;
  setup_grid_x(D);
  const double grid5[5] = {D.x_grid_min_local, D.x_min, D.x_max, D.x_min_local, D.x_max_local};
  TRY(cylgpu_window_shift(c, nullptr, nullptr, grid5));
  D.window_shifts_total += 1;
  return 0;
}

// what moving_window is about to decide, with its own arithmetic: the window is already moving and this step
// completes at least one cell.  (The step that STARTS the window changes bc_field first: not predicted.)
static bool window_will_shift(const Driver& D) {
  if (!D.cfg.move_window || !D.window_started) return false;
  if (D.time >= D.cfg.window_stop_time || D.cfg.window_v_x <= 0.0) return false;
  const double frac = D.window_shift_fraction + D.dt * D.cfg.window_v_x / D.dx;
  return (int)std::floor(frac) > 0;
}

// moving_window, window.F90:330-376
static int moving_window(cylgpu_ctx* c, Driver& D) {
  if (!D.cfg.move_window) return 0;
  if (!D.window_started) {
    if (D.cfg.window_start_time <= D.time && D.time < D.cfg.window_stop_time) {
      D.raw_bc_field[CYLGPU_BD_X_MIN] = D.cfg.bc_x_min_after_move;
      D.raw_bc_field[CYLGPU_BD_X_MAX] = D.cfg.bc_x_max_after_move;
      normalise_bc_field(D);
      int32_t bc[4];
      for (int i = 0; i < 4; ++i) bc[i] = D.bc_field[i];
      TRY(cylgpu_set_bc_field(c, bc));
      D.window_shift_fraction = 0.0;
      D.window_started = true;
    }
  }
  if (!D.window_started || D.time >= D.cfg.window_stop_time || D.cfg.window_v_x <= 0.0) return 0;
  D.window_shift_fraction = D.window_shift_fraction + D.dt * D.cfg.window_v_x / D.dx;
  const int cells = (int)std::floor(D.window_shift_fraction);
  if (cells > 0) {
    for (int k = 0; k < cells; ++k) TRY(shift_window_once(c, D));
    TRY(cylgpu_particle_bcs(c));
    D.window_shift_fraction = D.window_shift_fraction - (double)cells;
  }
  return 0;
}

}  // namespace cylgpu

using namespace cylgpu;

extern "C" {

int cylgpu_driver_configure(cylgpu_handle c, const cylgpu_driver_config* cfg) {
  if (!c || !cfg) { set_error("driver_configure: null argument"); return 2; }
  if (cfg->n_lasers < 0 || (cfg->n_lasers > 0 && !cfg->lasers)) { set_error("driver_configure: bad laser list"); return 2; }
  Driver* D = drv(c);
  if (!D) { D = new Driver(); c->driver = D; }
  D->cfg = *cfg;
  D->lasers.assign(cfg->lasers, cfg->lasers + cfg->n_lasers);
  D->cfg.lasers = nullptr;
  D->dx = c->cfg.dx;
  D->dt = c->dt;
  D->nx_global = c->cfg.nx_global;
  D->cell_x_min = cfg->cell_x_min;
  D->cell_x_max = cfg->cell_x_min + c->cfg.nx - 1;
  D->x_grid_min = cfg->x_grid_min;
  D->xb_min = cfg->xb_min;
  D->x_min = c->x_min; D->x_max = c->x_max;
  setup_grid_x(*D);
  for (int i = 0; i < 4; ++i) D->raw_bc_field[i] = cfg->raw_bc_field[i];
  normalise_bc_field(*D);
  D->time = cfg->time;
  D->step = cfg->step;
  D->window_started = cfg->window_started != 0;
  D->window_shift_fraction = cfg->window_shift_fraction;
  D->window_shifts_total = cfg->window_shifts_total;
  return 0;
}

// epoch2d.F90:143-161: particle_bcs, efield_bcs, then bfield_final_bcs over half a step
int cylgpu_driver_init_half_step(cylgpu_handle c) {
  Driver* D = c ? drv(c) : nullptr;
  if (!D) { set_error("driver: cylgpu_driver_configure first"); return 2; }
  TRY(cylgpu_particle_bcs(c));
  TRY(cylgpu_efield_bcs(c));
  const double dt_store = D->dt;
  D->dt = D->dt / 2.0;
  TRY(cylgpu_set_dt(c, D->dt));
  D->time = D->time + D->dt;
  all_sources(c, *D);
  TRY(cylgpu_bfield_final_bcs(c, D->src[0].data(), D->src[1].data(), D->src[2].data(), D->src[3].data()));
  D->dt = dt_store;
  return cylgpu_set_dt(c, dt_store);
}

// the loop body of epoch2d.F90:189-266 with the optional physics packages off, nsteps times
int cylgpu_driver_step(cylgpu_handle c, int64_t nsteps) {
  Driver* D = c ? drv(c) : nullptr;
  if (!D) { set_error("driver: cylgpu_driver_configure first"); return 2; }
  for (int64_t k = 0; k < nsteps; ++k) {
    TRY(cylgpu_fields_half(c));                         // update_eb_fields_half, :213
    TRY(cylgpu_push(c));                                // push_particles, :218
    TRY(cylgpu_current_finish(c));                      // :252
    D->step += 1;
    D->time = D->time + D->dt / 2.0;
    TRY(cylgpu_rng_flush_cache(c));                     // output_routines -> random_flush_cache, diagnostics.F90:235
    D->time = D->time + D->dt / 2.0;
    all_sources(c, *D);
    c->final_shift_follows = window_will_shift(*D);
    const int rc_final = cylgpu_fields_final(c, D->src[0].data(), D->src[1].data(), D->src[2].data(), D->src[3].data());   // :263
    c->final_shift_follows = false;
    TRY(rc_final);
    TRY(moving_window(c, *D));                          // :265
  }
  return 0;
}

int cylgpu_driver_get_state(cylgpu_handle c, cylgpu_driver_state* out) {
  Driver* D = c ? drv(c) : nullptr;
  if (!D || !out) { set_error("driver_get_state: not configured"); return 2; }
  out->time = D->time; out->step = D->step;
  out->window_started = D->window_started ? 1 : 0;
  out->window_shift_fraction = D->window_shift_fraction;
  out->window_shifts_total = D->window_shifts_total;
  out->x_grid_min = D->x_grid_min; out->x_min = D->x_min; out->x_max = D->x_max;
  out->x_grid_min_local = D->x_grid_min_local; out->x_min_local = D->x_min_local; out->x_max_local = D->x_max_local;
  for (int i = 0; i < 4; ++i) { out->bc_field[i] = D->bc_field[i]; out->raw_bc_field[i] = D->raw_bc_field[i]; }
  return 0;
}

int cylgpu_driver_set_time(cylgpu_handle c, double time, int64_t step) {
  Driver* D = c ? drv(c) : nullptr;
  if (!D) { set_error("driver: cylgpu_driver_configure first"); return 2; }
  D->time = time; D->step = step;
  return 0;
}

void cylgpu_driver_release(void* p) { delete static_cast<Driver*>(p); }

}  // extern "C"
