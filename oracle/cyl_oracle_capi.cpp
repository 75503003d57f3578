// cyl_oracle_capi.cpp -- flat C entry points over the CPU oracle so that tests/, smoke()
// and bench.py's cpu_baseline leg can drive it through ctypes.
// TEST INFRASTRUCTURE ONLY (see cyl_oracle.hpp).
#ifdef _OPENMP
#include <omp.h>
#endif
#include <chrono>
#include <cstring>

#include "cyl_oracle.hpp"

using namespace cylo;

extern "C" {

struct CyloConfig {   // must stay layout-identical to cylo::Config
  int32_t nx_global, ny_global, n_mode, nranks;
  double x_min, x_max, y_max;
  double dt_multiplier;
  int32_t bc_field[4];
  int32_t move_window;
  double window_v_x, window_start_time, window_stop_time;
  int32_t bc_x_min_after_move, bc_x_max_after_move;
};

void* cylo_create(const CyloConfig* c) {
  Config cfg;
  cfg.nx_global = c->nx_global; cfg.ny_global = c->ny_global; cfg.n_mode = c->n_mode; cfg.nranks = c->nranks;
  cfg.x_min = c->x_min; cfg.x_max = c->x_max; cfg.y_max = c->y_max;
  cfg.dt_multiplier = c->dt_multiplier;
  for (int i = 0; i < 4; ++i) cfg.bc_field[i] = c->bc_field[i];
  cfg.move_window = c->move_window;
  cfg.window_v_x = c->window_v_x; cfg.window_start_time = c->window_start_time;
  cfg.window_stop_time = c->window_stop_time;
  cfg.bc_x_min_after_move = c->bc_x_min_after_move; cfg.bc_x_max_after_move = c->bc_x_max_after_move;
  return new World(cfg);
}

void cylo_destroy(void* w) { delete (World*)w; }

int cylo_add_species(void* w, double charge, double mass, const int32_t* bc_particle, int immobile,
                     int zero_current, double ppc, double density, const double* temp3, const double* drift3) {
  Species s;
  s.charge = charge; s.mass = mass;
  for (int i = 0; i < 4; ++i) s.bc_particle[i] = bc_particle[i];
  s.immobile = immobile != 0; s.zero_current = zero_current != 0;
  s.npart_per_cell = ppc; s.density = density;
  for (int i = 0; i < 3; ++i) { s.temp[i] = temp3[i]; s.drift[i] = drift3[i]; }
  return ((World*)w)->add_species(s);
}

void cylo_add_laser(void* w, int boundary, double amp, double omega, double pol_angle, double t_start,
                    double t_end, double t_centre, double t_width, double r_width, double phase, double phase_curv) {
  Laser L;
  L.boundary = boundary; L.amp = amp; L.omega = omega; L.pol_angle = pol_angle;
  L.t_start = t_start; L.t_end = t_end; L.t_centre = t_centre; L.t_width = t_width;
  L.r_width = r_width; L.phase = phase; L.phase_curv = phase_curv;
  ((World*)w)->lasers.push_back(L);
}

void cylo_load_uniform(void* w, int isp) { ((World*)w)->load_uniform(isp); }

// out[0..]: dx, dy, dt, time, x_min, x_max, y_max, x_grid_min, xb_min, y_grid_min_local,
//           length_x, window_shift_fraction, step, window_started, window_shifts_total
void cylo_get_scalars(void* wp, double* out) {
  World* w = (World*)wp;
  out[0] = w->dx; out[1] = w->dy; out[2] = w->dt; out[3] = w->time; out[4] = w->x_min; out[5] = w->x_max;
  out[6] = w->y_max; out[7] = w->x_grid_min; out[8] = w->xb_min; out[9] = w->y_grid_min_local;
  out[10] = w->length_x; out[11] = w->window_shift_fraction; out[12] = (double)w->step;
  out[13] = w->window_started ? 1.0 : 0.0; out[14] = (double)w->window_shifts_total;
}
void cylo_set_dt(void* w, double dt) { ((World*)w)->dt = dt; }
void cylo_set_hc_push(void* w, int on) { ((World*)w)->hc_push = on != 0; }
// threads of the `omp parallel for` over ranks (the CPU-baseline timing): set explicitly, because launchers such as
// torchrun export OMP_NUM_THREADS=1 to their children
void cylo_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n > 0 ? n : 1);
#else
  (void)n;
#endif
}
int cylo_get_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void cylo_set_taylor_switch(void* w, double v) { ((World*)w)->taylor_switch = v; }
// calc_number_density_modes into each rank's work array; returns rank k's pointer afterwards via cylo_wk_ptr
void cylo_number_density_modes(void* w, int species) { ((World*)w)->calc_number_density_modes(species); }
void cylo_charge_density(void* w, int species) { ((World*)w)->calc_charge_density(species); }
// counter-based plasma column (cyl_philox.cpp)
void cylo_set_counter_insert(void* w, int on, uint64_t seed) {
  ((World*)w)->counter_insert = on != 0;
  ((World*)w)->counter_seed = seed;
}
// one column with a given column number (tests of the device kernel against a single column)
void cylo_insert_column(void* wp, uint64_t column) {
  World* w = (World*)wp;
  const int64_t keep = w->window_shifts_total;
  w->window_shifts_total = (int64_t)column;
  for (Rank& r : w->ranks) w->insert_particles_counter(r);
  w->window_shifts_total = keep;
}
void cylo_philox4x32(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }
// calc_df.F90 moments (cyl_moments.cpp): result in the real part of mode 0 of rank k's m0
void cylo_moment(void* w, int kind, int species, int direction) { ((World*)w)->calc_moment(kind, species, direction); }
void* cylo_moment_ptr(void* w, int k) { return (void*)((World*)w)->ranks[k].m0.d.data(); }
void* cylo_wk_ptr(void* w, int k) { return (void*)((World*)w)->ranks[k].wk.d.data(); }
void cylo_set_smoothing(void* wp, int enable, int its, int comp_its, int nstrides, const int32_t* strides) {
  World* w = (World*)wp;
  w->smooth_currents = enable != 0;
  w->smooth_its = its;
  w->smooth_comp_its = comp_its;
  w->smooth_strides.assign(strides, strides + nstrides);
}
void cylo_set_time(void* w, double t) { ((World*)w)->time = t; }
void cylo_get_bc_field(void* w, int32_t* out) { for (int i = 0; i < 4; ++i) out[i] = ((World*)w)->bc_field[i]; }
void cylo_get_bc_particle(void* w, int isp, int32_t* out) {
  for (int i = 0; i < 4; ++i) out[i] = ((World*)w)->species[isp].bc_particle[i];
}

// iout: nx, ny, cell_x_min, cell_x_max, x_min_boundary, x_max_boundary
// dout: x_grid_min_local, x_grid_max_local, x_min_local, x_max_local
void cylo_rank_info(void* wp, int k, int32_t* iout, double* dout) {
  Rank& r = ((World*)wp)->ranks[k];
  iout[0] = r.nx; iout[1] = r.ny; iout[2] = r.cell_x_min; iout[3] = r.cell_x_max;
  iout[4] = r.x_min_boundary; iout[5] = r.x_max_boundary;
  dout[0] = r.x_grid_min_local; dout[1] = r.x_grid_max_local; dout[2] = r.x_min_local; dout[3] = r.x_max_local;
}

// field ids 0..14: exm erm etm bxm brm btm jxm jrm jtm bxm_old brm_old btm_old jxm_old jrm_old jtm_old
// ids 15..26: {exm erm etm bxm brm btm}_x_min then _x_max
void* cylo_field_ptr(void* wp, int k, int id) {
  Rank& r = ((World*)wp)->ranks[k];
  Arr3* a3[] = {&r.exm, &r.erm, &r.etm, &r.bxm, &r.brm, &r.btm, &r.jxm, &r.jrm, &r.jtm,
                &r.bxm_old, &r.brm_old, &r.btm_old, &r.jxm_old, &r.jrm_old, &r.jtm_old};
  Arr2* a2[] = {&r.exm_x_min, &r.erm_x_min, &r.etm_x_min, &r.bxm_x_min, &r.brm_x_min, &r.btm_x_min,
                &r.exm_x_max, &r.erm_x_max, &r.etm_x_max, &r.bxm_x_max, &r.brm_x_max, &r.btm_x_max};
  if (id < 15) return a3[id]->d.data();
  return a2[id - 15]->d.data();
}

int64_t cylo_nparticles(void* wp, int k, int isp) { return (int64_t)((World*)wp)->ranks[k].parts[isp].size(); }

void cylo_get_particles(void* wp, int k, int isp, double* out) {
  auto& pl = ((World*)wp)->ranks[k].parts[isp];
  static_assert(sizeof(Particle) == 7 * sizeof(double), "wire format is 7 doubles");
  if (!pl.empty()) std::memcpy(out, pl.data(), pl.size() * sizeof(Particle));
}

void cylo_set_particles(void* wp, int k, int isp, int64_t n, const double* in) {
  auto& pl = ((World*)wp)->ranks[k].parts[isp];
  pl.resize((size_t)n);
  if (n > 0) std::memcpy(pl.data(), in, (size_t)n * sizeof(Particle));
}

// out: n_sent_left, n_sent_right, n_removed, n_recv of the last particle_bcs
void cylo_stats(void* wp, int k, int64_t* out) {
  Rank& r = ((World*)wp)->ranks[k];
  out[0] = r.n_sent_left; out[1] = r.n_sent_right; out[2] = r.n_removed; out[3] = r.n_recv;
}

void cylo_laser_sources(void* wp, int bd, int k, double* s1, double* s2) {
  World* w = (World*)wp;
  std::vector<double> a, b;
  w->laser_sources(bd, w->ranks[k], a, b);
  std::memcpy(s1, a.data(), a.size() * sizeof(double));
  std::memcpy(s2, b.data(), b.size() * sizeof(double));
}

enum Op {
  OP_STEP = 0, OP_FIELDS_HALF = 1, OP_PUSH = 2, OP_CURRENT_FINISH = 3, OP_FIELDS_FINAL = 4,
  OP_MOVING_WINDOW = 5, OP_INIT_HALF_STEP = 6, OP_PARTICLE_BCS = 7, OP_EFIELD_BCS = 8,
  OP_BFIELD_BCS_MPI = 9, OP_BFIELD_FINAL_BCS = 10, OP_UPDATE_E = 11, OP_UPDATE_B = 12,
  OP_SNAPSHOT_BOUNDARIES = 13, OP_ADVANCE_HALF_TIME = 14, OP_PUSH_NO_BCS = 15, OP_CURRENT_BCS = 16,
  OP_FLUSH_RNG = 17, OP_BFIELD_BCS = 18
};

// returns elapsed seconds of the call
void cylo_set_reference_quirks(void* wp, int on) { ((World*)wp)->reference_quirks = on != 0; }

// ghost cells of the arrays of this build (ng = png + 2: the particle shape is compiled in)
int cylo_ng() { return NG; }
int cylo_shape() { return CYLO_SHAPE; }

double cylo_call(void* wp, int op) {
  World* w = (World*)wp;
  auto t0 = std::chrono::steady_clock::now();
  switch (op) {
    case OP_STEP: w->step_once(); break;
    case OP_FIELDS_HALF: w->update_eb_fields_half(); break;
    case OP_PUSH: w->push_particles(); break;
    case OP_CURRENT_FINISH: w->current_finish(); break;
    case OP_FIELDS_FINAL: w->update_eb_fields_final(); break;
    case OP_MOVING_WINDOW: w->moving_window(); break;
    case OP_INIT_HALF_STEP: w->init_half_step(); break;
    case OP_PARTICLE_BCS: w->particle_bcs(); break;
    case OP_EFIELD_BCS: w->efield_bcs(); break;
    case OP_BFIELD_BCS_MPI: w->bfield_bcs(true); break;
    case OP_BFIELD_FINAL_BCS: w->bfield_final_bcs(); break;
    case OP_UPDATE_E: for (Rank& r : w->ranks) w->update_e_field(r); break;
    case OP_UPDATE_B: for (Rank& r : w->ranks) w->update_b_field(r); break;
    case OP_SNAPSHOT_BOUNDARIES: w->snapshot_field_boundaries(); break;
    case OP_ADVANCE_HALF_TIME: w->time = w->time + w->dt / 2.0; break;
    case OP_PUSH_NO_BCS: for (Rank& r : w->ranks) w->push_rank(r); break;
    case OP_CURRENT_BCS: w->current_bcs(); break;
    case OP_BFIELD_BCS: w->bfield_bcs(false); break;
    case OP_FLUSH_RNG: for (Rank& r : w->ranks) r.rng.flush_cache(); break;
    default: return -1.0;
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

double cylo_rng_uniform(void* wp, int k) { return ((World*)wp)->ranks[k].rng.uniform(); }

// the six members of random_state_type (random_generator.f90:26-30) of rank k
void cylo_rng_get_state(void* wp, int k, int32_t* xyzw, int* cached, double* cached_value) {
  const Rng& r = ((World*)wp)->ranks[k].rng;
  xyzw[0] = (int32_t)r.x; xyzw[1] = (int32_t)r.y; xyzw[2] = (int32_t)r.z; xyzw[3] = (int32_t)r.w;
  *cached = r.cached ? 1 : 0;
  *cached_value = r.cached_value;
}

}  // extern "C"
