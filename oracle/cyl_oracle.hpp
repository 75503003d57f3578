// cyl_oracle.hpp -- CPU oracle for the cylindrical-EPOCH per-timestep PIC hot path.
//
// TEST INFRASTRUCTURE ONLY.  This is a from-scratch FP64 restatement, in C++, of the
// algorithm the reference implements in Fortran (reference files are cited per function
// as `file:line`, relative to /root/reference/epoch_axial/src).  Nothing in the product
// library (cylindrical_epoch_b200/csrc) includes, links or calls this code; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// PARITY UNPINNED: the reference ships no golden vectors, fixtures or runnable tests for
// the cylindrical path, and neither this container nor the GPU box
// has a Fortran compiler or MPI, so the real reference cannot be run to produce any.  The
// oracle is therefore anchored only by (i) the analytic expectation of
// example_decks/current_density_test.deck (DOCUMENTATION.pdf section 7.2), (ii) exact
// discrete charge continuity of the deposit, (iii) the axis-condition identities and
// (iv) vacuum-propagation sanity checks, (v) the reference's gaussian_pulse example deck reaching the focus its
// own constants block designs, (vi) Boris rotation, plasma oscillation and Gauss-law invariants -- see
// tests/test_oracle*.py.  Only the SDF container (cylindrical_epoch_b200/csrc/sdf_io.cu) is checked by
// reference code proper: the reference's SDF C reader, compiled into oracle/_ref by oracle/sdf_ref/Makefile.
// tools/check_against_reference_dumps.py pins this oracle against two restart dumps of the real reference
// for anyone who can build it.
//
// Conventions (all mirror the reference so arrays can be compared element by element):
//   * mode arrays are Fortran column-major (ix, ir, im), lower bounds (1-ng, 1-ng, 0),
//     ng = 5, complex128 interleaved;
//   * "x" is the cylinder axis, grid "y" is r, particle pos/p are Cartesian (x, y, z);
//   * decomposition is x-slabs only (nprocx = nranks, nprocy = 1): every rank owns the
//     axis and r_max (y_min_boundary = y_max_boundary = true).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace cylo {

// The particle shape is a compile-time choice of the reference (-DPARTICLE_SHAPE_TOPHAT / _BSPLINE3, else the
// triangle; constants.F90:524-545) and so it is here: -DCYLO_SHAPE=0 triangle (default), 1 top-hat, 2 third-order
// B-spline.  It sets the ghost widths -- i.e. the layout of every array -- besides the weights.
#ifndef CYLO_SHAPE
#define CYLO_SHAPE 0
#endif
#if CYLO_SHAPE == 2
constexpr int SF_MIN = -2, SF_MAX = 2, PNG = 4;   // constants.F90:526-529
#elif CYLO_SHAPE == 1
constexpr int SF_MIN = 0, SF_MAX = 1, PNG = 2;    // constants.F90:530-533
#else
constexpr int SF_MIN = -1, SF_MAX = 1, PNG = 3;   // constants.F90:534-537
#endif
constexpr int NG = PNG + 2;   // constants.F90:544
constexpr int JNG = NG;       // constants.F90:545  MAX(ng, png)
constexpr int WO = 3;         // weight arrays hold offsets -3..3 (sf_min-1 : sf_max+1 of the widest shape): w[k + WO]
constexpr int NW = 7;

inline double pow4(double x) { const double t = x * x; return t * t; }   // x**4 by repeated squaring

// <shape>/gx.inc and hx_dcell.inc: the UNNORMALISED weights of one direction, placed at shift+sf_min .. shift+sf_max
// (shift = 0 for gx / gy, dcell for hx / hy).  w[] must be zeroed by the caller where the reference zeroes it.
inline void shape_weights(double cf, int shift, double* w) {
#if CYLO_SHAPE == 2
  const double cf2 = cf * cf;
  w[shift - 2 + WO] = pow4(0.5 + cf);
  w[shift - 1 + WO] = 4.75 + 11.0 * cf + 4.0 * cf2 * (1.5 - cf - cf2);
  w[shift + WO] = 14.375 + 6.0 * cf2 * (cf2 - 2.5);
  w[shift + 1 + WO] = 4.75 - 11.0 * cf + 4.0 * cf2 * (1.5 + cf - cf2);
  w[shift + 2 + WO] = pow4(0.5 - cf);
#elif CYLO_SHAPE == 1
  w[shift + WO] = 0.5 + cf;
  w[shift + 1 + WO] = 0.5 - cf;
#else
  const double cf2 = cf * cf;
  w[shift - 1 + WO] = 0.25 + cf2 + cf;
  w[shift + WO] = 1.5 - 2.0 * cf2;
  w[shift + 1 + WO] = 0.25 + cf2 - cf;
#endif
}
// the factor the weights above still need per direction pair (particles.F90:145-153)
#if CYLO_SHAPE == 2
constexpr double SHAPE_FAC = (1.0 / 24.0) * (1.0 / 24.0);
#elif CYLO_SHAPE == 1
constexpr double SHAPE_FAC = 1.0;
#else
constexpr double SHAPE_FAC = 0.25;
#endif
// top-hat: positions are measured from the cell edge (particles.F90:336-342, 540-546)
constexpr double SHAPE_CELL_SHIFT = (CYLO_SHAPE == 1) ? 0.5 : 0.0;

// <shape>/gxfac.inc without its fold at the axis: the NORMALISED weights of particle_to_grid.inc, w[k + WO]
inline void shape_weights_fac(double cf, double* w) {
#if CYLO_SHAPE == 2
  const double third = 1.0 / 3.0;
  const double fac1 = 0.125 * third, fac2 = 0.5 * third, fac3 = 7.1875 * third;   // particle_head.inc
  const double c2 = cf * cf;
  w[-2 + WO] = fac1 * pow4(0.5 + cf);
  w[-1 + WO] = fac2 * (1.1875 + 2.75 * cf + c2 * (1.5 - cf - c2));
  w[0 + WO] = 0.25 * (fac3 + c2 * (c2 - 2.5));
  w[1 + WO] = fac2 * (1.1875 - 2.75 * cf + c2 * (1.5 + cf - c2));
  w[2 + WO] = fac1 * pow4(0.5 - cf);
#elif CYLO_SHAPE == 1
  w[0 + WO] = 0.5 + cf;
  w[1 + WO] = 0.5 - cf;
#else
  const double c2 = cf * cf;
  w[-1 + WO] = 0.5 * (0.25 + c2 + cf);
  w[0 + WO] = 0.75 - c2;
  w[1 + WO] = 0.5 * (0.25 + c2 - cf);
#endif
}
// the fold of the radial weights of gxfac.inc for a particle next to the axis
inline void shape_axis_fold(double part_r, double dy, double* gy) {
#if CYLO_SHAPE == 2
  if (part_r < 2.0 * dy) {
    if (part_r < dy) {
      gy[0 + WO] = gy[0 + WO] + gy[-1 + WO];
      gy[1 + WO] = gy[1 + WO] + gy[-2 + WO];
      gy[-1 + WO] = 0.0;
      gy[-2 + WO] = 0.0;
    } else {
      gy[-1 + WO] = gy[-1 + WO] + gy[-2 + WO];
      gy[-2 + WO] = 0.0;
    }
  }
#elif CYLO_SHAPE == 1
  if (part_r < 0.5 * dy) {
    gy[1 + WO] = 1.0;
    gy[0 + WO] = 0.0;
  }
#else
  if (part_r < dy) {
    gy[0 + WO] = gy[0 + WO] + gy[-1 + WO];
    gy[-1 + WO] = 0.0;
  }
#endif
}
// include/particle_to_grid.inc: nearest cell, fractions and normalised weights of a particle at (x, r)
inline void particle_to_grid(double x_local, double r_local, double part_r, double dx, double dy, int* cell_x,
                             int* cell_y, double* gx, double* gy) {
  const double cell_x_r = x_local / dx - SHAPE_CELL_SHIFT;
  const double cell_y_r = r_local / dy - SHAPE_CELL_SHIFT;
  int cx = (int)std::floor(cell_x_r + 0.5);
  int cy = (int)std::floor(cell_y_r + 0.5);
  const double cfx = (double)cx - cell_x_r;
  const double cfy = (double)cy - cell_y_r;
  *cell_x = cx + 1;
  *cell_y = cy + 1;
  for (int k = 0; k < NW; ++k) { gx[k] = 0.0; gy[k] = 0.0; }
  shape_weights_fac(cfx, gx);
  shape_weights_fac(cfy, gy);
  shape_axis_fold(part_r, dy, gy);
}

// Boundary-condition codes, constants.F90:55-72
enum BC {
  BC_PERIODIC = 1, BC_OTHER = 2, BC_SIMPLE_LASER = 3, BC_SIMPLE_OUTFLOW = 4, BC_OPEN = 5,
  BC_ZERO_GRADIENT = 7, BC_CLAMP = 8, BC_REFLECT = 9, BC_CONDUCT = 10, BC_THERMAL = 11,
  BC_CPML_LASER = 12, BC_CPML_OUTFLOW = 13, BC_MIXED = 14, BC_ZERO_B = 16
};
// boundary location codes (0-based here; reference c_bd_x_min..c_bd_y_max = 1..4)
enum BD { BD_X_MIN = 0, BD_X_MAX = 1, BD_Y_MIN = 2, BD_Y_MAX = 3 };

// physical constants, constants.F90:171-201
constexpr double PI = 3.141592653589793238462643383279503;
constexpr double Q0 = 1.602176565e-19;
constexpr double M0 = 9.10938291e-31;
constexpr double C_LIGHT = 2.99792458e8;
constexpr double KB = 1.3806488e-23;
constexpr double EPSILON0 = 8.854187817620389850536563031710750e-12;

// Minimal complex type with the naive (Fortran-rule) products; no FMA contraction is
// allowed when this file is compiled (-ffp-contract=off), matching a generic x86-64
// gfortran build of the reference.
struct cplx {
  double re, im;
  cplx() : re(0.0), im(0.0) {}
  cplx(double r) : re(r), im(0.0) {}
  cplx(double r, double i) : re(r), im(i) {}
};
inline cplx operator+(cplx a, cplx b) { return cplx(a.re + b.re, a.im + b.im); }
inline cplx operator-(cplx a, cplx b) { return cplx(a.re - b.re, a.im - b.im); }
inline cplx operator-(cplx a) { return cplx(-a.re, -a.im); }
inline cplx operator*(cplx a, cplx b) {
  return cplx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
inline cplx operator*(double s, cplx a) { return cplx(s * a.re, s * a.im); }
inline cplx operator*(cplx a, double s) { return cplx(a.re * s, a.im * s); }
inline cplx operator/(cplx a, double s) { return cplx(a.re / s, a.im / s); }
inline cplx& operator+=(cplx& a, cplx b) { a.re += b.re; a.im += b.im; return a; }
// real / complex, Smith's method as used by libgcc for COMPLEX division
inline cplx recip(cplx z) {
  if (std::abs(z.re) >= std::abs(z.im)) {
    double r = z.im / z.re, den = z.re + z.im * r;
    return cplx(1.0 / den, -r / den);
  } else {
    double r = z.re / z.im, den = z.im + z.re * r;
    return cplx(r / den, -1.0 / den);
  }
}
const cplx IMAGI(0.0, 1.0);

// (ix, ir, im) array with Fortran bounds (1-ng:nx+ng, 1-ng:ny+ng, 0:M-1)
struct Arr3 {
  int nx = 0, ny = 0, M = 0;
  std::vector<cplx> d;
  void alloc(int nx_, int ny_, int M_) {
    nx = nx_; ny = ny_; M = M_;
    d.assign((size_t)(nx + 2 * NG) * (ny + 2 * NG) * M, cplx());
  }
  inline cplx& operator()(int ix, int ir, int im) {
    return d[((size_t)im * (ny + 2 * NG) + (size_t)(ir + NG - 1)) * (nx + 2 * NG) + (ix + NG - 1)];
  }
  inline const cplx& operator()(int ix, int ir, int im) const {
    return d[((size_t)im * (ny + 2 * NG) + (size_t)(ir + NG - 1)) * (nx + 2 * NG) + (ix + NG - 1)];
  }
  void zero() { std::fill(d.begin(), d.end(), cplx()); }
};
// (ir, im) boundary snapshot, bounds (1-ng:ny+ng, 0:M-1), setup.F90:397-402
struct Arr2 {
  int ny = 0, M = 0;
  std::vector<cplx> d;
  void alloc(int ny_, int M_) { ny = ny_; M = M_; d.assign((size_t)(ny + 2 * NG) * M, cplx()); }
  inline cplx& operator()(int ir, int im) { return d[(size_t)im * (ny + 2 * NG) + (ir + NG - 1)]; }
  inline const cplx& operator()(int ir, int im) const { return d[(size_t)im * (ny + 2 * NG) + (ir + NG - 1)]; }
};

struct Particle {  // shared_data.F90:93-142 (default build: pos, p, weight)
  double pos[3], p[3], w;
};

struct Species {   // shared_data.F90:190-280 (hot-path members only)
  double charge, mass;
  int bc_particle[4];
  bool immobile, zero_current;
  double npart_per_cell;                    // for window insertion
  double density, temp[3], drift[3];        // uniform initial conditions
};

// KISS generator + polar Box-Muller, random_generator.f90:45-173
struct Rng {
  uint32_t x, y, z, w;
  bool cached;
  double cached_value;
  void init(int seed);
  double uniform();
  double box_muller(double stdev, double mu);
  void flush_cache() { cached = false; }
};

struct Laser {  // laser.f90 laser_block, restricted to what the decks in scope use
  int boundary;            // BD_X_MIN or BD_X_MAX
  double amp, omega, pol_angle, t_start, t_end;
  double t_centre, t_width;   // t_profile = gauss(time, t_centre, t_width); t_width<=0 -> 1
  double r_width;             // profile = gauss(y, 0, r_width); r_width<=0 -> 1
  double phase;               // constant phase
  double phase_curv = 0.0;    // phase(y) = phase + phase_curv * y^2 (the deck's phase function, laser.f90:203-226,454)
};

// the real-valued particle moments of io/calc_df.F90 (cyl_moments.cpp)
enum Moment {
  MOM_MASS_DENSITY = 0, MOM_NUMBER_DENSITY = 1, MOM_EKBAR = 2, MOM_EKFLUX = 3, MOM_PPC = 4,
  MOM_AVERAGE_WEIGHT = 5, MOM_TEMPERATURE = 6, MOM_SPECIES_CURRENT = 7, MOM_AVERAGE_MOMENTUM = 8
};

struct Rank {
  int nx, ny, M;
  int x_coord, nprocx;
  int cell_x_min, cell_x_max;   // global cell range, mpi_routines.F90:312-337
  bool x_min_boundary, x_max_boundary;
  double x_grid_min_local, x_grid_max_local, x_min_local, x_max_local;
  Arr3 exm, erm, etm, bxm, brm, btm, jxm, jrm, jtm;
  Arr3 bxm_old, brm_old, btm_old, jxm_old, jrm_old, jtm_old;
  Arr3 wk;   // work array of smooth_mode_array (current_smooth.F90:145-196)
  Arr3 m0, m1, m2, m3, m4;   // work arrays of calc_moment (cyl_moments.cpp); real data in mode 0
  Arr2 exm_x_min, erm_x_min, etm_x_min, bxm_x_min, brm_x_min, btm_x_min;
  Arr2 exm_x_max, erm_x_max, etm_x_max, bxm_x_max, brm_x_max, btm_x_max;
  std::vector<std::vector<Particle>> parts;   // per species, in linked-list order
  Rng rng;
  // statistics of the last particle_bcs call (per species summed)
  int64_t n_sent_left = 0, n_sent_right = 0, n_removed = 0, n_recv = 0;
};

struct Config {
  int nx_global, ny_global, n_mode, nranks;
  double x_min, x_max, y_max;      // y_min = 0 always (deck_control_block.F90:74-75)
  double dt_multiplier;            // setup.F90:78 default 0.95
  int bc_field[4];                 // raw deck codes; normalised by setup_boundaries
  // moving window (window.F90:330-376)
  int move_window;
  double window_v_x, window_start_time, window_stop_time;
  int bc_x_min_after_move, bc_x_max_after_move;
};

void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);   // cyl_philox.cpp

struct World {
  Config cfg;
  int M;
  double dx, dy, dt, time;
  int step;
  double x_min, x_max, y_max, length_x;
  double x_grid_min, xb_min;       // global grid origin (cell centre / cell edge)
  double y_grid_min_local;
  int bc_field[4];
  bool add_laser[4];
  std::vector<Species> species;
  std::vector<Laser> lasers;
  std::vector<Rank> ranks;
  // window state
  bool window_started = false;
  double window_shift_fraction = 0.0;
  int64_t window_shifts_total = 0;

  explicit World(const Config& c);
  int add_species(const Species& s);
  void setup_boundaries();                       // boundary.F90:30-75
  void setup_grid_x();                           // utilities.f90:343-372
  void load_uniform(int ispecies);               // helper.F90:373-811, particle_temperature.F90:30-83
  void snapshot_field_boundaries();              // setup.F90:393-423
  void init_half_step();                         // epoch2d.F90:143-161

  // hot path, World-level (all ranks, with the MPI exchanges emulated in-process)
  void update_e_field(Rank& r);                  // fields.f90:53-182
  void update_b_field(Rank& r);                  // fields.f90:186-312
  void efield_bcs();                             // boundary.F90:1355-1413
  void bfield_bcs(bool mpi_only);                // boundary.F90:1417-1476
  void bfield_final_bcs();                       // boundary.F90:1505-1537
  void update_eb_fields_half();                  // fields.f90:316-337
  void update_eb_fields_final();                 // fields.f90:341-353
  void push_particles();                         // particles.F90:28-734
  void push_rank(Rank& r);                       //   the per-rank particle loop
  void current_bcs_r_min_final(Rank& r);         // boundary.F90:1909-1959
  void particle_bcs();                           // boundary.F90:1541-1889
  void current_bcs();                            // boundary.F90:1893-1905
  void current_finish();                         // current_smooth.F90:29-45
  void smooth_mode_array(Arr3 Rank::*f);         // current_smooth.F90:145-196
  void calc_number_density_modes(int species);   // calc_df.F90:588-661 -> Rank::wk (species < 0: all)
  void calc_charge_density(int species);         // calc_df.F90:442-519 -> Rank::wk, mode 0, real part
  void density_deposit_and_bcs(int species, bool charge);
  void calc_moment(int kind, int species, int direction);   // calc_df.F90:59-1221 -> Rank::m0 (cyl_moments.cpp)
  void moment_summation_bcs(Arr3 Rank::*f);      // calc_boundary, calc_df.F90:24-31
  void centre_zero_gradient(Arr3 Rank::*f);      // boundary.F90:597-707, c_stagger_centre
  bool smooth_currents = false;                  // shared_data.F90:468-472
  double taylor_switch = 1.0e-4;                 // |m dtheta| below which particles.F90:593-598 use the series (test knob)
  bool hc_push = false;                          // -DHC_PUSH, particles.F90:409-421
  // laser.f90: r_d_vals(0:ny) / source_t(0:ny) used whole against (1:ny) sections (:474,:587,:604) and icdt_2r
  // declared REAL but assigned an imaginary value (:640,:648).  true: what gfortran makes of that (the default, the
  // reference's behaviour); false: element for element and the imaginary coefficient (what the code spells)
  bool reference_quirks = true;
  int smooth_its = 1, smooth_comp_its = 0;
  std::vector<int> smooth_strides;
  void moving_window();                          // window.F90:330-376
  void step_once();                              // epoch2d.F90:189-266 loop body

  // pieces
  void halo_x(Arr3 Rank::*f, int row_lo_off, int row_hi_off);      // boundary.F90:500-553
  void clamp_zero(Rank& r, Arr3& f, bool stag_x, bool stag_y, int bd);     // boundary.F90:772-829
  void zero_gradient(Rank& r, Arr3& f, bool stag_x, bool stag_y, int bd);  // boundary.F90:654-707
  void laser_sources(int bd, const Rank& r, std::vector<double>& s1, std::vector<double>& s2);
  void outflow_bcs_x_min(Rank& r);               // laser.f90:411-520
  void outflow_bcs_x_max(Rank& r);               // laser.f90:524-633
  void outflow_bcs_r_max(Rank& r);               // laser.f90:637-690
  void reflection_bcs(Rank& r, Arr3& a, int im, int flip_dir);     // boundary.F90:918-1015
  void periodic_sum_x(Arr3 Rank::*f);            // boundary.F90:1133-1203
  void insert_particles(Rank& r);                // window.F90:157-300
  void insert_particles_counter(Rank& r);        // the same with Philox counters (cyl_philox.cpp)
  bool counter_insert = false;                   // moving_window uses insert_particles_counter
  uint64_t counter_seed = 0;
  void shift_fields();                           // window.F90:98-153
  int bc_allspecies(int bd) const;
};

}  // namespace cylo
