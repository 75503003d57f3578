// ctx.cuh -- internal state of one cylgpu handle (one x-slab on one B200).
// Product code: never includes, links or calls anything under oracle/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/cylgpu.h"

#include "geom.cuh"
#include "field_ranges.cuh"

namespace cylgpu {

void set_error(const char* fmt, ...);

#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      cylgpu::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,            \
                        cudaGetErrorString(_e));                                         \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

#define TRY(expr)                 \
  do {                            \
    int _r = (expr);              \
    if (_r != 0) return _r;       \
  } while (0)

struct Transport;   // transport.cu

struct SpeciesState {
  cylgpu_species sp;
  bool set = false;
  double* d[7] = {0, 0, 0, 0, 0, 0, 0};   // SoA: x y z px py pz w
  int64_t n = 0, cap = 0;
  double* alt[7] = {0, 0, 0, 0, 0, 0, 0};   // second buffer set: the strip push writes the sorted list here
  int64_t alt_cap = 0;
  int* cell_start = nullptr;   // exclusive scan of the sort buckets of the last sort, ncell + 1 entries
  int64_t cell_start_n = 0;
  uint32_t* perm = nullptr;    // sorted slot -> particle of the last sort (per species: all species are sorted ahead
  int64_t perm_cap = 0;        // of the push on the side stream, presort_fork)
  // Device-resident count (cylgpu_set_exchange_capacity > 0): the exact count then lives on the device
  // (cylgpu_ctx::n_dev[isp]) and `n` above is an UPPER BOUND of it -- enough for grid sizes and capacities --
  // until the next publish has been read back (poll_counts).  lazy = false: `n` is exact.
  bool lazy = false;
};

struct Timer {
  cudaEvent_t a = nullptr, b = nullptr;
};

// random_state_type, random_generator.f90:26-30
struct KissState {
  uint32_t x = 0, y = 0, z = 0, w = 0;
  int cached = 0;
  double cached_value = 0.0;
};

// Device-time accounting without host syncs: each timed span records a pair of events from a
// pool; the elapsed times are read and accumulated when the statistics are asked for (or the
// pool runs low), so that timing never stalls the step.
struct TimerPool {
  struct Span { cudaEvent_t a, b; double* acc; int64_t* count; };
  std::vector<cudaEvent_t> idle;
  std::vector<Span> pending;
  cudaEvent_t get() {
    if (idle.empty()) { cudaEvent_t e = nullptr; cudaEventCreate(&e); return e; }
    cudaEvent_t e = idle.back(); idle.pop_back(); return e;
  }
  void drain() {
    for (Span& s : pending) {
      cudaEventSynchronize(s.b);
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) { *s.acc += (double)ms; if (s.count) *s.count += 1; }
      idle.push_back(s.a); idle.push_back(s.b);
    }
    pending.clear();
  }
  void destroy() {
    drain();
    for (cudaEvent_t e : idle) cudaEventDestroy(e);
    idle.clear();
  }
};

// streams, events and staging of the host-resident particle path (do_push_host)
struct HostStream {
  cudaStream_t up = nullptr, down = nullptr;
  cudaEvent_t ev_up[3] = {0, 0, 0}, ev_unpacked[3] = {0, 0, 0}, ev_packed[2] = {0, 0}, ev_down[2] = {0, 0};
  double* in[3] = {0, 0, 0};    // AoS chunks on their way in
  double* out[2] = {0, 0};      // AoS chunks on their way out
  int64_t cap = 0;              // particles per staging buffer
};

}  // namespace cylgpu

struct cylgpu_ctx {
  cylgpu_config cfg;
  cylgpu::Geom g;
  double dt;
  double x_grid_min_local, x_min, x_max, x_min_local, x_max_local;
  int bc_field[4];
  int left, right;             // neighbour ranks for the Cartesian communicator (-1 = none)
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int device = 0;

  cylgpu::cplx* f[CYLGPU_NFIELDS] = {0};
  cylgpu::cplx* spare = nullptr;          // out-of-place window shift target
  cylgpu::cplx* snap[CYLGPU_NSNAPS] = {0};
  double* tables = nullptr;               // 4 x (ny + 2*JNG + 1) radial tables (particles.F90:190-217)
  int ntab = 0;
  double* src = nullptr;                  // 4 x (ny+1) laser sources (device)
  double* src_stage = nullptr;            // pinned staging: 8 slots of 4 x (ny+1)
  cudaEvent_t src_event[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int src_slot = 0;

  // halo staging: [3 comps][M][SY][NG] complex each
  cylgpu::cplx *sbuf_l = nullptr, *sbuf_r = nullptr, *rbuf_l = nullptr, *rbuf_r = nullptr;
  size_t halo_elems = 0;

  cylgpu::SpeciesState species[CYLGPU_MAX_SPECIES];
  // particle scratch
  double* ptmp = nullptr;        // one SoA component worth of doubles (sort scatter target)
  int64_t ptmp_cap = 0;
  uint32_t* perm = nullptr;      // sort destination per particle
  uint8_t* flag = nullptr;       // fate (left / right / gone) of the h-th leaver
  uint8_t* tailmark = nullptr;   // leavers among the last `nholes` slots
  uint32_t* hole_list = nullptr; // indices of leavers (stand-alone particle_bcs)
  uint32_t* lowhole = nullptr;   // holes below the new count
  uint32_t* hightail = nullptr;  // keepers above the new count
  int64_t pscratch_cap = 0;
  int* scan_blocks = nullptr;
  int64_t ncell = 0;
  double* psend_l = nullptr; double* psend_r = nullptr; double* precv = nullptr;
  int64_t psend_l_cap = 0, psend_r_cap = 0, precv_cap = 0;
  // ---- device-resident particle counts (no host sync inside a step) ----
  // xcap > 0: migrating particles travel in ONE fixed-size message per neighbour, [count header][xcap slots],
  // leaver counts, compaction and arrivals are handled by kernels that read the counts on the device, and the
  // host only keeps upper bounds of the list lengths.  xcap = 0: the exact protocol (counts first, then the
  // payload: two host syncs per species per step).
  int64_t xcap = 0;
  int64_t* n_dev = nullptr;                 // device: [CYLGPU_MAX_SPECIES] exact counts, then [PST_N] statistics
  cylgpu::CompactPlan* d_plan = nullptr;
  struct Publish {
    cudaEvent_t ev = nullptr;
    int64_t bound_at[CYLGPU_MAX_SPECIES];   // the host's upper bounds when the copy was enqueued
    bool pending = false;
  } pub[8];
  int64_t* h_pub = nullptr;                 // pinned: 8 slots of [CYLGPU_MAX_SPECIES + PST_N]
  uint64_t pub_head = 0;
  bool overflowed = false;                  // sticky: a fixed-size exchange met more migrants than xcap
  // window shift with device-resident counts: remove_particles waits for the particle_bcs that follows the shift
  bool pending_remove = false;
  double pending_remove_x = 0.0;
  bool r_clean = false;                     // every listed particle is inside in r (the last particle_bcs saw it there)
  // The cell sort of the next push does not touch the fields and update_eb_fields_half does not touch the
  // particles: with device-resident counts the sort is enqueued on a side stream at the start of
  // cylgpu_fields_half and the push joins it.  With a slab per GPU of an 8-GPU run the field phase is a chain of
  // short launches and neighbour exchanges, i.e. latency, and the sort (bandwidth) hides behind it.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool presorted = false;                   // the side stream holds the sort of the lists as they are now
  // ... and after the push kernel the particle chain of the last species (compaction of the leavers, neighbour
  // exchange, arrivals, count copy) goes to the side stream as well, while the library stream runs current_finish
  // and update_eb_fields_final: two latency chains side by side instead of one after the other
  cudaEvent_t ev_pfork = nullptr, ev_pdone = nullptr;
  bool side_pending = false;                // the library stream has not yet waited for that chain
  int presort_policy = -1;                  // -1: not decided yet, 0: off, 1: on
  // window columns on their way to the device: pinned ring (append_async)
  double* app_pin[4] = {0, 0, 0, 0};
  double* app_dev[4] = {0, 0, 0, 0};
  int64_t app_cap[4] = {0, 0, 0, 0};
  cudaEvent_t app_ev[4] = {0, 0, 0, 0};
  int app_slot = 0;
  unsigned long long* counters = nullptr;   // device counters (8 of them)
  unsigned long long* h_counters = nullptr; // pinned mirror
  double* d_energy = nullptr;

  // current smoothing (current_smooth.F90:49-57,145-196): control-block settings and the two
  // ping-pong work sets of three mode arrays, allocated on first use
  bool smooth_currents = false;
  int smooth_its = 1, smooth_comp_its = 0;
  std::vector<int> smooth_strides;
  cylgpu::cplx* smooth_wk[2][3] = {{0, 0, 0}, {0, 0, 0}};

  // The field phases and current_finish are chains of short launches: replayed as CUDA graphs once their
  // parameters have been stable for a few steps (dt, bc_field, array pointers, stream: `graph_epoch`).
  struct PhaseGraph {
    cudaGraphExec_t exec = nullptr;
    uint64_t epoch = ~0ull;       // epoch the executable was captured for
    uint64_t seen_epoch = ~0ull;  // epoch of the previous call
    int stable_calls = 0;
    int64_t launches = 0;         // kernel launches inside (for the statistics)
    bool failed = false;
  };
  PhaseGraph graphs[4];          // fields_half, fields_final, current_finish, fields_final before a window shift
  bool reference_quirks = true;       // laser.f90's section / REAL-for-imaginary quirks reproduced (cylgpu_set_reference_quirks)
  bool final_shift_follows = false;   // driver.cu: the window moves right after this update_eb_fields_final
  uint64_t graph_epoch = 0;
  bool use_graphs = true;

  void* driver = nullptr;         // driver.cu: the main-loop body run natively (cylgpu_driver_*)
  cylgpu::Transport* tr = nullptr;
  size_t p2p_cap_bytes = 0;
  int p2p_policy = 0;            // CYLGPU_P2P: 0 NCCL only, 1 every message that fits a slot, 2 ("particles") the counted ones
  bool msg_counted = false;      // the message being exchanged is [7-double header: count][slots] (particles)
  bool p2p_link_l = false, p2p_link_r = false;   // peer-memory mailboxes mapped on both ends of the link (transport.cu)
  cylgpu::HostStream hs;
  cylgpu::KissState rng;          // this rank's random stream (window insertion)
  // device-side column (cylgpu_insert_particles_device): profile + row-offset staging
  double* ins_pin = nullptr;
  double* ins_dev = nullptr;
  size_t ins_cap = 0;
  cudaEvent_t ins_ev = nullptr;
  int64_t host_chunk = 1 << 21;   // particles per chunk of the host-resident path (117 MB)
  // remove_particles (window.F90:304-325) for host-resident lists: the window shift only notes the new x_min; the
  // next cylgpu_push_host drops what lies behind it while the chunks stream through (no extra pass over the list)
  bool host_remove_active = false;
  bool sort_drops_behind = false;   // set by do_push_host around the sort of a chunk
  double host_remove_x = 0.0;

  int sort_interval = 1;
  double taylor_switch = 1.0e-4;  // particles.F90:593; moved only by the conditioning test (cylgpu_set_taylor_switch)
  bool hc_push = false;           // Higuera-Cary instead of Boris (the reference's -DHC_PUSH build)
  // 0 per-particle REDs, 1 warp-window shuffle deposit, 2 strip CTAs + shared-memory field patch,
  // 3 strip CTAs + DMMA outer-product deposit
  int push_variant = (CYL_SHAPE == 0) ? 3 : 4;   // 4: the generic per-particle kernel of push_shapes.cuh
  int64_t pushes_since_sort = 0;
  bool sorted_valid = false;

  cylgpu_stats_t stats;
  cylgpu::TimerPool timers;
  bool timing = true;
  bool blocking_wait = false;     // host syncs yield the core instead of spinning (multi-rank hosts)
  cudaEvent_t ev_wait = nullptr;
};

namespace cylgpu {

// fields.cu
int launch_update_e(cylgpu_ctx* c);
int launch_update_b(cylgpu_ctx* c, bool save_old = false);
int launch_update_e(cylgpu_ctx* c, int ix_lo, int ix_hi);
int launch_update_b(cylgpu_ctx* c, bool save_old, int ix_lo, int ix_hi);
// bcs.cu
int do_efield_bcs(cylgpu_ctx* c);
int do_bfield_bcs(cylgpu_ctx* c, bool mpi_only);
int do_bfield_final_bcs(cylgpu_ctx* c, const double* s1min, const double* s2min, const double* s1max,
                        const double* s2max);
int upload_laser_sources(cylgpu_ctx* c, const double* s1min, const double* s2min, const double* s1max,
                         const double* s2max);
int do_bfield_final_bcs_device(cylgpu_ctx* c, bool wide = false, int phase = 1);
int efield_edges(cylgpu_ctx* c);
int bfield_edges(cylgpu_ctx* c);
int halo_eb(cylgpu_ctx* c);
bool wide_fields(const cylgpu_ctx* c);
FieldRanges wide_ranges(const cylgpu_ctx* c, int phase);
int do_current_bcs(cylgpu_ctx* c);
int do_current_finish(cylgpu_ctx* c);
int do_number_density_modes(cylgpu_ctx* c, int species, bool charge);
int do_particle_moment(cylgpu_ctx* c, int kind, int species, int direction, double* host_out);
int download_real_part_mode0(cylgpu_ctx* c, const cplx* a, double* host_out);
int do_r_min_final(cylgpu_ctx* c);
int do_snapshot(cylgpu_ctx* c);
int do_shift_fields(cylgpu_ctx* c);
int halo_x(cylgpu_ctx* c, int f0, int f1, int f2, int skip0, int skip1, int skip2);
// particles.cu
int do_push(cylgpu_ctx* c);
int do_push_bcs(cylgpu_ctx* c);
int presort_fork(cylgpu_ctx* c);                 // enqueue the next push's cell sort on the side stream
int presort_join(cylgpu_ctx* c, bool still_valid);   // main stream waits for it; !still_valid: the lists change, forget it
int flush_pending_remove(cylgpu_ctx* c);         // remove_particles left pending by a window shift, now
int poll_counts(cylgpu_ctx* c, bool block);      // tighten (block: make exact) the host's particle counts
int publish_counts(cylgpu_ctx* c);               // enqueue the device -> host copy of counts + statistics
int set_count_exact(cylgpu_ctx* c, int isp);     // the host changed species[isp].n: mirror it on the device
int append_async(cylgpu_ctx* c, int isp, int64_t n, const double* host_aos);
int do_particle_bcs(cylgpu_ctx* c);
int do_push_host(cylgpu_ctx* c, const int64_t* n_in, double* const* host_aos, const int64_t* capacity,
                 int64_t* n_out);
int do_sort(cylgpu_ctx* c);
int do_remove_behind(cylgpu_ctx* c);
int do_energy(cylgpu_ctx* c, double* out2);
int do_cells(cylgpu_ctx* c, int isp, int64_t capn, int32_t* out);
int reserve_particles(cylgpu_ctx* c, int isp, int64_t n);
int build_tables(cylgpu_ctx* c);
// window_insert.cu
void kiss_init(KissState& s, int seed);
double kiss_uniform(KissState& s);
double kiss_box_muller(KissState& s, double stdev, double mu);
int do_insert_particles(cylgpu_ctx* c, int isp, double x_grid_max, double npart_per_cell, const double* density,
                        const double* temperature, const double* drift, double dmin, double dmax,
                        std::vector<double>& aos);
int do_insert_particles_device(cylgpu_ctx* c, int isp, double x_grid_max, double npart_per_cell, const double* density,
                               const double* temperature, const double* drift, double dmin, double dmax, uint64_t seed,
                               uint64_t column, int64_t* n_inserted);
// sdf_io.cu
int sdf_write_host(const char* path, const cylgpu_sdf_desc* d, const void* const* fields15,
                   const double* const* particles_aos, const double* const* derived);
int sdf_derived_count(const cylgpu_sdf_desc* d);
int sdf_read_host(const char* path, cylgpu_sdf_desc* d, void* const* fields15, double x_lo, double x_hi,
                  std::vector<std::vector<double>>* particles);
// transport.cu
Transport* make_transport(cylgpu_ctx* c);
int p2p_setup(cylgpu_ctx* c, size_t cap_bytes);   // collective over the neighbours; failure leaves NCCL in charge
int p2p_check(cylgpu_ctx* c);                     // error if an exchange through the mailboxes timed out
bool transport_two_streams(const cylgpu_ctx* c);  // exchanges may run on the side stream beside the library stream's
void destroy_transport(Transport* t);
int transport_sendrecv(cylgpu_ctx* c, const void* sl, size_t sl_b, void* rl, size_t rl_b, const void* sr,
                       size_t sr_b, void* rr, size_t rr_b);

struct PhaseTimer {   // accumulates the device ms of a span into a stats field (read at cylgpu_stats)
  cylgpu_ctx* c;
  double* acc;
  int64_t* count;
  cudaEvent_t a = nullptr;
  PhaseTimer(cylgpu_ctx* c_, double* acc_, int64_t* count_ = nullptr, bool enabled = true)
      : c(c_), acc(acc_), count(count_) {
    if (c->timing && enabled) { a = c->timers.get(); cudaEventRecord(a, c->stream); }
  }
  void stop() {
    if (!a) return;
    cudaEvent_t b = c->timers.get();
    cudaEventRecord(b, c->stream);
    c->timers.pending.push_back({a, b, acc, count});
    a = nullptr;
    if (c->timers.pending.size() >= 4096) c->timers.drain();
  }
  ~PhaseTimer() { stop(); }
};

}  // namespace cylgpu
