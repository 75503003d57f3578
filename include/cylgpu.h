/* cylgpu.h -- C-ABI of the B200-native per-timestep PIC hot path of cylindrical EPOCH.
 *
 * One `cylgpu_handle` == one MPI rank of the reference == one x-slab == one GPU.
 * The reference has no plugin API: the seam is the set of argument-less Fortran module
 * procedures that operate on `shared_data` globals.  Every entry point below names the
 * reference procedure (file:line under /root/reference/epoch_axial/src) it replaces; the
 * iso_c_binding stub a maintainer adds on the Fortran side is in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure (Fortran then calls
 *     abort_code, utilities.f90:261); cylgpu_last_error() gives the message;
 *   - mode arrays are Fortran column-major (ix, ir, im) with lower bounds
 *     (1-ng, 1-ng, 0), ng = 5, COMPLEX(num) == interleaved re/im doubles
 *     (shared_data.F90:474-478, mpi_routines.F90:385-400); the address passed is that of
 *     element (1-ng, 1-ng, 0);
 *   - "x" is the cylinder axis, grid "y" is r; particle pos/p are Cartesian (x, y, z);
 *   - decomposition is x-slabs only (nprocx = nranks, nprocy = 1).
 *   - no torch types, no C++ types: plain pointers and sizes only.
 */
#ifndef CYLGPU_H
#define CYLGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CYLGPU_NG 5            /* constants.F90:544-545 (ng = jng = png + 2, triangle shape) */
#define CYLGPU_NFIELDS 15      /* exm erm etm bxm brm btm jxm jrm jtm b*_old j*_old */
#define CYLGPU_NSNAPS 12       /* {exm erm etm bxm brm btm}_x_min then _x_max, setup.F90:393-423 */
#define CYLGPU_MAX_SPECIES 8

/* boundary-condition codes, constants.F90:55-72 (values after setup_boundaries()
 * normalisation, boundary.F90:44-57,109-123) */
enum {
  CYLGPU_BC_PERIODIC = 1, CYLGPU_BC_SIMPLE_LASER = 3, CYLGPU_BC_SIMPLE_OUTFLOW = 4,
  CYLGPU_BC_OPEN = 5, CYLGPU_BC_ZERO_GRADIENT = 7, CYLGPU_BC_CLAMP = 8, CYLGPU_BC_REFLECT = 9,
  CYLGPU_BC_CONDUCT = 10, CYLGPU_BC_CPML_LASER = 12, CYLGPU_BC_CPML_OUTFLOW = 13,
  CYLGPU_BC_ZERO_B = 16
};
/* boundary ids (0-based; reference c_bd_x_min..c_bd_y_max = 1..4) */
enum { CYLGPU_BD_X_MIN = 0, CYLGPU_BD_X_MAX = 1, CYLGPU_BD_Y_MIN = 2, CYLGPU_BD_Y_MAX = 3 };
/* field ids for upload/download/device_ptr */
enum {
  CYLGPU_EXM = 0, CYLGPU_ERM, CYLGPU_ETM, CYLGPU_BXM, CYLGPU_BRM, CYLGPU_BTM,
  CYLGPU_JXM, CYLGPU_JRM, CYLGPU_JTM, CYLGPU_BXM_OLD, CYLGPU_BRM_OLD, CYLGPU_BTM_OLD,
  CYLGPU_JXM_OLD, CYLGPU_JRM_OLD, CYLGPU_JTM_OLD
};
/* transports for the x-neighbour exchange (MPI_SENDRECV in the reference) */
enum {
  CYLGPU_TRANSPORT_NONE = 0,     /* nranks == 1 (periodic wraps onto itself on the device) */
  CYLGPU_TRANSPORT_NCCL = 1,     /* ncclSend/ncclRecv over NVLink; libnccl is dlopen()ed */
  CYLGPU_TRANSPORT_CALLBACK = 2, /* caller-supplied sendrecv (e.g. torch.distributed) */
  CYLGPU_TRANSPORT_FABRIC = 3    /* several handles inside ONE process (tests, 1 GPU) */
};

typedef struct cylgpu_ctx* cylgpu_handle;

/* Caller-supplied exchange used by CYLGPU_TRANSPORT_CALLBACK.  All four buffers are DEVICE
 * pointers; byte counts may be 0; rank -1 means "no neighbour" (MPI_PROC_NULL).  It must
 * behave like the pair of MPI_SENDRECVs at boundary.F90:528,541: send `send_left` to
 * `left`, `send_right` to `right`, receive `recv_left` from `left`, `recv_right` from
 * `right`; the work may be left in flight on `stream` (a cudaStream_t). */
typedef int (*cylgpu_sendrecv_fn)(void* user, int left, int right,
                                  const void* send_left, size_t send_left_bytes,
                                  void* recv_left, size_t recv_left_bytes,
                                  const void* send_right, size_t send_right_bytes,
                                  void* recv_right, size_t recv_right_bytes, void* stream);

/* Everything the hot path reads from shared_data (shared_data.F90:473-488,656-667) that is
 * not an array.  Filled by the Fortran shim after mpi_initialise / setup_grid / set_dt. */
typedef struct cylgpu_config {
  int32_t nx, ny;                /* local cells of this rank (mpi_routines.F90:312-337) */
  int32_t nx_global, ny_global;
  int32_t n_mode;                /* deck_control_block.F90:244 */
  int32_t rank, nranks;          /* x_coords, nprocx (nprocy must be 1) */
  int32_t x_min_boundary, x_max_boundary;   /* shared_data.F90:660 */
  int32_t bc_field[4];           /* after setup_boundaries */
  int32_t n_species;
  int32_t device;                /* CUDA device ordinal; -1 = current device */
  int32_t transport;             /* CYLGPU_TRANSPORT_* */
  double dx, dy, dt;
  double x_grid_min_local;       /* x(1) of this rank, utilities.f90:343-372 */
  double y_grid_min_local;       /* y(1) = dy/2 on the axis rank, setup.F90:173-183 */
  double x_min, x_max, y_max;    /* global physical domain (moves with the window) */
  double x_min_local, x_max_local;
  /* transport parameters */
  const void* nccl_unique_id;    /* 128 bytes from ncclGetUniqueId (NCCL transport) */
  cylgpu_sendrecv_fn sendrecv;   /* CALLBACK transport */
  void* sendrecv_user;
  void* fabric;                  /* FABRIC transport: from cylgpu_fabric_create */
  int64_t particle_capacity;     /* per species; 0 = grow on demand */
} cylgpu_config;

/* shared_data.F90:190-280, hot-path members only */
typedef struct cylgpu_species {
  double charge, mass;
  int32_t bc_particle[4];        /* after setup_boundaries: periodic / open / reflect */
  int32_t immobile, zero_current;
} cylgpu_species;

/* cylgpu_stats: integer outputs that must match the reference bit-exactly */
typedef struct cylgpu_stats_t {
  int64_t n_particles[CYLGPU_MAX_SPECIES];   /* species_list(:)%attached_list%count */
  int64_t n_sent_left, n_sent_right;         /* last particle_bcs, all species */
  int64_t n_removed, n_recv;
  int64_t n_window_removed;                  /* last window shift, remove_particles */
  int64_t n_sorts;                           /* cell-tile sorts performed so far */
  int64_t kernel_launches;                   /* kernels launched since create/reset */
  /* device time (ms, CUDA events on the library stream) accumulated since reset */
  double ms_fields, ms_push, ms_bcs, ms_sort, ms_exchange;
  /* the fused push+gather+deposit kernel(s) alone, and how many times it was launched */
  double ms_push_kernel;
  int64_t n_push_kernel;
} cylgpu_stats_t;

const char* cylgpu_last_error(void);
int cylgpu_version(void);

/* in-process fabric joining `nranks` handles of one process (tests on a single GPU) */
void* cylgpu_fabric_create(int nranks);
void cylgpu_fabric_destroy(void* fabric);
/* fills 128 bytes; wraps ncclGetUniqueId of the dlopen()ed libnccl */
int cylgpu_nccl_unique_id(void* out128);

/* replaces: allocation in mpi_initialise (mpi_routines.F90:385-400) + setup of the
 * per-rank constants used by particles.F90:146-217 */
int cylgpu_create(const cylgpu_config* cfg, cylgpu_handle* out);
int cylgpu_destroy(cylgpu_handle h);
int cylgpu_set_species(cylgpu_handle h, int ispecies, const cylgpu_species* sp);
/* epoch2d.F90:157-161 plays with dt around the start-up half step */
int cylgpu_set_dt(cylgpu_handle h, double dt);
/* window.F90:341-349: bc_field swapped to bc_*_after_move, setup_boundaries re-run */
int cylgpu_set_bc_field(cylgpu_handle h, const int32_t bc_field[4]);
/* run the library's work on a caller stream (cudaStream_t), NULL = library-owned stream */
int cylgpu_set_stream(cylgpu_handle h, void* stream);
int cylgpu_synchronize(cylgpu_handle h);

/* host <-> device mirrors of the shared_data arrays.  `host` has the full Fortran extent
 * (nx+2ng)*(ny+2ng)*n_mode complex; snapshots are (ny+2ng)*n_mode complex. */
int cylgpu_upload_field(cylgpu_handle h, int field_id, const void* host);
int cylgpu_download_field(cylgpu_handle h, int field_id, void* host);
int cylgpu_upload_snapshot(cylgpu_handle h, int snap_id, const void* host);
int cylgpu_download_snapshot(cylgpu_handle h, int snap_id, void* host);
void* cylgpu_field_device_ptr(cylgpu_handle h, int field_id);
/* setup.F90:393-423 setup_field_boundaries, evaluated on the device arrays */
int cylgpu_snapshot_field_boundaries(cylgpu_handle h);

/* particles: AoS on the wire exactly like pack_particle (partlist.F90:414-428):
 * 7 doubles per particle = pos(3), p(3), weight.  `host_aos` is n*7 doubles. */
int cylgpu_upload_particles(cylgpu_handle h, int ispecies, int64_t n, const double* host_aos);
int cylgpu_append_particles(cylgpu_handle h, int ispecies, int64_t n, const double* host_aos);
int cylgpu_download_particles(cylgpu_handle h, int ispecies, int64_t capacity, double* host_aos,
                              int64_t* n_out);
int cylgpu_particle_count(cylgpu_handle h, int ispecies, int64_t* n_out);
/* per-particle (cell_x, cell_y) of split_particle.F90:62-63, int32 pairs, device order */
int cylgpu_particle_cells(cylgpu_handle h, int ispecies, int64_t capacity, int32_t* cells_out);
/* SoA device pointers: comp 0..6 = x y z px py pz w */
void* cylgpu_particle_device_ptr(cylgpu_handle h, int ispecies, int comp);

/* ---- the hot path, in driver order (epoch2d.F90:189-266) ---- */
/* fields.f90:316-337 update_eb_fields_half */
int cylgpu_fields_half(cylgpu_handle h);
/* particles.F90:28-734 push_particles, including current_bcs_r_min_final
 * (boundary.F90:1909) and particle_bcs (boundary.F90:1541) */
int cylgpu_push(cylgpu_handle h);
/* The same push_particles + particle_bcs for species whose particle list STAYS IN HOST MEMORY
 * (the reference's linked lists, shared_data.F90:159-171, flattened in pack_particle order,
 * partlist.F90:414-428: x y z px py pz w).  host_aos[isp] holds n_in[isp] particles and has
 * room for capacity[isp]; the list is streamed through the GPU in chunks (upload | sort + push +
 * deposit + boundary conditions | download run concurrently on three streams), the survivors
 * are written back in place, compacted, followed by the particles received from the
 * neighbours; n_out[isp] is the new count.  A NULL host_aos[isp] leaves species isp to its
 * device-resident list (pushed as by cylgpu_push, WITHOUT particle_bcs).  Use pinned memory
 * (cudaHostRegister on the Fortran array) for full PCIe rate. */
int cylgpu_push_host(cylgpu_handle h, const int64_t* n_in, double* const* host_aos, const int64_t* capacity,
                     int64_t* n_out);
/* particles per chunk of cylgpu_push_host (default 2^21) */
int cylgpu_set_host_chunk(cylgpu_handle h, int64_t particles);
/* Current smoothing of current_finish (smooth_current, current_smooth.F90:49-57,145-196): the
 * control block's smooth_currents, smooth_its, smooth_compensation (0/1) and smooth_strides
 * (deck_control_block.F90:447-466; nstrides = 0 means stride 1; strides up to ng = 5). */
int cylgpu_set_current_smoothing(cylgpu_handle h, int enable, int its, int comp_its, int nstrides,
                                 const int32_t* strides);
/* current_smooth.F90:29-45 current_finish (smoothing off) */
int cylgpu_current_finish(cylgpu_handle h);
/* fields.f90:341-353 update_eb_fields_final.  source1/source2 are the host-evaluated laser
 * sources (laser.f90:442-461) on ir = 0..ny for x_min and x_max; NULL = no laser there. */
int cylgpu_fields_final(cylgpu_handle h, const double* src1_xmin, const double* src2_xmin,
                        const double* src1_xmax, const double* src2_xmax);
/* window.F90:62-94 shift_window for ONE cell: append the host-generated column
 * (insert_particles, window.F90:157-300; n_new[is] particles each, AoS, may be NULL),
 * take the grid the host's setup_grid_x (utilities.f90:343-372) computed for the shifted
 * window, grid5 = {x_grid_min_local, x_min, x_max, x_min_local, x_max_local},
 * remove_particles (x_min rank), shift_fields.  The caller runs cylgpu_particle_bcs
 * afterwards exactly like window.F90:364. */
int cylgpu_window_shift(cylgpu_handle h, const int64_t* n_new, const double* const* new_aos,
                        const double* grid5);

/* ---- plasma column of the moving window with the reference's random stream ----
 * insert_particles (window.F90:157-300) draws from the rank's KISS stream
 * (random_generator.f90:45-173), which the loader used before it: either hand the stream over
 * (cylgpu_rng_set_state with the six members of random_state_type after loading) or start it
 * as the reference does (cylgpu_rng_init(7842432 + rank), setup.F90:563-567).
 * cylgpu_rng_flush_cache mirrors random_flush_cache (diagnostics.F90:235, once per step). */
int cylgpu_rng_init(cylgpu_handle h, int seed);
int cylgpu_rng_set_state(cylgpu_handle h, const int32_t* xyzw, int box_muller_cached, double cached_random_value);
int cylgpu_rng_get_state(cylgpu_handle h, int32_t* xyzw, int* box_muller_cached, double* cached_random_value);
int cylgpu_rng_flush_cache(cylgpu_handle h);
int cylgpu_rng_uniform(cylgpu_handle h, double* out);
/* insert_particles for one species, to be called for every species in order BEFORE
 * cylgpu_window_shift (shift_window, window.F90:62-94).  x_grid_max = x_global(nx_global)
 * before the shift; density(0:ny+1), temperature(0:ny+1,1:3), drift(0:ny+1,1:3): the deck
 * functions on the column ix = nx as the reference evaluates them (window.F90:203-220),
 * Fortran order; dmin/dmax = initial_conditions%density_min/max.  A no-op on ranks that do
 * not own x_max.  The new particles are appended to the device list. */
int cylgpu_insert_particles(cylgpu_handle h, int ispecies, double x_grid_max, double npart_per_cell,
                            const double* density, const double* temperature, const double* drift, double dmin,
                            double dmax, int64_t* n_inserted);

/* ... and for a species whose list lives in host memory (cylgpu_push_host): the column is written behind the
 * *n_inout records of host_aos (capacity records of 7 doubles) and *n_inout grows by it.  The removal of the plasma
 * behind the window (remove_particles, window.F90:304-325) needs no call for such lists: cylgpu_window_shift notes
 * the new x_min and the next cylgpu_push_host drops what lies behind it while the list streams through the GPU. */
int cylgpu_insert_particles_host(cylgpu_handle h, int ispecies, double x_grid_max, double npart_per_cell,
                                 const double* density, const double* temperature, const double* drift, double dmin,
                                 double dmax, double* host_aos, int64_t capacity, int64_t* n_inout);

/* The same column generated ON THE DEVICE from a counter-based stream: one
 * kernel writes the new particles straight into the SoA list, with no host loop or upload, and
 * the plasma does not depend on the number of ranks.  Same arguments and per-particle arithmetic
 * as cylgpu_insert_particles (window.F90:187-298); NOT bit-identical to the reference's column
 * (documented non-bit-parity mode): the uniforms come from Philox4x32-10 with
 *   key = (seed low word, seed high word + ispecies),
 *   counter = (column low word, column high word, radial cell iy, 4 * ip + block)
 * for particle ip of cell iy (blocks 0..3; (.., .., iy, 0xFFFFFFFF) decides the cell's fractional
 * particle), 53-bit uniforms from word pairs, and the momenta use the trigonometric Box-Muller
 * transform.  column = a number that is unique per inserted column (the total number of window
 * shifts so far).  cylgpu_philox4x32 evaluates the generator on the host. */
int cylgpu_insert_particles_device(cylgpu_handle h, int ispecies, double x_grid_max, double npart_per_cell,
                                   const double* density, const double* temperature, const double* drift,
                                   double dmin, double dmax, uint64_t seed, uint64_t column, int64_t* n_inserted);
int cylgpu_philox4x32(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4);

/* ---- pieces, exposed because the reference calls them on their own ---- */
int cylgpu_update_e_field(cylgpu_handle h);                 /* fields.f90:53-182 */
int cylgpu_update_b_field(cylgpu_handle h);                 /* fields.f90:186-312 */
int cylgpu_efield_bcs(cylgpu_handle h);                     /* boundary.F90:1355-1413 */
int cylgpu_bfield_bcs(cylgpu_handle h, int mpi_only);       /* boundary.F90:1417-1476 */
int cylgpu_bfield_final_bcs(cylgpu_handle h, const double* src1_xmin, const double* src2_xmin,
                            const double* src1_xmax, const double* src2_xmax);  /* :1505-1537 */
int cylgpu_particle_bcs(cylgpu_handle h);                   /* boundary.F90:1541-1889 */
int cylgpu_push_no_bcs(cylgpu_handle h);                    /* particles.F90:163-730 + :1909 */
int cylgpu_current_bcs(cylgpu_handle h);                    /* boundary.F90:1893-1905 */
/* Momentum rotation: 0 = Boris (default build), 1 = Higuera-Cary (the reference's -DHC_PUSH
 * build, particles.F90:409-421). */
int cylgpu_set_pusher(cylgpu_handle h, int higuera_cary);
/* Test knob: |m dtheta| below which the deposit factors m_fac_1..4 use the small-angle series.  The reference
 * hard-codes 1.0e-4 (particles.F90:593), which is the default and what every run uses; tests/ move it on both
 * sides to show that the closed forms just above the switch, not the kernels, set the parity of hot decks. */
int cylgpu_set_taylor_switch(cylgpu_handle h, double threshold);
/* cell-tile sort of the SoA arrays (no reference counterpart: replaces the linked list) */
int cylgpu_sort_particles(cylgpu_handle h);
int cylgpu_set_sort_interval(cylgpu_handle h, int every_n_pushes);   /* 0 = never */
/* 0 = one thread per particle, reductions at L2; 1 = warp-window reduce-scatter; 2 = strip CTAs; 3 = strip CTAs with
 * the FP64 tensor-op deposit (default); 4 = the shape-generic per-particle kernel (csrc/push_shapes.cuh), the only
 * one in the top-hat / B-spline builds */
int cylgpu_set_push_variant(cylgpu_handle h, int variant);
/* The absorbing / laser boundaries of laser.f90 hold two pieces of non-conforming Fortran whose outcome the library
 * reproduces by default (on = 1), because that is what a gfortran build of the reference computes: r_d_vals(0:ny)
 * and, on x_max, source_t(0:ny) are used whole against (1:ny) sections (laser.f90:474,587,604), so element ir takes
 * the radius and the source of ir - 1; and icdt_2r of outflow_bcs_r_max is declared REAL but assigned 0.5 i c dt / r
 * (:640,648), so the azimuthal coupling terms of the r_max line updates vanish.  on = 0: element for element, and the
 * coefficient as the right-hand side spells it -- for a reference built with those lines corrected. */
int cylgpu_set_reference_quirks(cylgpu_handle h, int on);
/* The particle shape is a compile-time choice of the reference (-DPARTICLE_SHAPE_TOPHAT / _BSPLINE3, constants.F90:
 * 524-545) and of this library (-DCYL_SHAPE=1 / 2: libcylgpu_tophat.so, libcylgpu_bspline3.so): it sets ng = png + 2,
 * i.e. the layout (1-ng:nx+ng, 1-ng:ny+ng, 0:M-1) of every array that crosses this interface.  cylgpu_shape: 0
 * triangle, 1 top-hat, 2 third-order B-spline; cylgpu_ghost_cells: ng (5, 4, 6). */
int cylgpu_shape(void);
int cylgpu_ghost_cells(void);

/* calc_number_density_modes (calc_df.F90:588-661) of one species (ispecies >= 0) or of all
 * current-carrying species (ispecies < 0), computed from the device-resident lists including
 * calc_boundary_modes and the zero-gradient ghost fill; host_out is a complex(num) array
 * (1-ng:nx+ng, 1-ng:ny+ng, 0:n_mode-1).  Lets dump steps skip the particle download. */
int cylgpu_number_density_modes(cylgpu_handle h, int ispecies, void* host_out);
/* calc_charge_density (calc_df.F90:442-519): real array (1-ng:nx+ng, 1-ng:ny+ng) */
int cylgpu_charge_density(cylgpu_handle h, int ispecies, double* host_out);
/* The other particle moments of io/calc_df.F90, from the device-resident lists (no particle
 * download at dump steps).  host_out: real array (1-ng:nx+ng, 1-ng:ny+ng), with the reference's
 * calc_boundary and field_zero_gradient ghost treatment.  ispecies < 0 is the reference's
 * `current_species <= 0` (all species that carry current).  direction: 1, 2, 3 = c_dir_x, _y, _z
 * (constants.F90:231-233), negative for the backward ekflux, 0 = argument absent. */
enum {
  CYLGPU_MOM_MASS_DENSITY = 0,     /* calc_mass_density        calc_df.F90:59-136   */
  CYLGPU_MOM_NUMBER_DENSITY = 1,   /* calc_number_density      calc_df.F90:523-584  */
  CYLGPU_MOM_EKBAR = 2,            /* calc_ekbar               calc_df.F90:140-245  */
  CYLGPU_MOM_EKFLUX = 3,           /* calc_ekflux              calc_df.F90:249-391  (direction +-1..3) */
  CYLGPU_MOM_PPC = 4,              /* calc_ppc                 calc_df.F90:665-712  */
  CYLGPU_MOM_AVERAGE_WEIGHT = 5,   /* calc_average_weight      calc_df.F90:716-778  */
  CYLGPU_MOM_TEMPERATURE = 6,      /* calc_temperature         calc_df.F90:782-1033 (direction 0..3) */
  CYLGPU_MOM_SPECIES_CURRENT = 7,  /* calc_per_species_current calc_df.F90:1037-1139 (direction 1..3) */
  CYLGPU_MOM_AVERAGE_MOMENTUM = 8  /* calc_average_momentum    calc_df.F90:1143-1221 (direction 1..3) */
};
int cylgpu_particle_moment(cylgpu_handle h, int kind, int ispecies, int direction, double* host_out);
/* ---- SDF dump / restart of the hot-path state straight from the device mirrors ----
 * (SURVEY.md 8(f)4).  One SDF 1.4 file in the reference's layout: the 'grid' mesh, the 30 mode-array
 * blocks of write_mode_field (io/diagnostics.F90:497-575,2033-2110: ids exm_real .. jtm_old_imag,
 * dims (nx_global, ny_global, n_mode), the r-staggered arrays shifted by one row so that file row 1
 * is the axis) and per species 'grid/<name>' with weight/ px/ py/ pz/<name> (:3040-3160).  Every rank
 * of an x-slab run calls with the same path and writes its own pieces (pwrite at offsets that follow
 * from the global sizes); the rank that owns x_min adds the metadata.  The caller supplies what the
 * reference gets from MPI: npart_global and npart_offset (species_offset) per species.  With
 * have_extents = 0 the particle-grid extents in the metadata are those of the writing rank. */
#define CYLGPU_SDF_MAX_CONSTANTS 16
typedef struct {
  int32_t nx_global, ny_global, n_mode, n_species;
  int32_t nx_local, cell_x_min;          /* this rank's global cells cell_x_min .. cell_x_min + nx_local - 1 (1-based) */
  int32_t step, restart, jobid1, jobid2; /* file header: step, restart_flag, jobid (sdf_write_header) */
  int32_t have_extents;
  /* derived variables of write_nspecies_field (io/diagnostics.F90:765-835) to add to the dump, computed
   * on the device by cylgpu_sdf_dump: bit v of derived_mask selects entry v of the reference's list
   * (0 ekbar, 1 mass_density, 2 charge_density, 3 number_density, 4 ppc, 5 average_weight, 6..8
   * average_px/py/pz, 9 temperature, 10..12 temperature_x/y/z, 13..15 jx/jy/jz, 16..21 ekflux x_max,
   * y_max, z_max, x_min, y_min, z_min; 22 number_density_mode: write_nspecies_field_mode :781-783,2596-2676,
   * per species 'Number_Density_Mode/<species>/Real' and '/Imaginary', complex arrays behind the others);
   * derived_sum: the species-summed block 'Derived/<Name>'
   * (dump_sum), derived_species: one block per species 'Derived/<Name>/<species>' (dump_species) */
  uint32_t derived_mask;
  int32_t derived_sum, derived_species;
  double time, x_min, dx, dy;            /* xb_global(i) = x_min + (i-1) dx, yb_global(j) = (j-1) dy */
  const char* species_name[CYLGPU_MAX_SPECIES];
  int64_t npart_global[CYLGPU_MAX_SPECIES], npart_offset[CYLGPU_MAX_SPECIES], npart_local[CYLGPU_MAX_SPECIES];
  double part_extents[CYLGPU_MAX_SPECIES][6];   /* min x, y, z then max x, y, z */
  /* real-valued constant blocks (sdf_write_srl, io/diagnostics.F90:403-416): what the driver owns and a
   * restart needs back -- 'dt', 'window_shift_fraction', 'x_grid_min', 'elapsed_time', energies ...  Written
   * by the rank that owns x_min; on reading, constant_value[k] is filled for every listed id that the file
   * holds and bit k of constants_found is set */
  int32_t n_constants;
  uint32_t constants_found;
  const char* constant_id[CYLGPU_SDF_MAX_CONSTANTS];
  const char* constant_name[CYLGPU_SDF_MAX_CONSTANTS];
  double constant_value[CYLGPU_SDF_MAX_CONSTANTS];
} cylgpu_sdf_desc;
/* host arrays in, no device needed: fields15[id] = complex(num) (1-ng:nx_local+ng, 1-ng:ny+ng, 0:n_mode-1)
 * for the 15 field ids above, particles_aos[s] = npart_local[s] records of 7 doubles */
#define CYLGPU_SDF_NDERIVED 22
int cylgpu_sdf_write_host(const char* path, const cylgpu_sdf_desc* d, const void* const* fields15,
                          const double* const* particles_aos, const double* const* derived);
/* derived: one real array (1-ng:nx_local+ng, 1-ng:ny+ng) per derived block, in file order -- for every
 * selected variable the sum block (if derived_sum) then the species blocks (if derived_species);
 * cylgpu_sdf_derived_count(d) says how many.  NULL when derived_mask is 0. */
int cylgpu_sdf_derived_count(const cylgpu_sdf_desc* d);
/* the inverse for this rank's slab: interior of the 15 arrays (ghosts untouched; the shift undone as in
 * housekeeping/setup.F90:1199-1210) and the particles with x_lo <= x < x_hi; fills step, time,
 * npart_global and npart_local; particles_aos may be NULL (counts only), else capacity[s] records each */
int cylgpu_sdf_read_host(const char* path, cylgpu_sdf_desc* d, void* const* fields15, double x_lo, double x_hi,
                         double* const* particles_aos, const int64_t* capacity);
/* the same between the file and the device-resident state.  dump: npart_local is taken from the
 * device lists; on one rank npart_global / npart_offset / extents are filled in too.  load: ghosts
 * are zeroed, the caller re-derives them (cylgpu_efield_bcs, cylgpu_bfield_bcs, cylgpu_current_finish
 * as the restart of the reference does through its boundary routines) */
int cylgpu_sdf_dump(cylgpu_handle h, const char* path, cylgpu_sdf_desc* d);
int cylgpu_sdf_load(cylgpu_handle h, const char* path, cylgpu_sdf_desc* d);

/* diagnostics the new code must own: field + kinetic energy from the
 * mode arrays with cylindrical volume elements; out[0] = field J, out[1] = kinetic J */
int cylgpu_energy(cylgpu_handle h, double* out2);
int cylgpu_stats(cylgpu_handle h, cylgpu_stats_t* out);
int cylgpu_reset_stats(cylgpu_handle h);
/* Device-resident particle counts: replaces the count-then-data MPI_SENDRECV pair of partlist_sendrecv
 * (partlist.F90:842,869) and every host-side use of a list length inside a step.
 * capacity > 0: the migrants of one direction travel in ONE message of fixed size -- a count header and
 * `capacity` particle slots of 7 doubles -- the leaver counts, the compaction, the arrivals and the window's
 * removals / insertions are handled by kernels that read the counts on the device, and no call of the step
 * (fields_half, push, current_finish, fields_final, window_shift, insert_particles, particle_bcs) waits for the
 * device.  The host keeps upper bounds of the list lengths and tightens them from copies that trail behind;
 * any other entry point (cylgpu_particle_count, cylgpu_stats, downloads, diagnostics) first waits for the newest
 * copy and sees exact counts.  More than `capacity` particles leaving towards one neighbour in one step is an
 * error (sticky, reported by the next such call); particles move less than a cell per step, so
 * 4 * (ny + 2) * particles-per-cell is a safe capacity for x-slabs.  Every rank must set the same value.
 * capacity = 0 (default): the exact protocol, counts first and then the payload, two host syncs per species. */
int cylgpu_set_exchange_capacity(cylgpu_handle h, int64_t capacity);
/* ---- the main-loop body, natively ----
 * What the reference's PROGRAM pic does between the hot-path calls is host work there too: time / step
 * bookkeeping, the laser source evaluation of outflow_bcs_x_min / x_max (laser.f90:276-328,442-461,556-575), the
 * moving-window trigger and shift_window's grid update (window.F90:62-94,330-376, utilities.f90:343-372).  A host
 * that owns these (the Fortran driver) calls the entry points above one by one; a host that does not (a C / C++ /
 * Python caller, bench.py) configures them once and lets the library run whole steps in the reference's order
 * (epoch2d.F90:189-266, optional physics packages off) with a handful of launches per step and no interpreter in
 * between -- with x-slabs over 8 GPUs a step is a few milliseconds. */
typedef struct cylgpu_laser {      /* laser_block (laser.f90) restricted to what the decks in scope use */
  int32_t boundary, pad_;          /* CYLGPU_BD_X_MIN / _X_MAX */
  double amp, omega, pol_angle;    /* amp = 100 sqrt(I[W/cm2] / (c eps0 / 2)), deck_laser_block.f90:135-139 */
  double t_start, t_end;
  double t_centre, t_width;        /* t_profile = gauss(time, t_centre, t_width); t_width <= 0: constant */
  double r_width;                  /* profile = gauss(y, 0, r_width); <= 0: flat */
  double phase, phase_curv;        /* phase(y) = phase + phase_curv y^2 */
} cylgpu_laser;
typedef struct cylgpu_insert_profile {   /* uniform plasma of the window's new column (window.F90:203-220) */
  double npart_per_cell, density;        /* npart_per_cell <= 0 or density <= 0: nothing is inserted */
  double temp[3], drift[3];
  double density_min, density_max;
} cylgpu_insert_profile;
typedef struct cylgpu_driver_config {
  int32_t cell_x_min;              /* first global cell of this rank, 1-based (mpi_routines.F90:312-337) */
  int32_t move_window;
  int32_t raw_bc_field[4];         /* the deck's bc_field BEFORE setup_boundaries normalises it */
  int32_t bc_x_min_after_move, bc_x_max_after_move;
  int32_t n_lasers;
  int32_t insert_mode;             /* 0: cylgpu_insert_particles (the rank's KISS stream), 1: _device (Philox) */
  uint64_t insert_seed;
  double x_grid_min;               /* global x(1) = x_min + dx / 2 */
  double xb_min;                   /* global cell-edge origin: window.F90 advances it beside x_grid_min, so a run taken
                                      over mid-way hands both (a fresh run: x_min) */
  double window_v_x, window_start_time, window_stop_time;
  const cylgpu_laser* lasers;
  cylgpu_insert_profile insert[CYLGPU_MAX_SPECIES];
  /* where the run stands (0 for a fresh run; from a restart dump otherwise) */
  double time, window_shift_fraction;
  int64_t step, window_shifts_total;
  int32_t window_started, pad_;
} cylgpu_driver_config;
typedef struct cylgpu_driver_state {
  double time;
  int64_t step;
  int32_t window_started, pad_;
  double window_shift_fraction;
  int64_t window_shifts_total;
  double x_grid_min, x_min, x_max, x_grid_min_local, x_min_local, x_max_local;
  int32_t bc_field[4];
  int32_t raw_bc_field[4];         /* the deck's values, bc_x_*_after_move once the window started (window.F90:342-350) */
} cylgpu_driver_state;
int cylgpu_driver_configure(cylgpu_handle h, const cylgpu_driver_config* cfg);
int cylgpu_driver_init_half_step(cylgpu_handle h);          /* epoch2d.F90:143-161 */
int cylgpu_driver_step(cylgpu_handle h, int64_t nsteps);    /* epoch2d.F90:189-266, nsteps times */
int cylgpu_driver_get_state(cylgpu_handle h, cylgpu_driver_state* out);
int cylgpu_driver_set_time(cylgpu_handle h, double time, int64_t step);
/* ---- slab re-balancer (balance.F90 for nprocy = 1) ----
 * cylgpu_load_x: part_load_func (balance.F90:2453-2478) from the device-resident lists: macro-particles of all
 * species per column of this slab, load_out[ix + ng - 1] for ix = 1-ng .. nx+ng (nx + 2 ng entries).  The caller
 * sums the slabs' columns into the global profile (the reference's MPI_ALLREDUCE), multiplies by push_per_field
 * (5, shared_data.F90:761) and adds ny_global per interior column (get_load, balance.F90:2322-2365).
 * cylgpu_calculate_breaks: calculate_breaks (balance.F90:2510-2653) on such a profile, load[0 .. sz + 2 ng) =
 * load(1-ng : sz+ng); mins / maxs receive the 1-based inclusive cell range of each of the nproc slabs.  Host
 * arithmetic, no device needed.  Moving the columns and particles to their new owners (redistribute_domain,
 * distribute_particles) is done by the host through download / create / upload: cylindrical_epoch_b200/balance.py. */
int cylgpu_load_x(cylgpu_handle h, int64_t* load_out);
int cylgpu_calculate_breaks(const int64_t* load, int32_t sz, int32_t nproc, int32_t* mins, int32_t* maxs);
/* How the neighbour exchanges of this handle travel: out4 = {transport kind (CYLGPU_TRANSPORT_*), 1 if the left
 * link goes through peer-memory mailboxes over NVLink (CUDA IPC mapping of the neighbour's buffer: one kernel on
 * each side per message, no rendezvous), the same for the right link, slot size of the mailboxes in bytes / 1024}.
 * With the NCCL transport the mailboxes are set up at cylgpu_create / cylgpu_set_exchange_capacity where both ends of
 * a link can map each other (CYLGPU_P2P=0 disables them); ncclSend / ncclRecv remain the fallback per link. */
int cylgpu_transport_info(cylgpu_handle h, int32_t* out4);
/* per-phase CUDA-event timers in cylgpu_stats (adds a host sync per phase); default off */
int cylgpu_set_timing(cylgpu_handle h, int on);

#ifdef __cplusplus
}
#endif
#endif /* CYLGPU_H */
