// bcs.cu -- field / current boundary conditions, x-halo exchange, laser + outflow line
// updates, axis current fold, boundary snapshots and the moving-window field shift.
// Replaces boundary.F90 (field_mode_bc, field_mode_clamp_zero, field_mode_zero_gradient,
// particle_reflection_bcs_complex, particle_periodic_bcs_complex, efield_bcs, bfield_bcs,
// bfield_final_bcs, current_bcs, current_bcs_r_min_final), laser.f90 outflow_bcs_*,
// current_smooth.F90 current_finish and window.F90 shift_fields.
#include <cfloat>
#include <cstring>

#include "ctx.cuh"

namespace cylgpu {

#include "bc_kernels.cuh"

static void launch_edge(cylgpu_ctx* c, const Tri& t, int bd) {
  const Geom& g = c->g;
  if (bd == CYLGPU_BD_Y_MAX) {
    k_edge_y<<<dim3((g.SX + 127) / 128, g.M, 3), 128, 0, c->stream>>>(g, t);
  } else {
    k_edge_x<<<dim3((g.SY + 127) / 128, g.M, 3), 128, 0, c->stream>>>(g, t, bd);
  }
  c->stats.kernel_launches += 1;
}

// `gg`: geometry of the arrays when it is not the handle's (the single-plane work arrays of the
// particle moments, moments.cuh); the staging buffers are sized for the handle's n_mode >= 1.
static int exchange3(cylgpu_ctx* c, const Halo3& h, int mode, bool send_l, bool send_r, bool recv_l,
                     bool recv_r, const Geom* gg = nullptr) {
  const Geom& g = gg ? *gg : c->g;
  if (!(send_l || send_r || recv_l || recv_r)) return 0;
  const size_t halo_elems = gg ? (size_t)3 * g.M * g.SY * NG : c->halo_elems;
  const size_t bytes = (mode == 2 ? 2 : 1) * halo_elems * sizeof(cplx);
  dim3 grd((g.SY * NG + 127) / 128, g.M, 3);
  if (send_l || send_r) {
    k_halo_pack<<<grd, 128, 0, c->stream>>>(g, h, send_l ? c->sbuf_l : nullptr, send_r ? c->sbuf_r : nullptr,
                                            mode, halo_elems);
    c->stats.kernel_launches += 1;
  }
  TRY(transport_sendrecv(c, send_l ? c->sbuf_l : nullptr, send_l ? bytes : 0, recv_l ? c->rbuf_l : nullptr,
                         recv_l ? bytes : 0, send_r ? c->sbuf_r : nullptr, send_r ? bytes : 0,
                         recv_r ? c->rbuf_r : nullptr, recv_r ? bytes : 0));
  if (recv_l || recv_r) {
    k_halo_unpack<<<grd, 128, 0, c->stream>>>(g, h, recv_l ? c->rbuf_l : nullptr, recv_r ? c->rbuf_r : nullptr,
                                              mode, halo_elems);
    c->stats.kernel_launches += 1;
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// field_mode_bc on three arrays at once: one packed message per neighbour instead of the
// reference's 3 * n_mode MPI_SENDRECV pairs.
int halo_x(cylgpu_ctx* c, int f0, int f1, int f2, int skip0, int skip1, int skip2) {
  Halo3 h;
  h.f[0] = c->f[f0]; h.f[1] = c->f[f1]; h.f[2] = c->f[f2];
  h.skip[0] = skip0; h.skip[1] = skip1; h.skip[2] = skip2;
  // MPI_SENDRECV with MPI_PROC_NULL neighbours is a no-op; ghost columns are only filled
  // where `.NOT. x_max_boundary .OR. bc_field == periodic` (boundary.F90:531,544)
  const bool has_l = c->left >= 0, has_r = c->right >= 0;
  const bool fill_r = has_r && (!c->cfg.x_max_boundary || c->bc_field[CYLGPU_BD_X_MAX] == CYLGPU_BC_PERIODIC);
  const bool fill_l = has_l && (!c->cfg.x_min_boundary || c->bc_field[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC);
  // what I send left is what my left neighbour uses to fill ITS right ghosts, and vice versa;
  // with x-slabs the neighbour's fill condition mirrors mine
  return exchange3(c, h, 0, fill_l, fill_r, fill_l, fill_r);
}

// ---- efield_bcs / bfield_bcs: boundary.F90:1355-1476 ----
static void ops_for(int bc, int conduct0, int conduct1, int conduct2, int* op) {
  op[0] = op[1] = op[2] = OP_NONE;
  if (bc == CYLGPU_BC_CONDUCT) { op[0] = conduct0; op[1] = conduct1; op[2] = conduct2; }
}

static void general_ops(int bc, int* op) {
  op[0] = op[1] = op[2] = OP_NONE;
  if (bc == CYLGPU_BC_CLAMP || bc == CYLGPU_BC_SIMPLE_LASER || bc == CYLGPU_BC_SIMPLE_OUTFLOW)
    op[0] = op[1] = op[2] = OP_CLAMP;
  if (bc == CYLGPU_BC_ZERO_GRADIENT || bc == CYLGPU_BC_CPML_LASER || bc == CYLGPU_BC_CPML_OUTFLOW)
    op[0] = op[1] = op[2] = OP_ZEROGRAD;
}

static bool on_boundary(const cylgpu_ctx* c, int bd) {
  if (bd == CYLGPU_BD_X_MIN) return c->cfg.x_min_boundary != 0;
  if (bd == CYLGPU_BD_X_MAX) return c->cfg.x_max_boundary != 0;
  return true;   // nprocy = 1
}

// stagger flags (setup.F90:126-136): Exm (x:F,r:T) Erm (T,F) Etm (T,T) Bxm (T,F) Brm (F,T) Btm (F,F)
static const int STAG_X[6] = {0, 1, 1, 1, 0, 0};
static const int STAG_Y[6] = {1, 0, 1, 0, 1, 0};

static int edge_bcs(cylgpu_ctx* c, int base, const int cond_x[3], const int cond_y[3]) {
  Tri t;
  for (int k = 0; k < 3; ++k) t.f[k] = c->f[base + k];
  auto apply = [&](int bd, const int* op) {
    if (op[0] == OP_NONE && op[1] == OP_NONE && op[2] == OP_NONE) return;
    if (c->bc_field[bd] == CYLGPU_BC_PERIODIC) return;
    if (!on_boundary(c, bd)) return;
    for (int k = 0; k < 3; ++k) {
      t.op[k] = op[k];
      t.stag[k] = (bd == CYLGPU_BD_Y_MAX) ? STAG_Y[base + k] : STAG_X[base + k];
    }
    launch_edge(c, t, bd);
  };
  int op[3];
  for (int bd : {CYLGPU_BD_X_MIN, CYLGPU_BD_X_MAX}) {
    ops_for(c->bc_field[bd], cond_x[0], cond_x[1], cond_x[2], op);
    apply(bd, op);
  }
  ops_for(c->bc_field[CYLGPU_BD_Y_MAX], cond_y[0], cond_y[1], cond_y[2], op);
  apply(CYLGPU_BD_Y_MAX, op);
  for (int bd : {CYLGPU_BD_X_MIN, CYLGPU_BD_X_MAX, CYLGPU_BD_Y_MAX}) {
    general_ops(c->bc_field[bd], op);
    apply(bd, op);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// the ghost fills of efield_bcs / bfield_bcs on the domain boundaries, without the halo
int efield_edges(cylgpu_ctx* c) {
  const int cx[3] = {OP_CLAMP, OP_ZEROGRAD, OP_ZEROGRAD};
  const int cy[3] = {OP_ZEROGRAD, OP_CLAMP, OP_ZEROGRAD};
  return edge_bcs(c, CYLGPU_EXM, cx, cy);
}
int bfield_edges(cylgpu_ctx* c) {
  const int cx[3] = {OP_ZEROGRAD, OP_CLAMP, OP_CLAMP};
  const int cy[3] = {OP_CLAMP, OP_ZEROGRAD, OP_CLAMP};
  return edge_bcs(c, CYLGPU_BXM, cx, cy);
}

int do_efield_bcs(cylgpu_ctx* c) {
  TRY(halo_x(c, CYLGPU_EXM, CYLGPU_ERM, CYLGPU_ETM, 1, 0, 1));
  return efield_edges(c);
}

int do_bfield_bcs(cylgpu_ctx* c, bool mpi_only) {
  TRY(halo_x(c, CYLGPU_BXM, CYLGPU_BRM, CYLGPU_BTM, 0, 1, 0));
  if (mpi_only) return 0;
  return bfield_edges(c);
}

// which sides the field halo fills (boundary.F90:531,544)
void halo_sides(const cylgpu_ctx* c, bool* fill_l, bool* fill_r) {
  const bool has_l = c->left >= 0, has_r = c->right >= 0;
  *fill_r = has_r && (!c->cfg.x_max_boundary || c->bc_field[CYLGPU_BD_X_MAX] == CYLGPU_BC_PERIODIC);
  *fill_l = has_l && (!c->cfg.x_min_boundary || c->bc_field[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC);
}

// The communication-avoiding field phases (field_ranges.cuh) apply when a real neighbour fills the ghosts (a
// periodic wrap onto the slab itself is a local copy: nothing to save), the slab is at least two halos wide,
// and the currents are not smoothed (current_smooth.F90 leaves the ghosts of J unsmoothed, so a ghost column
// could not be advanced the way its owner advances it).  CYLGPU_WIDE_FIELDS=0 keeps the reference's exchanges.
bool wide_fields(const cylgpu_ctx* c) {
  static const bool env_on = [] { const char* e = getenv("CYLGPU_WIDE_FIELDS"); return !e || atoi(e) != 0; }();
  if (!env_on || c->smooth_currents || c->g.nx < 2 * NG) return false;
  bool fill_l, fill_r;
  halo_sides(c, &fill_l, &fill_r);
  if (!fill_l && !fill_r) return false;
  if (c->left == c->cfg.rank || c->right == c->cfg.rank) return false;
  return true;
}

FieldRanges wide_ranges(const cylgpu_ctx* c, int phase) {
  bool fill_l, fill_r;
  halo_sides(c, &fill_l, &fill_r);
  return field_ranges(phase, wide_fields(c), fill_l, fill_r, c->cfg.x_min_boundary != 0, c->cfg.x_max_boundary != 0,
                      c->g.nx);
}

// field_mode_bc on the six field arrays at once: the closing exchange of a wide field phase, one message per
// neighbour (the E block, then the B block; row skips as in efield_bcs / bfield_bcs)
int halo_eb(cylgpu_ctx* c) {
  const Geom& g = c->g;
  bool fill_l, fill_r;
  halo_sides(c, &fill_l, &fill_r);
  if (!fill_l && !fill_r) return 0;
  const size_t he = c->halo_elems;
  const dim3 grd((g.SY * NG + 127) / 128, g.M, 3);
  Halo3 h[2];
  for (int k = 0; k < 3; ++k) { h[0].f[k] = c->f[CYLGPU_EXM + k]; h[1].f[k] = c->f[CYLGPU_BXM + k]; }
  h[0].skip[0] = 1; h[0].skip[1] = 0; h[0].skip[2] = 1;
  h[1].skip[0] = 0; h[1].skip[1] = 1; h[1].skip[2] = 0;
  for (int q = 0; q < 2; ++q)
    k_halo_pack<<<grd, 128, 0, c->stream>>>(g, h[q], fill_l ? c->sbuf_l + q * he : nullptr,
                                            fill_r ? c->sbuf_r + q * he : nullptr, 0, he);
  const size_t bytes = 2 * he * sizeof(cplx);
  TRY(transport_sendrecv(c, fill_l ? c->sbuf_l : nullptr, fill_l ? bytes : 0, fill_l ? c->rbuf_l : nullptr,
                         fill_l ? bytes : 0, fill_r ? c->sbuf_r : nullptr, fill_r ? bytes : 0,
                         fill_r ? c->rbuf_r : nullptr, fill_r ? bytes : 0));
  for (int q = 0; q < 2; ++q)
    k_halo_unpack<<<grd, 128, 0, c->stream>>>(g, h[q], fill_l ? c->rbuf_l + q * he : nullptr,
                                              fill_r ? c->rbuf_r + q * he : nullptr, 0, he);
  c->stats.kernel_launches += 4;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

static FieldSet fieldset(cylgpu_ctx* c) {
  FieldSet F;
  F.exm = c->f[CYLGPU_EXM]; F.erm = c->f[CYLGPU_ERM]; F.etm = c->f[CYLGPU_ETM];
  F.bxm = c->f[CYLGPU_BXM]; F.brm = c->f[CYLGPU_BRM]; F.btm = c->f[CYLGPU_BTM];
  F.jxm = c->f[CYLGPU_JXM]; F.jrm = c->f[CYLGPU_JRM]; F.jtm = c->f[CYLGPU_JTM];
  F.bxo = c->f[CYLGPU_BXM_OLD]; F.bro = c->f[CYLGPU_BRM_OLD]; F.bto = c->f[CYLGPU_BTM_OLD];
  F.jxo = c->f[CYLGPU_JXM_OLD]; F.jro = c->f[CYLGPU_JRM_OLD]; F.jto = c->f[CYLGPU_JTM_OLD];
  return F;
}

// host laser sources -> device (outside any captured graph: the staging slot rotates)
int upload_laser_sources(cylgpu_ctx* c, const double* s1min, const double* s2min, const double* s1max,
                         const double* s2max) {
  const Geom& g = c->g;
  const int nsrc = g.ny + 1;
  // sources are host arrays evaluated by the Fortran laser blocks (laser.f90:442-461).  They are
  // copied into a pinned staging slot of the library (8 rotating slots, each guarded by an event),
  // so that the caller may reuse its arrays at once and the step needs no host sync here.
  const double* hs[4] = {s1min, s2min, s1max, s2max};
  {
    const size_t slot_doubles = (size_t)4 * nsrc;
    if (!c->src_stage) {
      CUDA_TRY(cudaMallocHost(&c->src_stage, 8 * slot_doubles * sizeof(double)));
      for (int k = 0; k < 8; ++k) CUDA_TRY(cudaEventCreateWithFlags(&c->src_event[k], cudaEventDisableTiming));
    }
    const int slot = c->src_slot;
    c->src_slot = (c->src_slot + 1) & 7;
    CUDA_TRY(cudaEventSynchronize(c->src_event[slot]));   // long done unless the host runs 8 steps ahead
    double* st = c->src_stage + slot * slot_doubles;
    for (int k = 0; k < 4; ++k) {
      if (hs[k]) std::memcpy(st + (size_t)k * nsrc, hs[k], nsrc * sizeof(double));
      else std::memset(st + (size_t)k * nsrc, 0, nsrc * sizeof(double));
    }
    CUDA_TRY(cudaMemcpyAsync(c->src, st, slot_doubles * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaEventRecord(c->src_event[slot], c->stream));
  }
  return 0;
}

// bfield_final_bcs, boundary.F90:1505-1537, with the sources already in c->src
// wide: inside a communication-avoiding final phase -- no halo before or after, the r_max line update also on
// the ghost columns field_ranges.cuh allows
int do_bfield_final_bcs_device(cylgpu_ctx* c, bool wide, int phase) {
  const Geom& g = c->g;
  if (wide) TRY(bfield_edges(c));
  else TRY(do_bfield_bcs(c, false));
  FieldSet F = fieldset(c);
  const int nsrc = g.ny + 1;
  dim3 grd((g.ny + 1 + 127) / 128, g.M);
  if (c->cfg.x_min_boundary) {
    const int b = c->bc_field[CYLGPU_BD_X_MIN];
    if (b == CYLGPU_BC_SIMPLE_LASER || b == CYLGPU_BC_SIMPLE_OUTFLOW) {
      k_outflow_x<<<grd, 128, 0, c->stream>>>(g, F, c->snap[1], c->snap[2], c->snap[3], c->snap[4], c->snap[5],
                                              c->src, c->src + nsrc, 0, c->cfg.dx, c->cfg.dy, c->dt,
                                              c->cfg.y_grid_min_local, c->reference_quirks ? 1 : 0);
      c->stats.kernel_launches += 1;
    }
  }
  if (c->cfg.x_max_boundary) {
    const int b = c->bc_field[CYLGPU_BD_X_MAX];
    if (b == CYLGPU_BC_SIMPLE_LASER || b == CYLGPU_BC_SIMPLE_OUTFLOW) {
      k_outflow_x<<<grd, 128, 0, c->stream>>>(g, F, c->snap[7], c->snap[8], c->snap[9], c->snap[10], c->snap[11],
                                              c->src + 2 * nsrc, c->src + 3 * nsrc, 1, c->cfg.dx, c->cfg.dy,
                                              c->dt, c->cfg.y_grid_min_local, c->reference_quirks ? 1 : 0);
      c->stats.kernel_launches += 1;
    }
  }
  if (c->bc_field[CYLGPU_BD_Y_MAX] == CYLGPU_BC_SIMPLE_OUTFLOW) {
    FieldRanges R = wide_ranges(c, phase);
    if (!wide) R = field_ranges(1, false, false, false, c->cfg.x_min_boundary != 0, c->cfg.x_max_boundary != 0, g.nx);
    const int ix0 = R.obx_lo < R.obt_lo ? R.obx_lo : R.obt_lo;
    const int ix1 = R.obx_hi > R.obt_hi ? R.obx_hi : R.obt_hi;
    k_outflow_r_max<<<dim3((ix1 - ix0 + 1 + 127) / 128, g.M), 128, 0, c->stream>>>(
        g, F, R.obx_lo, R.obx_hi, R.obt_lo, R.obt_hi, ix0, c->cfg.dx, c->cfg.dy, c->dt, c->cfg.y_grid_min_local,
        c->reference_quirks ? 1 : 0);
    c->stats.kernel_launches += 1;
  } else if (c->bc_field[CYLGPU_BD_Y_MAX] == CYLGPU_BC_ZERO_B) {
    k_zero_b_rmax<<<dim3((g.SX + 127) / 128, g.M), 128, 0, c->stream>>>(g, F.bxm, F.brm, F.btm);
    c->stats.kernel_launches += 1;
  }
  CUDA_TRY(cudaGetLastError());
  if (wide) return 0;
  return do_bfield_bcs(c, true);
}

int do_bfield_final_bcs(cylgpu_ctx* c, const double* s1min, const double* s2min, const double* s1max,
                        const double* s2max) {
  TRY(upload_laser_sources(c, s1min, s2min, s1max, s2max));
  return do_bfield_final_bcs_device(c, false);
}

int do_r_min_final(cylgpu_ctx* c) {
  const Geom& g = c->g;
  k_r_min_final<<<dim3((g.SX + 127) / 128, g.M), 128, 0, c->stream>>>(g, c->f[CYLGPU_JXM], c->f[CYLGPU_JRM],
                                                                     c->f[CYLGPU_JTM]);
  c->stats.kernel_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// bc_allspecies (boundary.F90:60-75): common particle bc of all species, -1 if mixed
static int bc_allspecies(const cylgpu_ctx* c, int bd) {
  int b = -2;
  for (int i = 0; i < c->cfg.n_species; ++i) {
    if (!c->species[i].set) continue;
    if (b == -2) b = c->species[i].sp.bc_particle[bd];
    else if (b != c->species[i].sp.bc_particle[bd]) return -1;
  }
  return b == -2 ? CYLGPU_BC_OPEN : b;
}

// `with_halo`: current_finish follows the sum with the J halo (field_mode_bc on jxm, jrm, jtm);
// when both exchanges go to the same neighbours they travel as one message.  *halo_done tells
// the caller whether the ghosts are already filled.
static int current_bcs_impl(cylgpu_ctx* c, bool with_halo, bool* halo_done) {
  const Geom& g = c->g;
  if (halo_done) *halo_done = false;
  int bca[4];
  for (int bd = 0; bd < 4; ++bd) {
    bca[bd] = bc_allspecies(c, bd);
    if (bd != CYLGPU_BD_Y_MIN && bca[bd] == -1) {
      set_error("mixed per-species particle boundary conditions are not supported");
      return 2;
    }
  }
  cplx *jx = c->f[CYLGPU_JXM], *jr = c->f[CYLGPU_JRM], *jt = c->f[CYLGPU_JTM];
  if (c->cfg.x_min_boundary && bca[CYLGPU_BD_X_MIN] == CYLGPU_BC_REFLECT) {
    k_jreflect_x<<<dim3((g.SY + 127) / 128, g.M, 3), 128, 0, c->stream>>>(g, jx, jr, jt, 0);
    c->stats.kernel_launches += 1;
  }
  if (c->cfg.x_max_boundary && bca[CYLGPU_BD_X_MAX] == CYLGPU_BC_REFLECT) {
    k_jreflect_x<<<dim3((g.SY + 127) / 128, g.M, 3), 128, 0, c->stream>>>(g, jx, jr, jt, 1);
    c->stats.kernel_launches += 1;
  }
  if (bca[CYLGPU_BD_Y_MAX] == CYLGPU_BC_REFLECT) {
    k_jreflect_y<<<dim3((g.SX + 127) / 128, g.M, 3), 128, 0, c->stream>>>(g, jx, jr, jt, c->cfg.dy,
                                                                         c->cfg.y_grid_min_local);
    c->stats.kernel_launches += 1;
  }
  CUDA_TRY(cudaGetLastError());
  // particle_periodic_bcs_complex, boundary.F90:1133-1203: ADD the neighbour's ghost columns
  // into my interior edge columns.  neighbour_local(+-1) is nulled on a domain boundary
  // whose particle bc is not periodic.
  const bool to_l = c->left >= 0 && !(c->cfg.x_min_boundary && bca[CYLGPU_BD_X_MIN] != CYLGPU_BC_PERIODIC);
  const bool to_r = c->right >= 0 && !(c->cfg.x_max_boundary && bca[CYLGPU_BD_X_MAX] != CYLGPU_BC_PERIODIC);
  Halo3 h;
  h.f[0] = jx; h.f[1] = jr; h.f[2] = jt;
  h.skip[0] = h.skip[1] = h.skip[2] = 0;
  if (with_halo) {
    const bool has_l = c->left >= 0, has_r = c->right >= 0;
    const bool fill_r = has_r && (!c->cfg.x_max_boundary || c->bc_field[CYLGPU_BD_X_MAX] == CYLGPU_BC_PERIODIC);
    const bool fill_l = has_l && (!c->cfg.x_min_boundary || c->bc_field[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC);
    if (fill_l == to_l && fill_r == to_r && (to_l || to_r)) {
      *halo_done = true;
      return exchange3(c, h, 2, to_l, to_r, to_l, to_r);
    }
  }
  return exchange3(c, h, 1, to_l, to_r, to_l, to_r);
}

int do_current_bcs(cylgpu_ctx* c) { return current_bcs_impl(c, false, nullptr); }

static int halo_ptrs(cylgpu_ctx* c, cplx* f0, cplx* f1, cplx* f2) {   // field_mode_bc on three arrays
  Halo3 h;
  h.f[0] = f0; h.f[1] = f1; h.f[2] = f2;
  h.skip[0] = h.skip[1] = h.skip[2] = 0;
  const bool has_l = c->left >= 0, has_r = c->right >= 0;
  const bool fill_r = has_r && (!c->cfg.x_max_boundary || c->bc_field[CYLGPU_BD_X_MAX] == CYLGPU_BC_PERIODIC);
  const bool fill_l = has_l && (!c->cfg.x_min_boundary || c->bc_field[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC);
  return exchange3(c, h, 0, fill_l, fill_r, fill_l, fill_r);
}

// The reference filters one array after the other through a work copy whose interior is
// rewritten after every pass and whose ghosts only change through field_mode_bc.  Here the three
// arrays go together (one packed halo message per pass) and the copy-back is replaced by
// ping-pong between two work sets that both start as full copies: outside the halo-filled
// columns their ghosts never change, and the halo-filled ones are refreshed from the source
// before every pass, so each pass reads exactly what the reference's wk_array holds.  As in the
// reference, beta keeps its initial value when alpha changes, alpha changes only after pass
// its+1, and the ghosts of J itself are left as current_finish's halo filled them.
static int do_smooth_current(cylgpu_ctx* c) {
  const Geom& g = c->g;
  std::vector<int> strides = c->smooth_strides.empty() ? std::vector<int>{1} : c->smooth_strides;
  for (int sdv : strides)
    if (sdv < 1 || sdv > NG) { set_error("smooth_strides must lie in 1..%d (sng <= jng)", NG); return 2; }
  const size_t bytes = g.plane * g.M * sizeof(cplx);
  cplx* J[3] = {c->f[CYLGPU_JXM], c->f[CYLGPU_JRM], c->f[CYLGPU_JTM]};
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < 3; ++k) {
      if (!c->smooth_wk[s][k]) CUDA_TRY(cudaMalloc(&c->smooth_wk[s][k], bytes));
      CUDA_TRY(cudaMemcpyAsync(c->smooth_wk[s][k], J[k], bytes, cudaMemcpyDeviceToDevice, c->stream));
    }
  double alpha = 0.5;
  const double beta = (1.0 - alpha) * 0.25;
  int cur = 0;
  const dim3 grd((g.nx + 127) / 128, g.ny, 3 * g.M);
  for (int it = 1; it <= c->smooth_its + c->smooth_comp_its; ++it) {
    for (int stride : strides) {
      cplx** W = c->smooth_wk[cur];
      cplx** D = c->smooth_wk[cur ^ 1];
      TRY(halo_ptrs(c, W[0], W[1], W[2]));
      Tri3 src{{W[0], W[1], W[2]}}, dst{{D[0], D[1], D[2]}};
      k_smooth<<<grd, 128, 0, c->stream>>>(g, src, dst, alpha, beta, stride);
      c->stats.kernel_launches += 1;
      cur ^= 1;
    }
    if (it > c->smooth_its) alpha = (double)c->smooth_its * 0.5 + 1.0;
  }
  cplx** W = c->smooth_wk[cur];
  Tri3 src{{W[0], W[1], W[2]}}, dst{{J[0], J[1], J[2]}};
  k_copy_interior<<<grd, 128, 0, c->stream>>>(g, src, dst);
  c->stats.kernel_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int do_current_finish(cylgpu_ctx* c) {   // current_smooth.F90:29-45
  bool halo_done = false;
  TRY(current_bcs_impl(c, true, &halo_done));
  if (!halo_done) TRY(halo_x(c, CYLGPU_JXM, CYLGPU_JRM, CYLGPU_JTM, 0, 0, 0));
  if (c->smooth_currents) TRY(do_smooth_current(c));
  return 0;
}

// species < 0: sum over the species that carry current (calc_df.F90:606-616).  Result in c->spare.
// charge: calc_charge_density (calc_df.F90:442-519) -- wdata = charge * weight, no azimuthal factors
// (mode 0 only; the real part is the answer).
int do_number_density_modes(cylgpu_ctx* c, int species, bool charge) {
  const Geom& g = c->g;
  int bca[4];
  for (int bd = 0; bd < 4; ++bd) {
    bca[bd] = bc_allspecies(c, bd);
    if (bd != CYLGPU_BD_Y_MIN && bca[bd] == -1) {
      set_error("mixed per-species particle boundary conditions are not supported");
      return 2;
    }
  }
  cplx* a = c->spare;
  CUDA_TRY(cudaMemsetAsync(a, 0, g.plane * g.M * sizeof(cplx), c->stream));
  for (int isp = 0; isp < c->cfg.n_species; ++isp) {
    const SpeciesState& S = c->species[isp];
    if (!S.set || S.n == 0) continue;
    if (species >= 0 && isp != species) continue;
    if (species < 0 && S.sp.zero_current) continue;
    k_number_density<<<(unsigned)((S.n + 255) / 256), 256, 0, c->stream>>>(
        g, S.d[0], S.d[1], S.d[2], S.d[6], S.n, (double*)a, c->x_grid_min_local, c->cfg.y_grid_min_local, c->cfg.dx,
        c->cfg.dy, charge ? S.sp.charge : 1.0, charge ? 1 : g.M);
    c->stats.kernel_launches += 1;
  }
  const dim3 gx_((g.SY + 127) / 128, g.M), gy_((g.SX + 127) / 128, g.M);
  if (c->cfg.x_min_boundary && bca[CYLGPU_BD_X_MIN] == CYLGPU_BC_REFLECT)
    k_density_reflect<<<gx_, 128, 0, c->stream>>>(g, a, CYLGPU_BD_X_MIN);
  if (c->cfg.x_max_boundary && bca[CYLGPU_BD_X_MAX] == CYLGPU_BC_REFLECT)
    k_density_reflect<<<gx_, 128, 0, c->stream>>>(g, a, CYLGPU_BD_X_MAX);
  if (bca[CYLGPU_BD_Y_MAX] == CYLGPU_BC_REFLECT) k_density_reflect<<<gy_, 128, 0, c->stream>>>(g, a, CYLGPU_BD_Y_MAX);
  // particle_periodic_bcs (boundary.F90:1019-1129): neighbours' ghost columns add into my edge columns
  {
    const bool to_l = c->left >= 0 && !(c->cfg.x_min_boundary && bca[CYLGPU_BD_X_MIN] != CYLGPU_BC_PERIODIC);
    const bool to_r = c->right >= 0 && !(c->cfg.x_max_boundary && bca[CYLGPU_BD_X_MAX] != CYLGPU_BC_PERIODIC);
    Halo3 h;
    h.f[0] = a; h.f[1] = nullptr; h.f[2] = nullptr;
    h.skip[0] = h.skip[1] = h.skip[2] = 0;
    TRY(exchange3(c, h, 1, to_l, to_r, to_l, to_r));
  }
  if (c->bc_field[CYLGPU_BD_X_MIN] != CYLGPU_BC_PERIODIC && c->cfg.x_min_boundary)
    k_density_zero_gradient<<<gx_, 128, 0, c->stream>>>(g, a, CYLGPU_BD_X_MIN);
  if (c->bc_field[CYLGPU_BD_X_MAX] != CYLGPU_BC_PERIODIC && c->cfg.x_max_boundary)
    k_density_zero_gradient<<<gx_, 128, 0, c->stream>>>(g, a, CYLGPU_BD_X_MAX);
  if (c->bc_field[CYLGPU_BD_Y_MIN] != CYLGPU_BC_PERIODIC)
    k_density_zero_gradient<<<gy_, 128, 0, c->stream>>>(g, a, CYLGPU_BD_Y_MIN);
  if (c->bc_field[CYLGPU_BD_Y_MAX] != CYLGPU_BC_PERIODIC)
    k_density_zero_gradient<<<gy_, 128, 0, c->stream>>>(g, a, CYLGPU_BD_Y_MAX);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int download_real_part_mode0(cylgpu_ctx* c, const cplx* a, double* host_out) {
  double* tmp = nullptr;
  CUDA_TRY(cudaMalloc(&tmp, c->g.plane * sizeof(double)));
  k_real_part<<<(unsigned)((c->g.plane + 255) / 256), 256, 0, c->stream>>>(a, tmp, c->g.plane);
  CUDA_TRY(cudaMemcpyAsync(host_out, tmp, c->g.plane * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaFree(tmp));
  return 0;
}

int do_snapshot(cylgpu_ctx* c) {
  const Geom& g = c->g;
  Snaps S;
  for (int k = 0; k < CYLGPU_NSNAPS; ++k) S.s[k] = c->snap[k];
  k_snapshot<<<dim3((g.SY + 127) / 128, g.M), 128, 0, c->stream>>>(g, fieldset(c), S);
  c->stats.kernel_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---- moving window: shift_fields, window.F90:98-153 ----
// All nine arrays move by one cell IN PLACE in one launch: a block owns one row of one array and walks it in
// chunks of 1024 elements -- every thread reads its four elements' right neighbours, the block synchronises, then
// writes; the only element a later chunk overwrites (its first) was read by the chunk before.  1 read + 1 write per
// element like the out-of-place copy (k_shift_x, kept for the CPU emulation), but the array pointers never change,
// so the captured field-phase graphs stay valid across window shifts.
struct Nine { cplx* f[9]; };
#define SHIFT_T 256
#define SHIFT_E 4
__global__ void __launch_bounds__(SHIFT_T) k_shift_x_inplace(Geom g, Nine A) {
  cplx* a = A.f[blockIdx.y] + (size_t)blockIdx.x * g.SX;   // row (im, ir) of array blockIdx.y
  const int SX = g.SX;
  for (int base = 0; base < SX; base += SHIFT_T * SHIFT_E) {
    cplx v[SHIFT_E];
#pragma unroll
    for (int k = 0; k < SHIFT_E; ++k) {
      const int i = base + k * SHIFT_T + threadIdx.x;
      if (i < SX) v[k] = a[i < SX - 1 ? i + 1 : i];   // the last ghost column keeps its value
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SHIFT_E; ++k) {
      const int i = base + k * SHIFT_T + threadIdx.x;
      if (i < SX) a[i] = v[k];
    }
  }
}

int do_shift_fields(cylgpu_ctx* c) {
  const Geom& g = c->g;
  {
    Nine A;
    for (int k = 0; k < 9; ++k) A.f[k] = c->f[k];   // exm erm etm bxm brm btm jxm jrm jtm
    k_shift_x_inplace<<<dim3((unsigned)(g.SY * g.M), 9), SHIFT_T, 0, c->stream>>>(g, A);
    c->stats.kernel_launches += 1;
  }
  CUDA_TRY(cudaGetLastError());
  // field_mode_bc on each shifted array (window.F90:147-150): the nine halos travel as ONE message per neighbour
  {
    const bool has_l = c->left >= 0, has_r = c->right >= 0;
    const bool fill_r = has_r && (!c->cfg.x_max_boundary || c->bc_field[CYLGPU_BD_X_MAX] == CYLGPU_BC_PERIODIC);
    const bool fill_l = has_l && (!c->cfg.x_min_boundary || c->bc_field[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC);
    if (fill_l || fill_r) {
      const size_t he = c->halo_elems;
      const dim3 grd((g.SY * NG + 127) / 128, g.M, 3);
      Halo3 h[3];
      for (int q = 0; q < 3; ++q) {
        for (int k = 0; k < 3; ++k) { h[q].f[k] = c->f[3 * q + k]; h[q].skip[k] = 0; }
        k_halo_pack<<<grd, 128, 0, c->stream>>>(g, h[q], fill_l ? c->sbuf_l + q * he : nullptr,
                                                fill_r ? c->sbuf_r + q * he : nullptr, 0, he);
      }
      const size_t bytes = 3 * he * sizeof(cplx);
      TRY(transport_sendrecv(c, fill_l ? c->sbuf_l : nullptr, fill_l ? bytes : 0, fill_l ? c->rbuf_l : nullptr,
                             fill_l ? bytes : 0, fill_r ? c->sbuf_r : nullptr, fill_r ? bytes : 0,
                             fill_r ? c->rbuf_r : nullptr, fill_r ? bytes : 0));
      for (int q = 0; q < 3; ++q)
        k_halo_unpack<<<grd, 128, 0, c->stream>>>(g, h[q], fill_l ? c->rbuf_l + q * he : nullptr,
                                                  fill_r ? c->rbuf_r + q * he : nullptr, 0, he);
      c->stats.kernel_launches += 6;
      CUDA_TRY(cudaGetLastError());
    }
  }
  if (c->cfg.x_max_boundary) {
    Snaps S;
    for (int k = 0; k < CYLGPU_NSNAPS; ++k) S.s[k] = c->snap[k];
    k_window_fill_xmax<<<dim3((g.SY + 127) / 128, g.M), 128, 0, c->stream>>>(g, fieldset(c), S);
    c->stats.kernel_launches += 1;
    CUDA_TRY(cudaGetLastError());
  }
  return 0;
}

#include "moments.cuh"

}  // namespace cylgpu
