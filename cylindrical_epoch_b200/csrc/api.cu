// api.cu -- the extern "C" boundary declared in include/cylgpu.h.
#include <algorithm>
#include <cstdarg>
#include <cstring>

#include <cstdlib>
#include <functional>

#include "ctx.cuh"
#include "philox.cuh"

namespace cylgpu {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// entry points of the step itself (field phases, push, current_finish, window, particle_bcs, column insertion):
// they work with device-resident particle counts and never wait for the device
static int check_handle_fields(cylgpu_handle h) {
  if (!h) { set_error("null cylgpu handle"); return 1; }
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) { set_error("cudaSetDevice(%d): %s", h->device, cudaGetErrorString(e)); return 1; }
  return 0;
}
// everything else sees exact particle counts on the host (no-op unless cylgpu_set_exchange_capacity > 0 has left
// the counts on the device: then this waits for the newest copy)
static int check_handle(cylgpu_handle h) {
  TRY(check_handle_fields(h));
  TRY(presort_join(h, false));
  TRY(flush_pending_remove(h));
  h->r_clean = false;   // whatever this call does to the lists, the next particle_bcs tests every rule
  return poll_counts(h, true);
}

static void set_neighbours(cylgpu_ctx* c) {
  // MPI_CART_CREATE periodicity, mpi_routines.F90:186-199: x is periodic when the x_min
  // field bc is periodic (or any species' x_min particle bc is)
  bool periodic = c->bc_field[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC;
  for (int i = 0; i < c->cfg.n_species; ++i)
    if (c->species[i].set && c->species[i].sp.bc_particle[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC) periodic = true;
  const int P = c->cfg.nranks, k = c->cfg.rank;
  c->left = (k - 1 >= 0) ? k - 1 : (periodic ? P - 1 : -1);
  c->right = (k + 1 < P) ? k + 1 : (periodic ? 0 : -1);
}

}  // namespace cylgpu

using namespace cylgpu;

extern "C" void cylgpu_driver_release(void* driver);   // driver.cu

extern "C" {

const char* cylgpu_last_error(void) { return g_err; }
int cylgpu_version(void) { return 100; }

static int create_impl(const cylgpu_config* cfg, cylgpu_ctx* c) {
  c->cfg = *cfg;
  if (cfg->device >= 0) c->device = cfg->device;
  else CUDA_TRY(cudaGetDevice(&c->device));
  CUDA_TRY(cudaSetDevice(c->device));
  Geom& g = c->g;
  g.nx = cfg->nx; g.ny = cfg->ny; g.M = cfg->n_mode;
  g.SX = g.nx + 2 * NG; g.SY = g.ny + 2 * NG;
  g.plane = (size_t)g.SX * g.SY;
  c->dt = cfg->dt;
  c->x_grid_min_local = cfg->x_grid_min_local;
  c->x_min = cfg->x_min; c->x_max = cfg->x_max;
  c->x_min_local = cfg->x_min_local; c->x_max_local = cfg->x_max_local;
  for (int i = 0; i < 4; ++i) c->bc_field[i] = cfg->bc_field[i];
  std::memset(&c->stats, 0, sizeof(c->stats));
  c->timing = false;
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
  const size_t nelem = g.plane * g.M;
  for (int k = 0; k < CYLGPU_NFIELDS; ++k) {
    CUDA_TRY(cudaMalloc(&c->f[k], nelem * sizeof(cplx)));
    CUDA_TRY(cudaMemsetAsync(c->f[k], 0, nelem * sizeof(cplx), c->stream));
  }
  CUDA_TRY(cudaMalloc(&c->spare, nelem * sizeof(cplx)));
  for (int k = 0; k < CYLGPU_NSNAPS; ++k) {
    CUDA_TRY(cudaMalloc(&c->snap[k], (size_t)g.SY * g.M * sizeof(cplx)));
    CUDA_TRY(cudaMemsetAsync(c->snap[k], 0, (size_t)g.SY * g.M * sizeof(cplx), c->stream));
  }
  CUDA_TRY(cudaMalloc(&c->src, 4 * (size_t)(g.ny + 1) * sizeof(double)));
  c->halo_elems = (size_t)3 * g.M * g.SY * NG;
  CUDA_TRY(cudaMalloc(&c->sbuf_l, 3 * c->halo_elems * sizeof(cplx)));   // x2: the merged J exchange, x3: the window shift
  CUDA_TRY(cudaMalloc(&c->sbuf_r, 3 * c->halo_elems * sizeof(cplx)));   // x2: the merged J exchange, x3: the window shift
  CUDA_TRY(cudaMalloc(&c->rbuf_l, 3 * c->halo_elems * sizeof(cplx)));   // x2: the merged J exchange, x3: the window shift
  CUDA_TRY(cudaMalloc(&c->rbuf_r, 3 * c->halo_elems * sizeof(cplx)));   // x2: the merged J exchange, x3: the window shift
  CUDA_TRY(cudaMalloc(&c->counters, 32 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemsetAsync(c->counters, 0, 32 * sizeof(unsigned long long), c->stream));
  CUDA_TRY(cudaMallocHost(&c->h_counters, 32 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMalloc(&c->d_energy, 2 * sizeof(double)));
  CUDA_TRY(cudaMalloc(&c->n_dev, (CYLGPU_MAX_SPECIES + PST_N) * sizeof(int64_t)));
  CUDA_TRY(cudaMemsetAsync(c->n_dev, 0, (CYLGPU_MAX_SPECIES + PST_N) * sizeof(int64_t), c->stream));
  CUDA_TRY(cudaMalloc(&c->d_plan, sizeof(CompactPlan)));
  CUDA_TRY(cudaMallocHost(&c->h_pub, 8 * (CYLGPU_MAX_SPECIES + PST_N) * sizeof(int64_t)));
  if (build_tables(c)) return 1;
  set_neighbours(c);
  c->tr = make_transport(c);
  if (!c->tr) return 6;
  // peer-memory mailboxes for the neighbour exchanges (collective over the neighbours; NCCL stays as the fallback)
  c->p2p_cap_bytes = 3 * c->halo_elems * sizeof(cplx);
  if (int rc = p2p_setup(c, c->p2p_cap_bytes)) return rc;
  // several ranks on one host: host syncs yield the core (CYLGPU_BLOCKING_WAIT=0/1 overrides)
  c->blocking_wait = c->cfg.nranks > 1 && c->cfg.transport != CYLGPU_TRANSPORT_FABRIC;
  if (const char* e = getenv("CYLGPU_BLOCKING_WAIT")) c->blocking_wait = atoi(e) != 0;
  if (const char* e = getenv("CYLGPU_GRAPHS")) c->use_graphs = atoi(e) != 0;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}


int cylgpu_destroy(cylgpu_handle c);

int cylgpu_create(const cylgpu_config* cfg, cylgpu_handle* out) {
  if (!cfg || !out) { set_error("null argument"); return 1; }
  *out = nullptr;
  if (cfg->nx < 2 * NG || cfg->ny < 2 * NG) { set_error("nx and ny must be >= %d", 2 * NG); return 2; }
  if (cfg->n_mode < 1 || cfg->n_mode > 6) { set_error("n_mode must be 1..6 (the push kernels are instantiated for these)"); return 2; }
  if (cfg->n_species < 0 || cfg->n_species > CYLGPU_MAX_SPECIES) { set_error("n_species out of range"); return 2; }
  if (cfg->rank < 0 || cfg->rank >= cfg->nranks) { set_error("bad rank/nranks"); return 2; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available (%s); this library has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return 10;
  }
  cylgpu_ctx* c = new cylgpu_ctx();
  const int rc = create_impl(cfg, c);
  if (rc != 0) {   // nothing of a half-built handle survives (device allocations, streams, communicator)
    char msg[1024];
    snprintf(msg, sizeof(msg), "%s", cylgpu_last_error());
    cylgpu_destroy(c);
    set_error("%s", msg);
    return rc;
  }
  *out = c;
  return 0;
}

int cylgpu_destroy(cylgpu_handle c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  // the captured field phases hold NCCL send/recv nodes: they go before the communicator
  for (int k = 0; k < 4; ++k) if (c->graphs[k].exec) { cudaGraphExecDestroy(c->graphs[k].exec); c->graphs[k].exec = nullptr; }
  destroy_transport(c->tr);
  if (c->driver) { cylgpu_driver_release(c->driver); c->driver = nullptr; }
  for (int k = 0; k < CYLGPU_NFIELDS; ++k) cudaFree(c->f[k]);
  cudaFree(c->spare);
  for (int k = 0; k < CYLGPU_NSNAPS; ++k) cudaFree(c->snap[k]);
  cudaFree(c->tables); cudaFree(c->src);
  for (int s = 0; s < 2; ++s) for (int k = 0; k < 3; ++k) cudaFree(c->smooth_wk[s][k]);
  cudaFree(c->sbuf_l); cudaFree(c->sbuf_r); cudaFree(c->rbuf_l); cudaFree(c->rbuf_r);
  for (int i = 0; i < CYLGPU_MAX_SPECIES; ++i)
{
    for (int q = 0; q < 7; ++q) cudaFree(c->species[i].d[q]);
    for (int q = 0; q < 7; ++q) cudaFree(c->species[i].alt[q]);
    cudaFree(c->species[i].cell_start);
    cudaFree(c->species[i].perm);
  }
  cudaFree(c->ptmp); cudaFree(c->perm); cudaFree(c->flag); cudaFree(c->tailmark); cudaFree(c->hole_list);
  cudaFree(c->lowhole); cudaFree(c->hightail); cudaFree(c->scan_blocks);
  cudaFree(c->psend_l); cudaFree(c->psend_r); cudaFree(c->precv);
  cudaFree(c->counters); cudaFreeHost(c->h_counters); cudaFree(c->d_energy);
  c->timers.destroy();
  if (c->ev_wait) cudaEventDestroy(c->ev_wait);
  if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->ev_pfork) cudaEventDestroy(c->ev_pfork);
  if (c->ev_pdone) cudaEventDestroy(c->ev_pdone);
  cudaFree(c->n_dev); cudaFree(c->d_plan);
  if (c->h_pub) cudaFreeHost(c->h_pub);
  for (int k = 0; k < 8; ++k) if (c->pub[k].ev) cudaEventDestroy(c->pub[k].ev);
  for (int k = 0; k < 4; ++k) {
    if (c->app_pin[k]) cudaFreeHost(c->app_pin[k]);
    cudaFree(c->app_dev[k]);
    if (c->app_ev[k]) cudaEventDestroy(c->app_ev[k]);
  }
  if (c->ins_pin) cudaFreeHost(c->ins_pin);
  cudaFree(c->ins_dev);
  if (c->ins_ev) cudaEventDestroy(c->ins_ev);
  if (c->src_stage) { cudaFreeHost(c->src_stage); for (int k = 0; k < 8; ++k) cudaEventDestroy(c->src_event[k]); }
  for (int k = 0; k < 3; ++k) {
    cudaFree(c->hs.in[k]);
    if (c->hs.ev_up[k]) cudaEventDestroy(c->hs.ev_up[k]);
    if (c->hs.ev_unpacked[k]) cudaEventDestroy(c->hs.ev_unpacked[k]);
  }
  for (int k = 0; k < 2; ++k) {
    cudaFree(c->hs.out[k]);
    if (c->hs.ev_packed[k]) cudaEventDestroy(c->hs.ev_packed[k]);
    if (c->hs.ev_down[k]) cudaEventDestroy(c->hs.ev_down[k]);
  }
  if (c->hs.up) cudaStreamDestroy(c->hs.up);
  if (c->hs.down) cudaStreamDestroy(c->hs.down);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

int cylgpu_set_species(cylgpu_handle c, int isp, const cylgpu_species* sp) {
  TRY(check_handle(c));
  if (isp < 0 || isp >= c->cfg.n_species || !sp) { set_error("bad species index"); return 2; }
  c->species[isp].sp = *sp;
  c->species[isp].set = true;
  c->graph_epoch += 1;
  const int l0 = c->left, r0 = c->right;
  set_neighbours(c);
  // a periodic particle boundary closes the chain of slabs into a ring (mpi_routines.F90:186-199): new links,
  // new mailboxes (every rank registers the same species: collective)
  if ((c->left != l0 || c->right != r0) && c->p2p_cap_bytes) TRY(p2p_setup(c, c->p2p_cap_bytes));
  return 0;
}

int cylgpu_set_dt(cylgpu_handle c, double dt) {
  TRY(check_handle_fields(c));
  c->dt = dt;
  c->graph_epoch += 1;
  return 0;
}

int cylgpu_set_bc_field(cylgpu_handle c, const int32_t bc[4]) {
  TRY(check_handle(c));
  for (int i = 0; i < 4; ++i) c->bc_field[i] = bc[i];
  c->graph_epoch += 1;
  // the Cartesian communicator is created once (mpi_routines.F90:179-227); neighbours stay
  return 0;
}

int cylgpu_set_stream(cylgpu_handle c, void* stream) {
  TRY(check_handle(c));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->graph_epoch += 1;
  if (c->own_stream) { cudaStreamDestroy(c->stream); c->own_stream = false; }
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  return 0;
}

int cylgpu_synchronize(cylgpu_handle c) {
  TRY(check_handle(c));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return p2p_check(c);
}

static int field_ok(cylgpu_handle c, int id, int n) {
  if (id < 0 || id >= n) { set_error("field/snapshot id %d out of range", id); return 2; }
  (void)c;
  return 0;
}

int cylgpu_upload_field(cylgpu_handle c, int id, const void* host) {
  TRY(check_handle(c)); TRY(field_ok(c, id, CYLGPU_NFIELDS));
  CUDA_TRY(cudaMemcpyAsync(c->f[id], host, c->g.plane * c->g.M * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}
int cylgpu_download_field(cylgpu_handle c, int id, void* host) {
  TRY(check_handle(c)); TRY(field_ok(c, id, CYLGPU_NFIELDS));
  CUDA_TRY(cudaMemcpyAsync(host, c->f[id], c->g.plane * c->g.M * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}
int cylgpu_upload_snapshot(cylgpu_handle c, int id, const void* host) {
  TRY(check_handle(c)); TRY(field_ok(c, id, CYLGPU_NSNAPS));
  CUDA_TRY(cudaMemcpyAsync(c->snap[id], host, (size_t)c->g.SY * c->g.M * sizeof(cplx), cudaMemcpyHostToDevice,
                           c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}
int cylgpu_download_snapshot(cylgpu_handle c, int id, void* host) {
  TRY(check_handle(c)); TRY(field_ok(c, id, CYLGPU_NSNAPS));
  CUDA_TRY(cudaMemcpyAsync(host, c->snap[id], (size_t)c->g.SY * c->g.M * sizeof(cplx), cudaMemcpyDeviceToHost,
                           c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}
void* cylgpu_field_device_ptr(cylgpu_handle c, int id) {
  if (!c || id < 0 || id >= CYLGPU_NFIELDS) return nullptr;
  return c->f[id];
}
int cylgpu_snapshot_field_boundaries(cylgpu_handle c) {
  TRY(check_handle(c));
  return do_snapshot(c);
}

// ---- particles ----
__global__ void __launch_bounds__(256) k_aos_to_soa(const double* __restrict__ aos, double* d0, double* d1,
                                                    double* d2, double* d3, double* d4, double* d5, double* d6,
                                                    int64_t base, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* p = aos + 7 * i;
  d0[base + i] = p[0]; d1[base + i] = p[1]; d2[base + i] = p[2];
  d3[base + i] = p[3]; d4[base + i] = p[4]; d5[base + i] = p[5]; d6[base + i] = p[6];
}
__global__ void __launch_bounds__(256) k_soa_to_aos(double* __restrict__ aos, const double* d0, const double* d1,
                                                    const double* d2, const double* d3, const double* d4,
                                                    const double* d5, const double* d6, int64_t base, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double* p = aos + 7 * i;
  p[0] = d0[base + i]; p[1] = d1[base + i]; p[2] = d2[base + i];
  p[3] = d3[base + i]; p[4] = d4[base + i]; p[5] = d5[base + i]; p[6] = d6[base + i];
}

static const int64_t STAGE_PARTS = 1 << 22;   // staging chunk: 4 Mi particles = 224 MiB

static int append_impl(cylgpu_handle c, int isp, int64_t n, const double* host_aos) {
  if (isp < 0 || isp >= c->cfg.n_species) { set_error("bad species index"); return 2; }
  if (n < 0) { set_error("negative particle count"); return 2; }
  SpeciesState& S = c->species[isp];
  if (n == 0) return 0;
  TRY(reserve_particles(c, isp, S.n + n));
  double* stage = nullptr;
  const int64_t chunk = n < STAGE_PARTS ? n : STAGE_PARTS;
  CUDA_TRY(cudaMalloc(&stage, (size_t)chunk * 7 * sizeof(double)));
  for (int64_t off = 0; off < n; off += chunk) {
    const int64_t m = (n - off < chunk) ? n - off : chunk;
    CUDA_TRY(cudaMemcpyAsync(stage, host_aos + 7 * off, (size_t)m * 7 * sizeof(double), cudaMemcpyHostToDevice,
                             c->stream));
    k_aos_to_soa<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(stage, S.d[0], S.d[1], S.d[2], S.d[3], S.d[4],
                                                                    S.d[5], S.d[6], S.n + off, m);
    c->stats.kernel_launches += 1;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  CUDA_TRY(cudaFree(stage));
  S.n += n;
  c->stats.n_particles[isp] = S.n;
  c->sorted_valid = false;
  return set_count_exact(c, isp);
}

}  // extern "C"

namespace cylgpu {
__global__ void k_add_count_api(int64_t* n_dev, long long add) { *n_dev += add; }
__global__ void __launch_bounds__(256) k_aos_to_soa_dev(const double* __restrict__ aos, double* d0, double* d1,
                                                        double* d2, double* d3, double* d4, double* d5, double* d6,
                                                        const int64_t* __restrict__ base_dev, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t base = *base_dev;
  const double* p = aos + 7 * i;
  d0[base + i] = p[0]; d1[base + i] = p[1]; d2[base + i] = p[2];
  d3[base + i] = p[3]; d4[base + i] = p[4]; d5[base + i] = p[5]; d6[base + i] = p[6];
}

// A freshly generated plasma column (window.F90:157-300) joins the list without a host sync: the records go
// through a ring of pinned staging buffers (the caller's array is free again on return), the kernel appends
// them behind the device-side count.  With host-side counts (xcap = 0) the same, at the host's count.
int append_async(cylgpu_ctx* c, int isp, int64_t n, const double* host_aos) {
  if (isp < 0 || isp >= c->cfg.n_species) { set_error("bad species index"); return 2; }
  if (n <= 0) return 0;
  TRY(presort_join(c, false));
  SpeciesState& S = c->species[isp];
  TRY(reserve_particles(c, isp, S.n + n));
  const int slot = c->app_slot;
  c->app_slot = (c->app_slot + 1) & 3;
  if (c->app_ev[slot]) CUDA_TRY(cudaEventSynchronize(c->app_ev[slot]));   // long done: 4 columns ago
  else CUDA_TRY(cudaEventCreateWithFlags(&c->app_ev[slot], cudaEventDisableTiming));
  if (c->app_cap[slot] < n) {
    if (c->app_pin[slot]) cudaFreeHost(c->app_pin[slot]);
    if (c->app_dev[slot]) cudaFree(c->app_dev[slot]);
    const int64_t cap = n + n / 4 + 256;
    CUDA_TRY(cudaMallocHost(&c->app_pin[slot], (size_t)cap * 7 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&c->app_dev[slot], (size_t)cap * 7 * sizeof(double)));
    c->app_cap[slot] = cap;
  }
  memcpy(c->app_pin[slot], host_aos, (size_t)n * 7 * sizeof(double));
  CUDA_TRY(cudaMemcpyAsync(c->app_dev[slot], c->app_pin[slot], (size_t)n * 7 * sizeof(double), cudaMemcpyHostToDevice,
                           c->stream));
  CUDA_TRY(cudaEventRecord(c->app_ev[slot], c->stream));
  if (S.lazy) {
    k_aos_to_soa_dev<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->app_dev[slot], S.d[0], S.d[1], S.d[2], S.d[3],
                                                                        S.d[4], S.d[5], S.d[6], c->n_dev + isp, n);
    k_add_count_api<<<1, 1, 0, c->stream>>>(c->n_dev + isp, (long long)n);
    c->stats.kernel_launches += 2;
    S.n += n;   // the bound moves with the count
  } else {
    k_aos_to_soa<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->app_dev[slot], S.d[0], S.d[1], S.d[2], S.d[3],
                                                                    S.d[4], S.d[5], S.d[6], S.n, n);
    c->stats.kernel_launches += 1;
    S.n += n;
    TRY(set_count_exact(c, isp));
  }
  CUDA_TRY(cudaGetLastError());
  c->stats.n_particles[isp] = S.n;
  c->sorted_valid = false;
  return 0;
}
}  // namespace cylgpu

extern "C" {

int cylgpu_upload_particles(cylgpu_handle c, int isp, int64_t n, const double* host_aos) {
  TRY(check_handle(c));
  if (isp < 0 || isp >= c->cfg.n_species) { set_error("bad species index"); return 2; }
  c->species[isp].n = 0;
  c->stats.n_particles[isp] = 0;
  return append_impl(c, isp, n, host_aos);
}
int cylgpu_append_particles(cylgpu_handle c, int isp, int64_t n, const double* host_aos) {
  TRY(check_handle(c));
  return append_impl(c, isp, n, host_aos);
}

int cylgpu_download_particles(cylgpu_handle c, int isp, int64_t capacity, double* host_aos, int64_t* n_out) {
  TRY(check_handle(c));
  if (isp < 0 || isp >= c->cfg.n_species) { set_error("bad species index"); return 2; }
  SpeciesState& S = c->species[isp];
  if (n_out) *n_out = S.n;
  if (capacity < S.n) { set_error("download_particles: capacity %lld < count %lld", (long long)capacity, (long long)S.n); return 2; }
  if (S.n == 0) return 0;
  double* stage = nullptr;
  const int64_t chunk = S.n < STAGE_PARTS ? S.n : STAGE_PARTS;
  CUDA_TRY(cudaMalloc(&stage, (size_t)chunk * 7 * sizeof(double)));
  for (int64_t off = 0; off < S.n; off += chunk) {
    const int64_t m = (S.n - off < chunk) ? S.n - off : chunk;
    k_soa_to_aos<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(stage, S.d[0], S.d[1], S.d[2], S.d[3], S.d[4],
                                                                    S.d[5], S.d[6], off, m);
    c->stats.kernel_launches += 1;
    CUDA_TRY(cudaMemcpyAsync(host_aos + 7 * off, stage, (size_t)m * 7 * sizeof(double), cudaMemcpyDeviceToHost,
                             c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  CUDA_TRY(cudaFree(stage));
  return 0;
}

int cylgpu_particle_count(cylgpu_handle c, int isp, int64_t* n_out) {
  if (!c || isp < 0 || isp >= c->cfg.n_species || !n_out) { set_error("bad argument"); return 2; }
  TRY(flush_pending_remove(c));
  TRY(poll_counts(c, true));   // device-resident counts: wait for the newest copy
  *n_out = c->species[isp].n;
  return 0;
}

int cylgpu_particle_cells(cylgpu_handle c, int isp, int64_t capacity, int32_t* cells_out) {
  TRY(check_handle(c));
  if (isp < 0 || isp >= c->cfg.n_species) { set_error("bad species index"); return 2; }
  return do_cells(c, isp, capacity, cells_out);
}

void* cylgpu_particle_device_ptr(cylgpu_handle c, int isp, int comp) {
  if (!c || isp < 0 || isp >= c->cfg.n_species || comp < 0 || comp > 6) return nullptr;
  return c->species[isp].d[comp];
}

// ---- hot path ----
int cylgpu_update_e_field(cylgpu_handle c) { TRY(check_handle_fields(c)); return launch_update_e(c); }
int cylgpu_update_b_field(cylgpu_handle c) { TRY(check_handle_fields(c)); return launch_update_b(c); }
int cylgpu_efield_bcs(cylgpu_handle c) { TRY(check_handle_fields(c)); return do_efield_bcs(c); }
int cylgpu_bfield_bcs(cylgpu_handle c, int mpi_only) { TRY(check_handle_fields(c)); return do_bfield_bcs(c, mpi_only != 0); }
int cylgpu_bfield_final_bcs(cylgpu_handle c, const double* a, const double* b, const double* d, const double* e) {
  TRY(check_handle_fields(c));
  return do_bfield_final_bcs(c, a, b, d, e);
}
int cylgpu_particle_bcs(cylgpu_handle c) { TRY(check_handle_fields(c)); return do_particle_bcs(c); }
int cylgpu_push_no_bcs(cylgpu_handle c) { TRY(check_handle(c)); return do_push(c); }
int cylgpu_current_bcs(cylgpu_handle c) { TRY(check_handle_fields(c)); return do_current_bcs(c); }
int cylgpu_sort_particles(cylgpu_handle c) { TRY(check_handle(c)); return do_sort(c); }
int cylgpu_set_taylor_switch(cylgpu_handle c, double v) { TRY(check_handle(c)); c->taylor_switch = v; return 0; }
int cylgpu_set_pusher(cylgpu_handle c, int higuera_cary) { TRY(check_handle(c)); c->hc_push = higuera_cary != 0; return 0; }
int cylgpu_set_sort_interval(cylgpu_handle c, int n) { TRY(check_handle(c)); c->sort_interval = n; return 0; }
int cylgpu_set_push_variant(cylgpu_handle c, int v) {
  TRY(check_handle(c));
  if (v < 0 || v > 4) { set_error("push variant must be 0..4"); return 2; }
#if CYL_SHAPE != 0
  if (v != 4) { set_error("this build (particle shape %d) only has the generic push kernel (variant 4)", CYL_SHAPE); return 2; }
#endif
  c->push_variant = v;
  return 0;
}
int cylgpu_set_reference_quirks(cylgpu_handle c, int on) {
  TRY(check_handle(c));
  c->reference_quirks = on != 0;
  c->graph_epoch += 1;   // (kernel arguments of the captured field phases)
  return 0;
}
// the particle shape this library was built for (0 triangle, 1 top-hat, 2 third-order B-spline) and its ng
int cylgpu_shape(void) { return CYL_SHAPE; }
int cylgpu_ghost_cells(void) { return NG; }

// Run one field phase: directly, or -- once its parameters have been the same for three calls and
// the transport can be captured (none / NCCL) -- as a replayed CUDA graph, so that the ~15 short
// dependent launches of the phase cost one launch and do not depend on the host keeping ahead.
static int run_field_phase(cylgpu_ctx* c, int which, const std::function<int()>& body) {
  cylgpu_ctx::PhaseGraph& G = c->graphs[which];
  const bool capturable = c->use_graphs && !G.failed &&
                          (c->cfg.transport == CYLGPU_TRANSPORT_NONE || c->cfg.transport == CYLGPU_TRANSPORT_NCCL);
  if (!capturable) return body();
  if (G.seen_epoch == c->graph_epoch) G.stable_calls += 1;
  else { G.seen_epoch = c->graph_epoch; G.stable_calls = 0; }
  if (G.exec && G.epoch == c->graph_epoch) {
    CUDA_TRY(cudaGraphLaunch(G.exec, c->stream));
    c->stats.kernel_launches += G.launches;
    return 0;
  }
  if (G.stable_calls < 3) return body();
  // capture
  if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
  const bool timing = c->timing;
  const int64_t launches0 = c->stats.kernel_launches;
  c->timing = false;   // events recorded inside a capture cannot be read back
  cudaGraph_t graph = nullptr;
  int rc = 0;
  if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    c->timing = timing;
    G.failed = true;
    return body();
  }
  rc = body();
  const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
  c->timing = timing;
  G.launches = c->stats.kernel_launches - launches0;
  c->stats.kernel_launches = launches0;
  if (rc != 0 || ce != cudaSuccess || !graph ||
      cudaGraphInstantiate(&G.exec, graph, nullptr, nullptr, 0) != cudaSuccess) {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    G.exec = nullptr;
    G.failed = true;   // this configuration cannot be captured: stay on direct launches
    return body();
  }
  cudaGraphDestroy(graph);
  G.epoch = c->graph_epoch;
  CUDA_TRY(cudaGraphLaunch(G.exec, c->stream));
  c->stats.kernel_launches += G.launches;
  return 0;
}

static int fields_half_body(cylgpu_ctx* c) {
  if (wide_fields(c)) {   // field_ranges.cuh: ghost columns advanced here, ONE closing exchange for E and B
    const FieldRanges R = wide_ranges(c, 0);
    TRY(launch_update_e(c, R.e_lo, R.e_hi));
    TRY(efield_edges(c));
    TRY(launch_update_b(c, true, R.b_lo, R.b_hi));
    return halo_eb(c);
  }
  TRY(launch_update_e(c));
  TRY(do_efield_bcs(c));
  // bxm_old = bxm etc. (fields.f90:326-328) ride on the B sweep
  TRY(launch_update_b(c, true));
  return do_bfield_bcs(c, true);
}

// fields.f90:316-337
int cylgpu_fields_half(cylgpu_handle c) {
  TRY(check_handle_fields(c));
  TRY(presort_fork(c));   // the next push's cell sort, on the side stream behind this phase
  PhaseTimer t(c, &c->stats.ms_fields);
  return run_field_phase(c, 0, [c]() { return fields_half_body(c); });
}

// particles.F90:28-734 (push + r_min fold + particle_bcs)
int cylgpu_push(cylgpu_handle c) {
  TRY(check_handle_fields(c));
  PhaseTimer t(c, &c->stats.ms_push);
  return do_push_bcs(c);
}

// the same for particle lists that stay in host memory (streamed through the GPU in chunks)
int cylgpu_push_host(cylgpu_handle c, const int64_t* n_in, double* const* host_aos, const int64_t* capacity,
                     int64_t* n_out) {
  TRY(check_handle(c));
  if (!n_in || !host_aos || !capacity || !n_out) { set_error("push_host: null argument"); return 2; }
  PhaseTimer t(c, &c->stats.ms_push);
  return do_push_host(c, n_in, host_aos, capacity, n_out);
}
int cylgpu_set_host_chunk(cylgpu_handle c, int64_t particles) {
  TRY(check_handle(c));
  if (particles < 1024 || particles > (int64_t)1 << 30) { set_error("host chunk must be in [1024, 2^30] particles"); return 2; }
  c->host_chunk = particles;
  return 0;
}

// smooth_currents / smooth_its / smooth_compensation / smooth_strides of the control block
// (deck_control_block.F90:447-466, shared_data.F90:468-472)
int cylgpu_set_current_smoothing(cylgpu_handle c, int enable, int its, int comp_its, int nstrides,
                                 const int32_t* strides) {
  TRY(check_handle(c));
  if (its < 0 || comp_its < 0 || nstrides < 0 || (nstrides > 0 && !strides)) { set_error("bad smoothing settings"); return 2; }
  for (int k = 0; k < nstrides; ++k)
    if (strides[k] < 1 || strides[k] > NG) { set_error("smooth_strides must lie in 1..%d", NG); return 2; }
  c->graph_epoch += 1;
  c->smooth_currents = enable != 0;
  c->smooth_its = its;
  c->smooth_comp_its = comp_its;
  c->smooth_strides.assign(strides, strides + nstrides);
  return 0;
}

int cylgpu_current_finish(cylgpu_handle c) {
  TRY(check_handle_fields(c));
  PhaseTimer t(c, &c->stats.ms_bcs);
  return run_field_phase(c, 2, [c]() { return do_current_finish(c); });
}

// fields.f90:341-353
int cylgpu_fields_final(cylgpu_handle c, const double* s1min, const double* s2min, const double* s1max,
                        const double* s2max) {
  TRY(check_handle_fields(c));
  PhaseTimer t(c, &c->stats.ms_fields);
  TRY(upload_laser_sources(c, s1min, s2min, s1max, s2max));
  // final_shift_follows (driver.cu): shift_fields comes next and ends with its own exchange of all nine arrays;
  // this phase then advances column nx+1 -- the one the shift moves into the interior -- itself and leaves the
  // ghosts to the window's halo: no exchange here at all
  const bool to_shift = c->final_shift_follows && wide_fields(c);
  return run_field_phase(c, to_shift ? 3 : 1, [c, to_shift]() -> int {
    if (wide_fields(c)) {
      const FieldRanges R = wide_ranges(c, to_shift ? 2 : 1);
      TRY(launch_update_b(c, false, R.b_lo, R.b_hi));
      TRY(do_bfield_final_bcs_device(c, true, to_shift ? 2 : 1));
      TRY(launch_update_e(c, R.e_lo, R.e_hi));
      TRY(efield_edges(c));
      return to_shift ? 0 : halo_eb(c);
    }
    TRY(launch_update_b(c));
    TRY(do_bfield_final_bcs_device(c));
    TRY(launch_update_e(c));
    return do_efield_bcs(c);
  });
}

// window.F90:62-94, one cell.  grid5 = {x_grid_min_local, x_min, x_max, x_min_local,
// x_max_local} AFTER the host's setup_grid_x for the shifted window.
int cylgpu_window_shift(cylgpu_handle c, const int64_t* n_new, const double* const* new_aos, const double* grid5) {
  TRY(check_handle_fields(c));
  if (!grid5) { set_error("window_shift needs the shifted grid"); return 2; }
  if (n_new && new_aos) {
    for (int isp = 0; isp < c->cfg.n_species; ++isp)
      if (n_new[isp] > 0) {
        TRY(append_async(c, isp, n_new[isp], new_aos[isp]));
        c->r_clean = false;   // a caller-supplied column: the next particle_bcs tests every rule
      }
  }
  c->x_grid_min_local = grid5[0];
  c->x_min = grid5[1]; c->x_max = grid5[2];
  c->x_min_local = grid5[3]; c->x_max_local = grid5[4];
  TRY(do_remove_behind(c));
  if (c->cfg.x_min_boundary) { c->host_remove_active = true; c->host_remove_x = c->x_min; }   // (host-resident lists)
  if (c->xcap > 0) TRY(publish_counts(c));
  return do_shift_fields(c);
}

// ---- moving-window plasma column with the reference's random stream ----
int cylgpu_rng_init(cylgpu_handle c, int seed) {
  TRY(check_handle_fields(c));   // host-side generator state only
  kiss_init(c->rng, seed);
  return 0;
}
int cylgpu_rng_set_state(cylgpu_handle c, const int32_t* xyzw, int cached, double cached_value) {
  TRY(check_handle_fields(c));   // host-side generator state only
  if (!xyzw) { set_error("rng_set_state: null state"); return 2; }
  c->rng.x = (uint32_t)xyzw[0]; c->rng.y = (uint32_t)xyzw[1]; c->rng.z = (uint32_t)xyzw[2]; c->rng.w = (uint32_t)xyzw[3];
  c->rng.cached = cached != 0;
  c->rng.cached_value = cached_value;
  return 0;
}
int cylgpu_rng_get_state(cylgpu_handle c, int32_t* xyzw, int* cached, double* cached_value) {
  TRY(check_handle_fields(c));   // host-side generator state only
  if (!xyzw || !cached || !cached_value) { set_error("rng_get_state: null argument"); return 2; }
  xyzw[0] = (int32_t)c->rng.x; xyzw[1] = (int32_t)c->rng.y; xyzw[2] = (int32_t)c->rng.z; xyzw[3] = (int32_t)c->rng.w;
  *cached = c->rng.cached;
  *cached_value = c->rng.cached_value;
  return 0;
}
int cylgpu_rng_flush_cache(cylgpu_handle c) {   // random_flush_cache, called by output_routines every step
  TRY(check_handle_fields(c));   // host-side generator state only: a deferred particle_bcs stays outstanding
  c->rng.cached = 0;
  return 0;
}
int cylgpu_rng_uniform(cylgpu_handle c, double* out) {
  TRY(check_handle_fields(c));   // host-side generator state only
  if (!out) { set_error("rng_uniform: null argument"); return 2; }
  *out = kiss_uniform(c->rng);
  return 0;
}
int cylgpu_insert_particles(cylgpu_handle c, int isp, double x_grid_max, double npart_per_cell, const double* density,
                            const double* temperature, const double* drift, double dmin, double dmax,
                            int64_t* n_inserted) {
  TRY(check_handle_fields(c));
  if (isp < 0 || isp >= c->cfg.n_species || !c->species[isp].set) { set_error("bad species index"); return 2; }
  if (!density || !temperature || !drift) { set_error("insert_particles: null profile"); return 2; }
  std::vector<double> aos;
  TRY(do_insert_particles(c, isp, x_grid_max, npart_per_cell, density, temperature, drift, dmin, dmax, aos));
  const int64_t n = (int64_t)(aos.size() / 7);
  if (n_inserted) *n_inserted = n;
  if (n > 0) TRY(append_async(c, isp, n, aos.data()));
  return 0;
}

// insert_particles for a species whose list lives in host memory (cylgpu_push_host): the column goes behind the
// *n_inout records of host_aos
int cylgpu_insert_particles_host(cylgpu_handle c, int isp, double x_grid_max, double npart_per_cell, const double* density,
                                 const double* temperature, const double* drift, double dmin, double dmax,
                                 double* host_aos, int64_t capacity, int64_t* n_inout) {
  TRY(check_handle_fields(c));
  if (isp < 0 || isp >= c->cfg.n_species || !c->species[isp].set) { set_error("bad species index"); return 2; }
  if (!density || !temperature || !drift || !host_aos || !n_inout) { set_error("insert_particles_host: null argument"); return 2; }
  std::vector<double> aos;
  TRY(do_insert_particles(c, isp, x_grid_max, npart_per_cell, density, temperature, drift, dmin, dmax, aos));
  const int64_t n = (int64_t)(aos.size() / 7);
  if (*n_inout + n > capacity) {
    set_error("insert_particles_host: species %d needs room for %lld particles, capacity %lld", isp,
              (long long)(*n_inout + n), (long long)capacity);
    return 2;
  }
  if (n > 0) memcpy(host_aos + 7 * *n_inout, aos.data(), (size_t)n * 7 * sizeof(double));
  *n_inout += n;
  return 0;
}

// ---- SDF dump / restart (sdf_io.cu) ----
int cylgpu_sdf_write_host(const char* path, const cylgpu_sdf_desc* d, const void* const* fields15,
                          const double* const* particles_aos, const double* const* derived) {
  return sdf_write_host(path, d, fields15, particles_aos, derived);
}
int cylgpu_sdf_derived_count(const cylgpu_sdf_desc* d) { return d ? sdf_derived_count(d) : 0; }
int cylgpu_sdf_read_host(const char* path, cylgpu_sdf_desc* d, void* const* fields15, double x_lo, double x_hi,
                         double* const* particles_aos, const int64_t* capacity) {
  std::vector<std::vector<double>> parts;
  TRY(sdf_read_host(path, d, fields15, x_lo, x_hi, &parts));
  if (!particles_aos) return 0;
  for (int s = 0; s < d->n_species; ++s) {
    const int64_t n = (int64_t)(parts[(size_t)s].size() / 7);
    if (n == 0) continue;
    if (!capacity || !particles_aos[s] || capacity[s] < n) {
      set_error("sdf_read_host: species %d needs room for %lld particles", s, (long long)n);
      return 2;
    }
    memcpy(particles_aos[s], parts[(size_t)s].data(), (size_t)n * 7 * sizeof(double));
  }
  return 0;
}
int cylgpu_sdf_dump(cylgpu_handle c, const char* path, cylgpu_sdf_desc* d) {
  TRY(check_handle(c));
  if (!d || !path) { set_error("sdf_dump: null argument"); return 2; }
  const Geom& g = c->g;
  if (d->nx_local != g.nx || d->ny_global != g.ny || d->n_mode != g.M || d->n_species != c->cfg.n_species) {
    set_error("sdf_dump: the descriptor does not match the handle");
    return 2;
  }
  const size_t nf = g.plane * g.M;
  std::vector<std::vector<cplx>> host((size_t)CYLGPU_NFIELDS, std::vector<cplx>(nf));
  const void* fptr[CYLGPU_NFIELDS];
  for (int k = 0; k < CYLGPU_NFIELDS; ++k) {
    CUDA_TRY(cudaMemcpyAsync(host[(size_t)k].data(), c->f[k], nf * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
    fptr[k] = host[(size_t)k].data();
  }
  std::vector<std::vector<double>> parts((size_t)d->n_species);
  const double* pptr[CYLGPU_MAX_SPECIES] = {0};
  for (int s = 0; s < d->n_species; ++s) {
    const int64_t n = c->species[s].n;
    d->npart_local[s] = n;
    parts[(size_t)s].resize((size_t)(7 * n));
    if (n > 0) TRY(cylgpu_download_particles(c, s, n, parts[(size_t)s].data(), nullptr));
    pptr[s] = parts[(size_t)s].data();
    if (c->cfg.nranks == 1) { d->npart_global[s] = n; d->npart_offset[s] = 0; }
    if (!d->have_extents) {
      double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
      for (int64_t i = 0; i < n; ++i)
        for (int q = 0; q < 3; ++q) {
          const double v = parts[(size_t)s][(size_t)(7 * i + q)];
          if (v < lo[q]) lo[q] = v;
          if (v > hi[q]) hi[q] = v;
        }
      for (int q = 0; q < 3; ++q) { d->part_extents[s][q] = n ? lo[q] : 0.0; d->part_extents[s][3 + q] = n ? hi[q] : 0.0; }
    }
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  // derived variables straight from the device lists (moments.cuh); table order of sdf_io.cu / cylgpu.h
  static const int kind_of[CYLGPU_SDF_NDERIVED] = {
      CYLGPU_MOM_EKBAR, CYLGPU_MOM_MASS_DENSITY, -1 /* charge density */, CYLGPU_MOM_NUMBER_DENSITY, CYLGPU_MOM_PPC,
      CYLGPU_MOM_AVERAGE_WEIGHT, CYLGPU_MOM_AVERAGE_MOMENTUM, CYLGPU_MOM_AVERAGE_MOMENTUM, CYLGPU_MOM_AVERAGE_MOMENTUM,
      CYLGPU_MOM_TEMPERATURE, CYLGPU_MOM_TEMPERATURE, CYLGPU_MOM_TEMPERATURE, CYLGPU_MOM_TEMPERATURE,
      CYLGPU_MOM_SPECIES_CURRENT, CYLGPU_MOM_SPECIES_CURRENT, CYLGPU_MOM_SPECIES_CURRENT,
      CYLGPU_MOM_EKFLUX, CYLGPU_MOM_EKFLUX, CYLGPU_MOM_EKFLUX, CYLGPU_MOM_EKFLUX, CYLGPU_MOM_EKFLUX, CYLGPU_MOM_EKFLUX};
  static const int dir_of[CYLGPU_SDF_NDERIVED] = {0, 0, 0, 0, 0, 0, 1, 2, 3, 0, 1, 2, 3, 1, 2, 3, 1, 2, 3, -1, -2, -3};
  const int nder = sdf_derived_count(d);
  const int nreal = nder - ((d->derived_mask & (1u << CYLGPU_SDF_NDERIVED)) ? d->n_species : 0);
  std::vector<std::vector<double>> der((size_t)nreal, std::vector<double>(g.plane));
  std::vector<const double*> dptr((size_t)nder);
  int q = 0;
  for (int v = 0; v < CYLGPU_SDF_NDERIVED; ++v) {
    if (!(d->derived_mask & (1u << v))) continue;
    for (int s = d->derived_sum ? -1 : 0; s < (d->derived_species ? d->n_species : 0); ++s, ++q) {
      if (kind_of[v] < 0) {
        TRY(do_number_density_modes(c, s, true));
        TRY(download_real_part_mode0(c, c->spare, der[(size_t)q].data()));
      } else {
        TRY(do_particle_moment(c, kind_of[v], s, dir_of[v], der[(size_t)q].data()));
      }
      dptr[(size_t)q] = der[(size_t)q].data();
    }
  }
  std::vector<std::vector<cplx>> dmode;
  if (d->derived_mask & (1u << CYLGPU_SDF_NDERIVED)) {   // calc_number_density_modes per species
    dmode.assign((size_t)d->n_species, std::vector<cplx>(nf));
    for (int s = 0; s < d->n_species; ++s, ++q) {
      TRY(do_number_density_modes(c, s, false));
      CUDA_TRY(cudaMemcpyAsync(dmode[(size_t)s].data(), c->spare, nf * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      dptr[(size_t)q] = reinterpret_cast<const double*>(dmode[(size_t)s].data());
    }
  }
  return sdf_write_host(path, d, fptr, pptr, nder ? dptr.data() : nullptr);
}
int cylgpu_sdf_load(cylgpu_handle c, const char* path, cylgpu_sdf_desc* d) {
  TRY(check_handle(c));
  if (!d || !path) { set_error("sdf_load: null argument"); return 2; }
  const Geom& g = c->g;
  if (d->nx_local != g.nx || d->ny_global != g.ny || d->n_mode != g.M || d->n_species != c->cfg.n_species) {
    set_error("sdf_load: the descriptor does not match the handle");
    return 2;
  }
  const size_t nf = g.plane * g.M;
  std::vector<std::vector<cplx>> host((size_t)CYLGPU_NFIELDS, std::vector<cplx>(nf, C(0.0, 0.0)));
  void* fptr[CYLGPU_NFIELDS];
  for (int k = 0; k < CYLGPU_NFIELDS; ++k) fptr[k] = host[(size_t)k].data();
  std::vector<std::vector<double>> parts;
  // the slab owns x_min_local <= x < x_max_local (boundary.F90:1607,1685)
  TRY(sdf_read_host(path, d, fptr, c->x_min_local, c->x_max_local, &parts));
  for (int k = 0; k < CYLGPU_NFIELDS; ++k)
    CUDA_TRY(cudaMemcpyAsync(c->f[k], host[(size_t)k].data(), nf * sizeof(cplx), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  for (int s = 0; s < d->n_species; ++s)
    TRY(cylgpu_upload_particles(c, s, (int64_t)(parts[(size_t)s].size() / 7), parts[(size_t)s].data()));
  c->graph_epoch++;
  return 0;
}

// insert_particles generated on the device from a counter-based stream (window_insert.cu)
int cylgpu_insert_particles_device(cylgpu_handle c, int isp, double x_grid_max, double npart_per_cell,
                                   const double* density, const double* temperature, const double* drift, double dmin,
                                   double dmax, uint64_t seed, uint64_t column, int64_t* n_inserted) {
  TRY(check_handle_fields(c));
  if (isp < 0 || isp >= c->cfg.n_species || !c->species[isp].set) { set_error("bad species index"); return 2; }
  if (!density || !temperature || !drift) { set_error("insert_particles_device: null profile"); return 2; }
  return do_insert_particles_device(c, isp, x_grid_max, npart_per_cell, density, temperature, drift, dmin, dmax, seed,
                                    column, n_inserted);
}
// Philox4x32-10 on the host (no device needed): lets the stream be checked against the published
// known-answer vectors and lets a host reproduce any particle of a device-generated column
int cylgpu_philox4x32(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) {
  if (!ctr4 || !key2 || !out4) { set_error("philox4x32: null argument"); return 2; }
  const Philox4 p = philox4x32_10(ctr4[0], ctr4[1], ctr4[2], ctr4[3], key2[0], key2[1]);
  for (int i = 0; i < 4; ++i) out4[i] = p.v[i];
  return 0;
}

// calc_number_density_modes (calc_df.F90:588-661) from the device-resident lists: no particle
// download on dump steps.  host_out: (nx+2ng, ny+2ng, n_mode) complex128, Fortran order.
int cylgpu_number_density_modes(cylgpu_handle c, int species, void* host_out) {
  TRY(check_handle(c));
  if (species >= c->cfg.n_species || !host_out) { set_error("number_density_modes: bad argument"); return 2; }
  TRY(do_number_density_modes(c, species, false));
  CUDA_TRY(cudaMemcpyAsync(host_out, c->spare, c->g.plane * c->g.M * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}

// the other calc_df.F90 moments (moments.cuh); host_out: real (nx+2ng, ny+2ng) array
int cylgpu_particle_moment(cylgpu_handle c, int kind, int species, int direction, double* host_out) {
  TRY(check_handle(c));
  if (species >= c->cfg.n_species || !host_out) { set_error("particle_moment: bad argument"); return 2; }
  return do_particle_moment(c, kind, species, direction, host_out);
}

// calc_charge_density (calc_df.F90:442-519); host_out: real (nx+2ng, ny+2ng) array
int cylgpu_charge_density(cylgpu_handle c, int species, double* host_out) {
  TRY(check_handle(c));
  if (species >= c->cfg.n_species || !host_out) { set_error("charge_density: bad argument"); return 2; }
  TRY(do_number_density_modes(c, species, true));
  return download_real_part_mode0(c, c->spare, host_out);
}

int cylgpu_energy(cylgpu_handle c, double* out2) { TRY(check_handle(c)); return do_energy(c, out2); }

int cylgpu_stats(cylgpu_handle c, cylgpu_stats_t* out) {
  if (!c || !out) { set_error("bad argument"); return 2; }
  TRY(flush_pending_remove(c));
  TRY(poll_counts(c, true));
  c->timers.drain();
  for (int i = 0; i < CYLGPU_MAX_SPECIES; ++i) c->stats.n_particles[i] = c->species[i].n;
  *out = c->stats;
  return 0;
}
int cylgpu_reset_stats(cylgpu_handle c) {
  if (!c) return 2;
  c->timers.drain();
  c->stats.kernel_launches = 0;
  c->stats.ms_fields = c->stats.ms_push = c->stats.ms_bcs = c->stats.ms_sort = c->stats.ms_exchange = 0.0;
  c->stats.ms_push_kernel = 0.0;
  c->stats.n_push_kernel = 0;
  return 0;
}
// Device-resident particle counts (ctx.cuh): capacity > 0 = particles per direction of the fixed-size migration
// message, and no host sync inside a step; 0 = the exact two-message protocol.  Every rank sets the same value.
int cylgpu_set_exchange_capacity(cylgpu_handle c, int64_t capacity) {
  TRY(check_handle(c));
  if (capacity < 0 || capacity > ((int64_t)1 << 28)) { set_error("exchange capacity out of range"); return 2; }
  c->xcap = capacity;
  for (int i = 0; i < c->cfg.n_species; ++i) TRY(set_count_exact(c, i));
  // the mailboxes must hold the particle message as well (every rank passes the same capacity: collective)
  const size_t need = std::max<size_t>(3 * c->halo_elems * sizeof(cplx), (size_t)(7 * capacity + 7) * sizeof(double));
  if (need > c->p2p_cap_bytes) { TRY(p2p_setup(c, need)); }
  c->p2p_cap_bytes = std::max(c->p2p_cap_bytes, need);
  return 0;
}
int cylgpu_transport_info(cylgpu_handle c, int32_t* out4) {
  if (!c || !out4) { set_error("bad argument"); return 2; }
  out4[0] = c->cfg.transport;
  out4[1] = c->p2p_link_l ? 1 : 0;
  out4[2] = c->p2p_link_r ? 1 : 0;
  out4[3] = (int32_t)(c->p2p_cap_bytes / 1024);
  return 0;
}
int cylgpu_set_timing(cylgpu_handle c, int on) {
  if (!c) return 2;
  c->timing = on != 0;
  return 0;
}

}  // extern "C"
