// balance.cu -- the slab re-balancer's arithmetic for a z-only (x-slab) decomposition: the per-column load of
// part_load_func / get_load (balance.F90:2322-2470) from the device-resident lists, and calculate_breaks
// (balance.F90:2510-2653), which places the slab boundaries.  The redistribution itself (redistribute_domain /
// distribute_particles, balance.F90:303-2300) moves whole columns and their particles between handles and lives in
// the host mirror (cylindrical_epoch_b200/balance.py): with nprocy = 1 a rank's new slab is a contiguous range of
// global columns.  Product code: never includes, links or calls anything under oracle/.
#include <climits>
#include <vector>

#include "ctx.cuh"

namespace cylgpu {

// part_load_func, balance.F90:2453-2478: one count per particle at cell_x = FLOOR((x - x_grid_min) / dx + 1.5);
// local index ix - (1 - ng) in [0, nx + 2 ng), clamped (a particle further out than the ghosts has left the slab)
__global__ void __launch_bounds__(256) k_load_x(const double* __restrict__ x, int64_t n, double x_grid_min_local,
                                                double idx, int nx, unsigned long long* __restrict__ load) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
#if CYL_SHAPE == 1   // balance.F90:2470-2473
  int cell = (int)floor((x[i] - x_grid_min_local) * idx) + 1;
#else
  int cell = (int)floor((x[i] - x_grid_min_local) * idx + 1.5);
#endif
  cell = max(1 - NG, min(nx + NG, cell));
  atomicAdd(&load[cell - (1 - NG)], 1ULL);
}

}  // namespace cylgpu

using namespace cylgpu;

extern "C" {

// load_out[ix + ng - 1], ix = 1-ng .. nx+ng: macro-particles of all species per column of this slab
int cylgpu_load_x(cylgpu_handle c, int64_t* load_out) {
  if (!c || !load_out) { set_error("load_x: null argument"); return 2; }
  CUDA_TRY(cudaSetDevice(c->device));
  TRY(presort_join(c, false));
  TRY(flush_pending_remove(c));
  TRY(poll_counts(c, true));
  const int n = c->g.nx + 2 * NG;
  unsigned long long* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, (size_t)n * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemsetAsync(d, 0, (size_t)n * sizeof(unsigned long long), c->stream));
  for (int isp = 0; isp < c->cfg.n_species; ++isp) {
    const SpeciesState& S = c->species[isp];
    if (!S.set || S.n == 0) continue;
    k_load_x<<<(unsigned)((S.n + 255) / 256), 256, 0, c->stream>>>(S.d[0], S.n, c->x_grid_min_local, 1.0 / c->cfg.dx,
                                                                  c->g.nx, d);
    c->stats.kernel_launches += 1;
  }
  CUDA_TRY(cudaMemcpyAsync(load_out, d, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaFree(d));
  return 0;
}

// calculate_breaks, balance.F90:2510-2653.  load[0 .. sz + 2 ng): the reference's load(1-ng : sz+ng), of which only
// 1..sz enter; mins / maxs: nproc entries each, 1-based inclusive cell ranges.  Host arithmetic only (no device).
int cylgpu_calculate_breaks(const int64_t* load_in, int32_t sz, int32_t nproc, int32_t* mins, int32_t* maxs) {
  if (!load_in || !mins || !maxs || sz < 1 || nproc < 1) { set_error("calculate_breaks: bad argument"); return 2; }
  const int ng = NG;
  const int ncell_min = (PNG + 1) / 2 + 1;   // constants.F90:548
  auto load = [&](int i) -> long long { return (long long)load_in[i - (1 - ng)]; };   // Fortran index
  for (int p = 0; p < nproc; ++p) { mins[p] = 1; maxs[p] = sz; }
  if (nproc < 2) return 0;
  if ((long long)nproc * ncell_min > sz) { set_error("calculate_breaks: %d cells cannot hold %d slabs", sz, nproc); return 2; }
  std::vector<int> mx((size_t)nproc + 1, sz);   // 1-based
  long long sum = 0;
  for (int i = 1; i <= sz; ++i) sum += load(i);
  const long long ideal = (long long)floor((double)sum / nproc + 0.5);
  int proc = 0, old = 1;
  long long total = 0;
  for (int idim = 1; idim <= sz; ++idim) {
    const long long total_old = total;
    total = total + load(idim);
    if (total >= ideal) {
      proc = proc + 1;
      if (ideal - total_old < total - ideal) mx[proc] = idim - 1;
      else mx[proc] = idim;
      const int nextra = old - mx[proc] + ncell_min;
      if (nextra > 0) mx[proc] = mx[proc] + nextra;
      if (proc == nproc - 1) break;
      old = mx[proc];
      total = total - ideal;
    }
  }
  auto sanity_backwards = [&]() {
    int o = sz;
    for (int p = nproc - 1; p >= 1; --p) {
      if (o - mx[p] < ncell_min) mx[p] = o - ncell_min;
      o = mx[p];
    }
  };
  sanity_backwards();
  auto spread = [&](long long& lmax, long long& lmin) {
    lmax = -1;
    lmin = LLONG_MAX;
    int i0 = 1;
    for (int p = 1; p <= nproc; ++p) {
      const int i1 = mx[p];
      long long l = 0;
      for (int i = i0; i <= i1; ++i) l += load(i);
      if (l > lmax) lmax = l;
      if (l < lmin) lmin = l;
      i0 = i1 + 1;
    }
  };
  // perturb the splits by one cell while that narrows the spread (balance.F90:2575-2630)
  long long best = LLONG_MAX, lmax = 0, lmin = 0;
  for (int iter = 1; iter <= 1000; ++iter) {
    bool left_early = false;
    for (int i = 1; i <= nproc - 1 && !left_early; ++i) {
      // minus
      int old_maxs = mx[i];
      int o = (i == 1) ? 0 : mx[i - 1];
      int new_maxs = old_maxs;
      if (old_maxs - o - 1 >= ng) new_maxs = old_maxs - 1;
      if (new_maxs != old_maxs) {
        mx[i] = new_maxs;
        spread(lmax, lmin);
        if (lmax - lmin < best) { left_early = true; break; }
        mx[i] = old_maxs;
      }
      // plus
      old_maxs = mx[i];
      o = mx[i + 1];
      new_maxs = old_maxs;
      if (o - old_maxs - 1 >= ng) new_maxs = old_maxs + 1;
      if (new_maxs != old_maxs) {
        mx[i] = new_maxs;
        spread(lmax, lmin);
        if (lmax - lmin < best) { left_early = true; break; }
        mx[i] = old_maxs;
      }
    }
    if (lmax - lmin < best) best = lmax - lmin;
    else break;
  }
  sanity_backwards();
  int o = 0;
  for (int p = 1; p <= nproc - 1; ++p) {   // forwards
    if (mx[p] - o < ncell_min) mx[p] = o + ncell_min;
    o = mx[p];
  }
  mins[0] = 1;
  for (int p = 1; p <= nproc; ++p) maxs[p - 1] = mx[p];
  maxs[nproc - 1] = sz;
  for (int p = 2; p <= nproc; ++p) mins[p - 1] = mx[p - 1] + 1;
  return 0;
}

}  // extern "C"
