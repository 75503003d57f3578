// window_insert.cu -- insert_particles of the moving window (window.F90:157-300) with the
// reference's own random stream, so that a moving-window run loads bit-identical plasma:
// the KISS generator and polar Box-Muller transform of random_generator.f90:45-173 (one stream
// per rank, seeded 7842432 + rank with 1000 warm-up draws, setup.F90:563-567; the spare
// Gaussian is dropped once per step by output_routines, diagnostics.F90:235).  The column is
// generated on the host, as in the reference, and appended to the device list.
// Product code: never includes, links or calls anything under oracle/.
#include <cmath>
#include <limits>
#include <vector>

#include "ctx.cuh"
#include "philox.cuh"

namespace cylgpu {

constexpr double KB = 1.3806488e-23;   // constants.F90:187

// random(), random_generator.f90:45-77: Fortran default INTEGER arithmetic wraps modulo 2^32 and
// ISHFT is a logical shift, hence unsigned words here and one signed reinterpretation at the end
double kiss_uniform(KissState& s) {
  s.x = 69069u * s.x + 1327217885u;
  uint32_t a = s.y;
  a ^= a << 13;
  a ^= a >> 17;
  a ^= a << 5;
  s.y = a;
  s.z = 18000u * (s.z & 65535u) + (s.z >> 16);
  s.w = 30903u * (s.w & 65535u) + (s.w >> 16);
  const uint32_t kiss = s.x + s.y + (s.z << 16) + s.w;
  return ((double)(int32_t)kiss + 2147483648.0) / 4294967296.0;
}

void kiss_init(KissState& s, int seed) {   // random_init, random_generator.f90:81-108
  s.x = (uint32_t)(123456789 + seed);
  s.y = (uint32_t)(362436069 + seed);
  s.z = (uint32_t)(521288629 + seed);
  s.w = (uint32_t)(916191069 + seed);
  s.cached = 0;
  s.cached_value = 0.0;
  for (int i = 0; i < 1000; ++i) (void)kiss_uniform(s);
}

double kiss_box_muller(KissState& s, double stdev, double mu) {   // random_generator.f90:112-173
  if (s.cached) {
    s.cached = 0;
    return s.cached_value * stdev + mu;
  }
  s.cached = 1;
  double r1, r2, w;
  do {
    r1 = 2.0 * kiss_uniform(s) - 1.0;
    r2 = 2.0 * kiss_uniform(s) - 1.0;
    w = r1 * r1 + r2 * r2;
  } while (!(w > std::numeric_limits<double>::min() && w < 1.0));
  w = std::sqrt((-2.0 * std::log(w)) / w);
  s.cached_value = r2 * w;
  return r1 * w * stdev + mu;
}

// window.F90:187-298 for one species.  density(0:ny+1), temperature(0:ny+1,1:3) and
// drift(0:ny+1,1:3) are the deck functions evaluated by the host on the column ix = nx
// (window.F90:203-220; Fortran order, radial index fastest).
int do_insert_particles(cylgpu_ctx* c, int isp, double x_grid_max, double npart_per_cell_real, const double* density_in,
                        const double* temperature, const double* drift, double dmin, double dmax,
                        std::vector<double>& aos) {
  aos.clear();
  if (!c->cfg.x_max_boundary) return 0;   // only the rightmost rank injects
  const SpeciesState& S = c->species[isp];
  const int ny = c->g.ny, nrow = ny + 2;
  const double dx = c->cfg.dx, dy = c->cfg.dy;
  std::vector<double> density(density_in, density_in + nrow);
  for (int iy = 0; iy < nrow; ++iy) {
    if (density[iy] > dmax) density[iy] = dmax;
    if (density[iy] < dmin) density[iy] = 0.0;
  }
  const int64_t npart_per_cell = (int64_t)std::floor(npart_per_cell_real);
  const double npart_frac = npart_per_cell_real - (double)npart_per_cell;
  const double x0 = x_grid_max + 0.5 * dx;
  KissState& rng = c->rng;
  for (int iy = 1; iy <= ny; ++iy) {
    if (density[iy] < dmin) continue;
    int64_t n_frac = 0;
    if (npart_frac > 0.0 && kiss_uniform(rng) < npart_frac) n_frac = 1;
    const int64_t ncell = npart_per_cell + n_frac;
    const double y_iy = c->cfg.y_grid_min_local + (double)(iy - 1) * dy;
    for (int64_t ip = 0; ip < ncell; ++ip) {
      const double cell_frac_y = 0.5 - kiss_uniform(rng);
      const double part_r = y_iy - cell_frac_y * dy;
      const double part_theta = 2.0 * PI * kiss_uniform(rng);
      double p[7];
      p[0] = x0 + kiss_uniform(rng) * dx;
      p[1] = part_r * std::cos(part_theta);
      p[2] = part_r * std::sin(part_theta);
      const double wdata = (2.0 * PI * dx * dy * part_r) / (double)ncell;
      const double cy2 = cell_frac_y * cell_frac_y;
      const double gy[3] = {0.5 * (0.25 + cy2 + cell_frac_y), 0.75 - cy2, 0.5 * (0.25 + cy2 - cell_frac_y)};
      for (int i = 0; i < 3; ++i) {
        double temp_local = 0.0, drift_local = 0.0;
        for (int k = -1; k <= 1; ++k) {
          temp_local = temp_local + gy[k + 1] * temperature[i * nrow + iy + k];
          drift_local = drift_local + gy[k + 1] * drift[i * nrow + iy + k];
        }
        // momentum_from_temperature, particle_temperature.F90:388-398
        p[3 + i] = kiss_box_muller(rng, std::sqrt(temp_local * KB * S.sp.mass), drift_local);
      }
      double weight_local = 0.0;
      for (int k = -1; k <= 1; ++k) weight_local = weight_local + gy[k + 1] * density[iy + k];
      p[6] = weight_local * wdata;
      aos.insert(aos.end(), p, p + 7);
    }
  }
  return 0;
}


// ------------------------------------------------------------------------------------------
// Device-side column: the same per-particle arithmetic as above
// (window.F90:226-296), but every particle draws from its own counter of a Philox4x32-10 stream
// instead of the rank's sequential KISS stream, so the column is generated by one kernel straight
// into the SoA list -- no host loop over particles, no upload, and the plasma no longer depends
// on how many ranks share the grid.  Not bit-identical to the reference's column by construction
// (a different generator; Gaussians by the trigonometric Box-Muller transform instead of the
// polar rejection loop); the oracle restates this very stream for the parity tests.
//   block 0: (r offset, theta)   block 1: (x offset, ua1)   block 2: (ub1, ua2)   block 3: (ub2, -)
//   (px, py) = sqrt(-2 ln(1 - ua1)) (cos, sin)(2 pi ub1),  pz = sqrt(-2 ln(1 - ua2)) cos(2 pi ub2)
// ------------------------------------------------------------------------------------------
#include "insert_kernel.cuh"

__global__ void k_add_count_ins(int64_t* n_dev, long long add) { *n_dev += add; }

int do_insert_particles_device(cylgpu_ctx* c, int isp, double x_grid_max, double npart_per_cell_real,
                               const double* density_in, const double* temperature, const double* drift, double dmin,
                               double dmax, uint64_t seed, uint64_t column, int64_t* n_inserted) {
  if (n_inserted) *n_inserted = 0;
  if (!c->cfg.x_max_boundary) return 0;   // only the rightmost rank injects (window.F90:176)
  SpeciesState& S = c->species[isp];
  const int ny = c->g.ny, nrow = ny + 2;
  const int64_t npart_per_cell = (int64_t)std::floor(npart_per_cell_real);
  const double npart_frac = npart_per_cell_real - (double)npart_per_cell;
  if (npart_per_cell + 1 >= (int64_t)1 << 30) { set_error("insert_particles_device: npart_per_cell too large"); return 2; }
  const ColumnStream rs = column_stream(seed, isp, column);
  // staging: 7 profile rows then the row offsets, pinned so that the upload never blocks
  const size_t words = (size_t)7 * nrow + (size_t)(ny + 1);
  if (c->ins_cap < words) {
    if (c->ins_ev) CUDA_TRY(cudaEventSynchronize(c->ins_ev));
    if (c->ins_pin) cudaFreeHost(c->ins_pin);
    if (c->ins_dev) cudaFree(c->ins_dev);
    c->ins_pin = nullptr; c->ins_dev = nullptr; c->ins_cap = 0;
    CUDA_TRY(cudaMallocHost(&c->ins_pin, words * sizeof(double)));
    CUDA_TRY(cudaMalloc(&c->ins_dev, words * sizeof(double)));
    if (!c->ins_ev) CUDA_TRY(cudaEventCreateWithFlags(&c->ins_ev, cudaEventDisableTiming));
    c->ins_cap = words;
  } else {
    CUDA_TRY(cudaEventSynchronize(c->ins_ev));   // the previous column's upload has left the pinned buffer
  }
  double* prof = c->ins_pin;
  int64_t* row_start = reinterpret_cast<int64_t*>(c->ins_pin + (size_t)7 * nrow);
  for (int iy = 0; iy < nrow; ++iy) {   // window.F90:203-209
    double d = density_in[iy];
    if (d > dmax) d = dmax;
    if (d < dmin) d = 0.0;
    prof[iy] = d;
  }
  for (int i = 0; i < 3 * nrow; ++i) {
    prof[nrow + i] = temperature[i];
    prof[4 * nrow + i] = drift[i];
  }
  // every rank owns all radial cells (nprocy = 1): the global radial index is the local one
  const int iy_global_offset = 0;
  int64_t total = 0;
  row_start[0] = 0;
  for (int iy = 1; iy <= ny; ++iy) {
    int64_t ncell = 0;
    if (!(prof[iy] < dmin)) {
      int64_t n_frac = 0;
      if (npart_frac > 0.0 && rs.cell_uniform((uint32_t)(iy + iy_global_offset)) < npart_frac) n_frac = 1;
      ncell = npart_per_cell + n_frac;
    }
    total += ncell;
    row_start[iy] = total;
  }
  if (n_inserted) *n_inserted = total;
  if (total == 0) return 0;
  TRY(presort_join(c, false));
  TRY(reserve_particles(c, isp, S.n + total));
  CUDA_TRY(cudaMemcpyAsync(c->ins_dev, c->ins_pin, words * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaEventRecord(c->ins_ev, c->stream));
  ColumnArgs a;
  a.rs = rs;
  a.prof = c->ins_dev;
  a.row_start = reinterpret_cast<const int64_t*>(c->ins_dev + (size_t)7 * nrow);
  a.x = S.d[0]; a.y = S.d[1]; a.z = S.d[2]; a.px = S.d[3]; a.py = S.d[4]; a.pz = S.d[5]; a.w = S.d[6];
  a.base = S.n;
  a.base_dev = S.lazy ? c->n_dev + isp : nullptr;
  a.ny = ny;
  a.iy_global_offset = iy_global_offset;
  a.dx = c->cfg.dx; a.dy = c->cfg.dy;
  a.x0 = x_grid_max + 0.5 * c->cfg.dx;
  a.y_grid_min_local = c->cfg.y_grid_min_local;
  a.mass = S.sp.mass;
  k_insert_column<<<ny, 128, 0, c->stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  c->stats.kernel_launches += 1;
  S.n += total;   // exact, or the upper bound moving with the device-side count
  if (S.lazy) {
    k_add_count_ins<<<1, 1, 0, c->stream>>>(c->n_dev + isp, (long long)total);
    c->stats.kernel_launches += 1;
  } else {
    TRY(set_count_exact(c, isp));
  }
  c->stats.n_particles[isp] = S.n;
  c->sorted_valid = false;
  return 0;
}

}  // namespace cylgpu
