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

}  // namespace cylgpu
