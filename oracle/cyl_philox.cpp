// cyl_philox.cpp -- CPU oracle, part 3: the counter-based plasma column of the moving window.
//
// TEST INFRASTRUCTURE ONLY (see cyl_oracle.hpp).  The reference generates the new column of
// insert_particles (window.F90:157-300) from the rank's sequential KISS stream; the product's
// device-side variant (cylgpu_insert_particles_device) keeps the reference's
// per-particle arithmetic (window.F90:226-296) but draws from Philox4x32-10 counters.  This file
// restates (i) Philox4x32-10 from its publication (Salmon et al., SC'11; Random123 1.09
// philox.h: multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key increments 0x9E3779B9 / 0xBB67AE85,
// ten rounds), pinned by the Random123 known-answer vectors in tests/test_oracle_moments.py,
// and (ii) the stream layout documented in include/cylgpu.h, written independently of the
// product source.
#include <cmath>
#include <cstdint>

#include "cyl_oracle.hpp"

namespace cylo {

void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  uint32_t k[2] = {key[0], key[1]};
  for (int round = 0; round < 10; ++round) {
    if (round != 0) {
      k[0] += 0x9E3779B9u;
      k[1] += 0xBB67AE85u;
    }
    const uint64_t prod_a = (uint64_t)0xD2511F53u * (uint64_t)c[0];
    const uint64_t prod_b = (uint64_t)0xCD9E8D57u * (uint64_t)c[2];
    const uint32_t hi_a = (uint32_t)(prod_a >> 32), lo_a = (uint32_t)(prod_a & 0xFFFFFFFFu);
    const uint32_t hi_b = (uint32_t)(prod_b >> 32), lo_b = (uint32_t)(prod_b & 0xFFFFFFFFu);
    const uint32_t n[4] = {hi_b ^ c[1] ^ k[0], lo_b, hi_a ^ c[3] ^ k[1], lo_a};
    for (int i = 0; i < 4; ++i) c[i] = n[i];
  }
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}

namespace {

// 53 random bits -> [0, 1): the high word supplies bits 52..21, the low word's top 21 bits the rest
double uniform53(uint32_t high, uint32_t low) {
  const uint64_t bits = ((uint64_t)high << 21) | ((uint64_t)low >> 11);
  return std::ldexp((double)bits, -53);
}

struct Draws {   // the seven uniforms of one particle
  double r_offset, theta, x_offset, ua1, ub1, ua2, ub2;
};

Draws particle_draws(uint64_t seed, int isp, uint64_t column, uint32_t iy, uint32_t ip) {
  const uint32_t key[2] = {(uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32) + (uint32_t)isp};
  uint32_t w[4][4];
  for (uint32_t b = 0; b < 4; ++b) {
    const uint32_t ctr[4] = {(uint32_t)(column & 0xFFFFFFFFu), (uint32_t)(column >> 32), iy, 4u * ip + b};
    philox4x32_10(ctr, key, w[b]);
  }
  Draws d;
  d.r_offset = uniform53(w[0][0], w[0][1]);
  d.theta = uniform53(w[0][2], w[0][3]);
  d.x_offset = uniform53(w[1][0], w[1][1]);
  d.ua1 = uniform53(w[1][2], w[1][3]);
  d.ub1 = uniform53(w[2][0], w[2][1]);
  d.ua2 = uniform53(w[2][2], w[2][3]);
  d.ub2 = uniform53(w[3][0], w[3][1]);
  return d;
}

double cell_draw(uint64_t seed, int isp, uint64_t column, uint32_t iy) {
  const uint32_t key[2] = {(uint32_t)(seed & 0xFFFFFFFFu), (uint32_t)(seed >> 32) + (uint32_t)isp};
  const uint32_t ctr[4] = {(uint32_t)(column & 0xFFFFFFFFu), (uint32_t)(column >> 32), iy, 0xFFFFFFFFu};
  uint32_t w[4];
  philox4x32_10(ctr, key, w);
  return uniform53(w[0], w[1]);
}

}  // namespace

// window.F90:157-300 with counter-based draws; column = number of shifts made so far
void World::insert_particles_counter(Rank& r) {
  if (!r.x_max_boundary) return;
  const uint64_t column = (uint64_t)window_shifts_total;
  const double x_grid_max = x_grid_min + (double)(cfg.nx_global - 1) * dx;
  for (size_t isp = 0; isp < species.size(); ++isp) {
    const Species& s = species[isp];
    if (!(s.npart_per_cell > 0.0) || !(s.density > 0.0)) continue;
    const int64_t npart_per_cell = (int64_t)std::floor(s.npart_per_cell);
    const double npart_frac = s.npart_per_cell - (double)npart_per_cell;
    const double x0 = x_grid_max + 0.5 * dx;
    for (int iy = 1; iy <= r.ny; ++iy) {
      int64_t n_frac = 0;
      if (npart_frac > 0.0 && cell_draw(counter_seed, (int)isp, column, (uint32_t)iy) < npart_frac) n_frac = 1;
      const int64_t ncell = npart_per_cell + n_frac;
      for (int64_t ip = 0; ip < ncell; ++ip) {
        const Draws d = particle_draws(counter_seed, (int)isp, column, (uint32_t)iy, (uint32_t)ip);
        Particle p;
        const double cell_frac_y = 0.5 - d.r_offset;
        const double yc = y_grid_min_local + (double)(iy - 1) * dy;
        const double part_r = yc - cell_frac_y * dy;
        const double part_theta = 2.0 * PI * d.theta;
        p.pos[0] = x0 + d.x_offset * dx;
        p.pos[1] = part_r * std::cos(part_theta);
        p.pos[2] = part_r * std::sin(part_theta);
        const double wdata = (2.0 * PI * dx * dy * part_r) / (double)ncell;
        const double cy2 = cell_frac_y * cell_frac_y;
        const double gy[3] = {0.5 * (0.25 + cy2 + cell_frac_y), 0.75 - cy2, 0.5 * (0.25 + cy2 - cell_frac_y)};
        // trigonometric Box-Muller: two Gaussians from the first pair, one from the second
        const double rad1 = std::sqrt(-2.0 * std::log(1.0 - d.ua1));
        const double rad2 = std::sqrt(-2.0 * std::log(1.0 - d.ua2));
        const double gauss[3] = {rad1 * std::cos(2.0 * PI * d.ub1), rad1 * std::sin(2.0 * PI * d.ub1),
                                 rad2 * std::cos(2.0 * PI * d.ub2)};
        for (int i = 0; i < 3; ++i) {
          double temp_local = 0.0, drift_local = 0.0;
          for (int k = 0; k < 3; ++k) {
            temp_local = temp_local + gy[k] * s.temp[i];
            drift_local = drift_local + gy[k] * s.drift[i];
          }
          p.p[i] = gauss[i] * std::sqrt(temp_local * KB * s.mass) + drift_local;
        }
        double weight_local = 0.0;
        for (int k = 0; k < 3; ++k) weight_local = weight_local + gy[k] * s.density;
        p.w = weight_local * wdata;
        r.parts[isp].push_back(p);
      }
    }
  }
}

}  // namespace cylo
