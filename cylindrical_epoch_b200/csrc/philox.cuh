// philox.cuh -- Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as
// 1, 2, 3", SC'11), the counter-based generator behind the device-side plasma column of the
// moving window (window_insert.cu).  Host + device, pinned by the Random123 known-answer vectors
// in tests/test_host_logic.py through cylgpu_philox4x32.
#pragma once
#include <stdint.h>

namespace cylgpu {

struct Philox4 { uint32_t v[4]; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    if (r > 0) { k0 += W0; k1 += W1; }
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
  }
  Philox4 o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

// 53-bit uniform in [0, 1) from two words (exact in double)
__host__ __device__ __forceinline__ double philox_u53(uint32_t hi, uint32_t lo) {
  return (double)(((uint64_t)hi << 21) | (uint64_t)(lo >> 11)) * (1.0 / 9007199254740992.0);
}

// The stream of the device-side column (documented in include/cylgpu.h):
//   key     = (seed low word, seed high word + species index)
//   counter = (column low word, column high word, global radial cell iy, 4 * ip + block), ip = particle in cell,
//             and (.., .., iy, 0xFFFFFFFF) for the cell's fractional-particle decision
struct ColumnStream {
  uint32_t k0, k1, col_lo, col_hi;
  __host__ __device__ __forceinline__ Philox4 block(uint32_t iy, uint32_t ip, uint32_t b) const {
    return philox4x32_10(col_lo, col_hi, iy, 4u * ip + b, k0, k1);
  }
  __host__ __device__ __forceinline__ double cell_uniform(uint32_t iy) const {
    const Philox4 p = philox4x32_10(col_lo, col_hi, iy, 0xFFFFFFFFu, k0, k1);
    return philox_u53(p.v[0], p.v[1]);
  }
};

__host__ __device__ __forceinline__ ColumnStream column_stream(uint64_t seed, int species, uint64_t column) {
  ColumnStream s;
  s.k0 = (uint32_t)seed;
  s.k1 = (uint32_t)(seed >> 32) + (uint32_t)species;
  s.col_lo = (uint32_t)column;
  s.col_hi = (uint32_t)(column >> 32);
  return s;
}

}  // namespace cylgpu
