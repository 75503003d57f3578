// insert_kernel.cuh -- the device-side plasma column of the moving window (host side:
// window_insert.cu; stream layout: philox.cuh / include/cylgpu.h).  Kernel only, included inside
// namespace cylgpu after philox.cuh; also compiled for the CPU by the kernel-emulation tests
// (tests/emul/).  Product code: no oracle here.
#pragma once

struct ColumnArgs {
  ColumnStream rs;
  const double* prof;        // [density | temperature(3) | drift(3)] x (ny + 2), density already clamped
  const int64_t* row_start;  // ny + 1 entries, row iy -> [row_start[iy-1], row_start[iy])
  double *x, *y, *z, *px, *py, *pz, *w;
  int64_t base;
  const int64_t* base_dev;   // not null: the list length lives on the device (device-resident counts)
  int ny, iy_global_offset;
  double dx, dy, x0, y_grid_min_local, mass;
};

__global__ void __launch_bounds__(128) k_insert_column(ColumnArgs a) {
  const int iy = blockIdx.x + 1;
  const int64_t r0 = a.row_start[iy - 1];
  const int64_t ncell = a.row_start[iy] - r0;
  const int nrow = a.ny + 2;
  const double y_iy = a.y_grid_min_local + (double)(iy - 1) * a.dy;
  for (int64_t ip = threadIdx.x; ip < ncell; ip += blockDim.x) {
    const uint32_t iyg = (uint32_t)(iy + a.iy_global_offset);
    const Philox4 b0 = a.rs.block(iyg, (uint32_t)ip, 0), b1 = a.rs.block(iyg, (uint32_t)ip, 1);
    const Philox4 b2 = a.rs.block(iyg, (uint32_t)ip, 2), b3 = a.rs.block(iyg, (uint32_t)ip, 3);
    const double cell_frac_y = 0.5 - philox_u53(b0.v[0], b0.v[1]);
    const double part_r = y_iy - cell_frac_y * a.dy;
    const double part_theta = 2.0 * PI * philox_u53(b0.v[2], b0.v[3]);
    const double X = a.x0 + philox_u53(b1.v[0], b1.v[1]) * a.dx;
    const double wdata = (2.0 * PI * a.dx * a.dy * part_r) / (double)ncell;
    const double cy2 = cell_frac_y * cell_frac_y;
    const double gy[3] = {0.5 * (0.25 + cy2 + cell_frac_y), 0.75 - cy2, 0.5 * (0.25 + cy2 - cell_frac_y)};
    const double rad1 = sqrt(-2.0 * log(1.0 - philox_u53(b1.v[2], b1.v[3])));
    const double ang1 = 2.0 * PI * philox_u53(b2.v[0], b2.v[1]);
    const double rad2 = sqrt(-2.0 * log(1.0 - philox_u53(b2.v[2], b2.v[3])));
    const double ang2 = 2.0 * PI * philox_u53(b3.v[0], b3.v[1]);
    const double gauss[3] = {rad1 * cos(ang1), rad1 * sin(ang1), rad2 * cos(ang2)};
    double p[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double temp_local = 0.0, drift_local = 0.0;
#pragma unroll
      for (int k = -1; k <= 1; ++k) {
        temp_local = temp_local + gy[k + 1] * a.prof[(1 + i) * nrow + iy + k];
        drift_local = drift_local + gy[k + 1] * a.prof[(4 + i) * nrow + iy + k];
      }
      p[i] = gauss[i] * sqrt(temp_local * KB * a.mass) + drift_local;   // particle_temperature.F90:388-398
    }
    double weight_local = 0.0;
#pragma unroll
    for (int k = -1; k <= 1; ++k) weight_local = weight_local + gy[k + 1] * a.prof[iy + k];
    const int64_t o = (a.base_dev ? *a.base_dev : a.base) + r0 + ip;
    a.x[o] = X;
    a.y[o] = part_r * cos(part_theta);
    a.z[o] = part_r * sin(part_theta);
    a.px[o] = p[0];
    a.py[o] = p[1];
    a.pz[o] = p[2];
    a.w[o] = weight_local * wdata;
  }
}
