// emul_kernels.cpp -- TEST INFRASTRUCTURE: runs the product's kernel SOURCES (the very headers nvcc
// compiles into libcylgpu.so) on the CPU through cuda_emul.hpp, with the launch sequences of
// csrc/moments.cuh::do_particle_moment and csrc/window_insert.cu::do_insert_particles_device
// restated for one slab.  tests/test_kernel_emulation.py compares the results with the oracle.
#include <cstring>
#include <vector>

#include "cuda_emul.hpp"

#include "../../include/cylgpu.h"
#include "../../cylindrical_epoch_b200/csrc/geom.cuh"
#include "../../cylindrical_epoch_b200/csrc/field_ranges.cuh"
#include "../../cylindrical_epoch_b200/csrc/philox.cuh"
// the push arithmetic: IEEE division / square root instead of the PTX seeds (inline asm cannot run here)
#define CYL_EMUL
#define CYL_REFERENCE_MATH
#include "../../cylindrical_epoch_b200/csrc/push.cuh"

namespace cylgpu {
#include "../../cylindrical_epoch_b200/csrc/bc_kernels.cuh"
#include "../../cylindrical_epoch_b200/csrc/moments_kernels.cuh"
#include "../../cylindrical_epoch_b200/csrc/insert_kernel.cuh"
#include "../../cylindrical_epoch_b200/csrc/push_v0.cuh"
#include "../../cylindrical_epoch_b200/csrc/push_shapes.cuh"
#include "../../cylindrical_epoch_b200/csrc/pbcs_kernels.cuh"
#include "../../cylindrical_epoch_b200/csrc/compact_kernels.cuh"
#include "../../cylindrical_epoch_b200/csrc/field_kernels.cuh"
}  // namespace cylgpu

using namespace cylgpu;

namespace {

struct SlabCfg {
  Geom g1;
  int bca[4], bc_field[4];
  bool periodic_self;   // one slab whose x neighbours are itself (left = right = rank 0)
};

// exchange3 of bcs.cu for one slab: pack, what goes left arrives from the right and vice versa, unpack
void exchange_self(const SlabCfg& S, cplx* a0, cplx* a1, int mode) {
  const Geom& g = S.g1;
  Halo3 h;
  h.f[0] = a0; h.f[1] = a1; h.f[2] = nullptr;
  h.skip[0] = h.skip[1] = h.skip[2] = 0;
  const size_t elems = (size_t)3 * g.M * g.SY * NG;
  std::vector<cplx> sl(2 * elems), sr(2 * elems);
  const dim3 grd((g.SY * NG + 127) / 128, g.M, 3);
  emul_launch(k_halo_pack, grd, dim3(128), g, h, sl.data(), sr.data(), mode, elems);
  emul_launch(k_halo_unpack, grd, dim3(128), g, h, (const cplx*)sr.data(), (const cplx*)sl.data(), mode, elems);
}

void summation_bcs(const SlabCfg& S, cplx* a0, cplx* a1) {   // moments.cuh::moment_summation_bcs
  const Geom& g1 = S.g1;
  const dim3 gx_((g1.SY + 127) / 128, 1), gy_((g1.SX + 127) / 128, 1);
  cplx* arr[2] = {a0, a1};
  for (int k = 0; k < 2; ++k) {
    cplx* a = arr[k];
    if (!a) continue;
    if (S.bca[CYLGPU_BD_X_MIN] == CYLGPU_BC_REFLECT) emul_launch(k_density_reflect, gx_, dim3(128), g1, a, (int)CYLGPU_BD_X_MIN);
    if (S.bca[CYLGPU_BD_X_MAX] == CYLGPU_BC_REFLECT) emul_launch(k_density_reflect, gx_, dim3(128), g1, a, (int)CYLGPU_BD_X_MAX);
    if (S.bca[CYLGPU_BD_Y_MAX] == CYLGPU_BC_REFLECT) emul_launch(k_density_reflect, gy_, dim3(128), g1, a, (int)CYLGPU_BD_Y_MAX);
  }
  const bool to_l = S.periodic_self && S.bca[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC;
  const bool to_r = S.periodic_self && S.bca[CYLGPU_BD_X_MAX] == CYLGPU_BC_PERIODIC;
  if (to_l && to_r) exchange_self(S, a0, a1, 1);
}

void zero_gradient(const SlabCfg& S, cplx* a) {   // moments.cuh::moment_zero_gradient
  const Geom& g1 = S.g1;
  const dim3 gx_((g1.SY + 127) / 128, 1), gy_((g1.SX + 127) / 128, 1);
  if (S.bc_field[CYLGPU_BD_X_MIN] != CYLGPU_BC_PERIODIC) emul_launch(k_density_zero_gradient, gx_, dim3(128), g1, a, (int)CYLGPU_BD_X_MIN);
  if (S.bc_field[CYLGPU_BD_X_MAX] != CYLGPU_BC_PERIODIC) emul_launch(k_density_zero_gradient, gx_, dim3(128), g1, a, (int)CYLGPU_BD_X_MAX);
  if (S.bc_field[CYLGPU_BD_Y_MIN] != CYLGPU_BC_PERIODIC) emul_launch(k_density_zero_gradient, gy_, dim3(128), g1, a, (int)CYLGPU_BD_Y_MIN);
  if (S.bc_field[CYLGPU_BD_Y_MAX] != CYLGPU_BC_PERIODIC) emul_launch(k_density_zero_gradient, gy_, dim3(128), g1, a, (int)CYLGPU_BD_Y_MAX);
}

}  // namespace

#define EMUL_API __attribute__((visibility("default")))

extern "C" {

// One slab (x_min_boundary = x_max_boundary = true).  soa[s]: 7 arrays of n[s] doubles (x y z px py pz w);
// bca: common particle bc per boundary; out: real (nx+2ng) x (ny+2ng) plane.  Returns 0, or 2 for an
// unknown moment (the product's error path).
EMUL_API int emul_particle_moment(int nx, int ny, int kind, int direction, int nsel, const double* const* soa, const int64_t* n,
                         const double* mass, const double* charge, double x_grid_min_local, double y_grid_min_local,
                         double dx, double dy, const int32_t* bca, const int32_t* bc_field, double* out) {
  SlabCfg S;
  S.g1.nx = nx; S.g1.ny = ny; S.g1.M = 1;
  S.g1.SX = nx + 2 * NG; S.g1.SY = ny + 2 * NG;
  S.g1.plane = (size_t)S.g1.SX * S.g1.SY;
  for (int i = 0; i < 4; ++i) { S.bca[i] = bca[i]; S.bc_field[i] = bc_field[i]; }
  S.periodic_self = bc_field[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC;
  const Geom g1 = S.g1;
  const size_t np = g1.plane;
  std::vector<cplx> A(np, C(0.0, 0.0)), B(np, C(0.0, 0.0)), D(np, C(0.0, 0.0));
  auto for_each_species = [&](auto&& launch) {
    for (int s = 0; s < nsel; ++s) {
      if (n[s] == 0) continue;
      MomentArgs a;
      a.x = soa[7 * s + 0]; a.y = soa[7 * s + 1]; a.z = soa[7 * s + 2];
      a.px = soa[7 * s + 3]; a.py = soa[7 * s + 4]; a.pz = soa[7 * s + 5]; a.w = soa[7 * s + 6];
      a.n = n[s];
      a.x_grid_min_local = x_grid_min_local; a.y_grid_min_local = y_grid_min_local;
      a.dx = dx; a.dy = dy;
      a.mass = mass[s]; a.charge = charge[s];
      a.kind = kind; a.direction = direction;
      launch(a, dim3((unsigned)((n[s] + 255) / 256)));
    }
  };
  const dim3 pb((unsigned)((np + 255) / 256));
  int finish_mode = 0;
  double dof = 1.0;
  cplx* result = A.data();
  if (kind == CYLGPU_MOM_PPC || kind == CYLGPU_MOM_AVERAGE_WEIGHT) {
    for_each_species([&](const MomentArgs& a, dim3 blocks) { emul_launch(k_moment_count, blocks, dim3(256), g1, a, (double*)A.data()); });
    finish_mode = kind == CYLGPU_MOM_AVERAGE_WEIGHT ? 1 : 0;
  } else if (kind == CYLGPU_MOM_TEMPERATURE) {
    for_each_species([&](const MomentArgs& a, dim3 blocks) {
      emul_launch(k_temperature_means, blocks, dim3(256), g1, a, (double*)A.data(), (double*)B.data());
    });
    summation_bcs(S, A.data(), B.data());
    emul_launch(k_temperature_normalise, pb, dim3(256), A.data(), B.data(), np);
    if (S.periodic_self) exchange_self(S, A.data(), B.data(), 0);
    for_each_species([&](const MomentArgs& a, dim3 blocks) {
      emul_launch(k_temperature_sigma, blocks, dim3(256), g1, a, (const cplx*)A.data(), (const cplx*)B.data(), (double*)D.data());
    });
    summation_bcs(S, D.data(), nullptr);
    finish_mode = 2;
    dof = direction > 0 ? 1.0 : 3.0;
    result = D.data();
  } else if (kind == CYLGPU_MOM_MASS_DENSITY || kind == CYLGPU_MOM_NUMBER_DENSITY || kind == CYLGPU_MOM_SPECIES_CURRENT ||
             kind == CYLGPU_MOM_EKBAR || kind == CYLGPU_MOM_EKFLUX || kind == CYLGPU_MOM_AVERAGE_MOMENTUM) {
    for_each_species([&](const MomentArgs& a, dim3 blocks) { emul_launch(k_moment_deposit, blocks, dim3(256), g1, a, (double*)A.data()); });
    summation_bcs(S, A.data(), nullptr);
    zero_gradient(S, A.data());
    finish_mode = (kind == CYLGPU_MOM_EKBAR || kind == CYLGPU_MOM_EKFLUX || kind == CYLGPU_MOM_AVERAGE_MOMENTUM) ? 1 : 0;
  } else {
    return 2;
  }
  emul_launch(k_moment_finish, pb, dim3(256), (const cplx*)result, out, np, finish_mode, dof);
  return 0;
}

// calc_number_density_modes (charge = 0) / calc_charge_density (charge = 1) on ONE slab: bcs.cu::
// do_number_density_modes -- deposit with the mode factors, reflection, additive ghost exchange, zero gradient.
// out: complex (nx+2ng, ny+2ng, M) array, zeroed by the caller.
EMUL_API void emul_number_density_modes(int nx, int ny, int M, int nsel, const double* const* soa, const int64_t* n,
                                        const double* charge_of, int charge, double x_grid_min_local,
                                        double y_grid_min_local, double dx, double dy, const int32_t* bca,
                                        const int32_t* bc_field, void* out) {
  Geom g;
  g.nx = nx; g.ny = ny; g.M = M;
  g.SX = nx + 2 * NG; g.SY = ny + 2 * NG;
  g.plane = (size_t)g.SX * g.SY;
  cplx* a = (cplx*)out;
  for (int s = 0; s < nsel; ++s)
    if (n[s] > 0)
      emul_launch(k_number_density, dim3((unsigned)((n[s] + 255) / 256)), dim3(256), g, soa[7 * s + 0], soa[7 * s + 1],
                  soa[7 * s + 2], soa[7 * s + 6], n[s], (double*)a, x_grid_min_local, y_grid_min_local, dx, dy,
                  charge ? charge_of[s] : 1.0, charge ? 1 : M);
  const dim3 gx_((g.SY + 127) / 128, g.M), gy_((g.SX + 127) / 128, g.M);
  if (bca[CYLGPU_BD_X_MIN] == CYLGPU_BC_REFLECT) emul_launch(k_density_reflect, gx_, dim3(128), g, a, (int)CYLGPU_BD_X_MIN);
  if (bca[CYLGPU_BD_X_MAX] == CYLGPU_BC_REFLECT) emul_launch(k_density_reflect, gx_, dim3(128), g, a, (int)CYLGPU_BD_X_MAX);
  if (bca[CYLGPU_BD_Y_MAX] == CYLGPU_BC_REFLECT) emul_launch(k_density_reflect, gy_, dim3(128), g, a, (int)CYLGPU_BD_Y_MAX);
  if (bc_field[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC && bca[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC) {
    Halo3 h;
    h.f[0] = a; h.f[1] = nullptr; h.f[2] = nullptr;
    h.skip[0] = h.skip[1] = h.skip[2] = 0;
    const size_t elems = (size_t)3 * g.M * g.SY * NG;
    std::vector<cplx> sl(elems), sr(elems);
    const dim3 grd((g.SY * NG + 127) / 128, g.M, 3);
    emul_launch(k_halo_pack, grd, dim3(128), g, h, sl.data(), sr.data(), 1, elems);
    emul_launch(k_halo_unpack, grd, dim3(128), g, h, (const cplx*)sr.data(), (const cplx*)sl.data(), 1, elems);
  }
  if (bc_field[CYLGPU_BD_X_MIN] != CYLGPU_BC_PERIODIC) emul_launch(k_density_zero_gradient, gx_, dim3(128), g, a, (int)CYLGPU_BD_X_MIN);
  if (bc_field[CYLGPU_BD_X_MAX] != CYLGPU_BC_PERIODIC) emul_launch(k_density_zero_gradient, gx_, dim3(128), g, a, (int)CYLGPU_BD_X_MAX);
  if (bc_field[CYLGPU_BD_Y_MIN] != CYLGPU_BC_PERIODIC) emul_launch(k_density_zero_gradient, gy_, dim3(128), g, a, (int)CYLGPU_BD_Y_MIN);
  if (bc_field[CYLGPU_BD_Y_MAX] != CYLGPU_BC_PERIODIC) emul_launch(k_density_zero_gradient, gy_, dim3(128), g, a, (int)CYLGPU_BD_Y_MAX);
}

// The density / averaged moments on TWO slabs that are x neighbours (slab 0 owns x_min, slab 1 owns x_max;
// periodic: they are also each other's outer neighbours, as in a 2-rank ring): deposit, reflection, the
// additive ghost exchange of moments.cuh::moment_summation_bcs with the send / receive flags of the product,
// zero-gradient fill, finish.  kind: one of the k_moment_deposit moments.  soa[k]: 7 arrays of slab k.
EMUL_API int emul_moment_two_slabs(const int* nx2, int ny, int kind, int direction, const double* const* soa,
                                   const int64_t* n2, double mass, double charge, const double* x_grid_min_local2,
                                   double y_grid_min_local, double dx, double dy, const int32_t* bca,
                                   const int32_t* bc_field, double* const* out2) {
  const bool averaged = (kind == CYLGPU_MOM_EKBAR || kind == CYLGPU_MOM_EKFLUX || kind == CYLGPU_MOM_AVERAGE_MOMENTUM);
  if (!(averaged || kind == CYLGPU_MOM_MASS_DENSITY || kind == CYLGPU_MOM_NUMBER_DENSITY ||
        kind == CYLGPU_MOM_SPECIES_CURRENT)) return 2;
  Geom g[2];
  std::vector<cplx> A[2];
  std::vector<cplx> sl[2], sr[2];
  const bool periodic = bca[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC;
  for (int k = 0; k < 2; ++k) {
    g[k].nx = nx2[k]; g[k].ny = ny; g[k].M = 1;
    g[k].SX = nx2[k] + 2 * NG; g[k].SY = ny + 2 * NG;
    g[k].plane = (size_t)g[k].SX * g[k].SY;
    A[k].assign(g[k].plane, C(0.0, 0.0));
    MomentArgs a;
    a.x = soa[7 * k + 0]; a.y = soa[7 * k + 1]; a.z = soa[7 * k + 2];
    a.px = soa[7 * k + 3]; a.py = soa[7 * k + 4]; a.pz = soa[7 * k + 5]; a.w = soa[7 * k + 6];
    a.n = n2[k];
    a.x_grid_min_local = x_grid_min_local2[k]; a.y_grid_min_local = y_grid_min_local;
    a.dx = dx; a.dy = dy; a.mass = mass; a.charge = charge; a.kind = kind; a.direction = direction;
    if (a.n > 0) emul_launch(k_moment_deposit, dim3((unsigned)((a.n + 255) / 256)), dim3(256), g[k], a, (double*)A[k].data());
    const bool xminb = (k == 0), xmaxb = (k == 1);
    const dim3 gx_((g[k].SY + 127) / 128, 1), gy_((g[k].SX + 127) / 128, 1);
    if (xminb && bca[CYLGPU_BD_X_MIN] == CYLGPU_BC_REFLECT) emul_launch(k_density_reflect, gx_, dim3(128), g[k], A[k].data(), (int)CYLGPU_BD_X_MIN);
    if (xmaxb && bca[CYLGPU_BD_X_MAX] == CYLGPU_BC_REFLECT) emul_launch(k_density_reflect, gx_, dim3(128), g[k], A[k].data(), (int)CYLGPU_BD_X_MAX);
    if (bca[CYLGPU_BD_Y_MAX] == CYLGPU_BC_REFLECT) emul_launch(k_density_reflect, gy_, dim3(128), g[k], A[k].data(), (int)CYLGPU_BD_Y_MAX);
  }
  // pack on both slabs, then each unpacks what its neighbours sent (moments.cuh: to_l / to_r)
  bool to_l[2], to_r[2];
  const size_t elems = (size_t)3 * g[0].SY * NG;
  for (int k = 0; k < 2; ++k) {
    const bool xminb = (k == 0), xmaxb = (k == 1);
    const bool has_l = !xminb || periodic, has_r = !xmaxb || periodic;   // decomp: the ring closes only when periodic
    to_l[k] = has_l && !(xminb && bca[CYLGPU_BD_X_MIN] != CYLGPU_BC_PERIODIC);
    to_r[k] = has_r && !(xmaxb && bca[CYLGPU_BD_X_MAX] != CYLGPU_BC_PERIODIC);
    sl[k].assign(elems, C(0.0, 0.0)); sr[k].assign(elems, C(0.0, 0.0));
    Halo3 h;
    h.f[0] = A[k].data(); h.f[1] = nullptr; h.f[2] = nullptr;
    h.skip[0] = h.skip[1] = h.skip[2] = 0;
    const dim3 grd((g[k].SY * NG + 127) / 128, 1, 3);
    if (to_l[k] || to_r[k])
      emul_launch(k_halo_pack, grd, dim3(128), g[k], h, to_l[k] ? sl[k].data() : (cplx*)nullptr,
                  to_r[k] ? sr[k].data() : (cplx*)nullptr, 1, elems);
  }
  for (int k = 0; k < 2; ++k) {
    const int other = 1 - k;   // left and right neighbour of a 2-ring are the same slab
    Halo3 h;
    h.f[0] = A[k].data(); h.f[1] = nullptr; h.f[2] = nullptr;
    h.skip[0] = h.skip[1] = h.skip[2] = 0;
    const dim3 grd((g[k].SY * NG + 127) / 128, 1, 3);
    // what arrives from my left is my left neighbour's right-going message, and vice versa
    const cplx* recv_l = to_l[k] ? sr[other].data() : nullptr;
    const cplx* recv_r = to_r[k] ? sl[other].data() : nullptr;
    if (recv_l || recv_r) emul_launch(k_halo_unpack, grd, dim3(128), g[k], h, recv_l, recv_r, 1, elems);
  }
  for (int k = 0; k < 2; ++k) {
    const bool xminb = (k == 0), xmaxb = (k == 1);
    const dim3 gx_((g[k].SY + 127) / 128, 1), gy_((g[k].SX + 127) / 128, 1);
    if (bc_field[CYLGPU_BD_X_MIN] != CYLGPU_BC_PERIODIC && xminb) emul_launch(k_density_zero_gradient, gx_, dim3(128), g[k], A[k].data(), (int)CYLGPU_BD_X_MIN);
    if (bc_field[CYLGPU_BD_X_MAX] != CYLGPU_BC_PERIODIC && xmaxb) emul_launch(k_density_zero_gradient, gx_, dim3(128), g[k], A[k].data(), (int)CYLGPU_BD_X_MAX);
    if (bc_field[CYLGPU_BD_Y_MIN] != CYLGPU_BC_PERIODIC) emul_launch(k_density_zero_gradient, gy_, dim3(128), g[k], A[k].data(), (int)CYLGPU_BD_Y_MIN);
    if (bc_field[CYLGPU_BD_Y_MAX] != CYLGPU_BC_PERIODIC) emul_launch(k_density_zero_gradient, gy_, dim3(128), g[k], A[k].data(), (int)CYLGPU_BD_Y_MAX);
    emul_launch(k_moment_finish, dim3((unsigned)((g[k].plane + 255) / 256)), dim3(256), (const cplx*)A[k].data(), out2[k],
                g[k].plane, averaged ? 1 : 0, 1.0);
  }
  return 0;
}

// push_particles without particle_bcs for one species of one slab (cylgpu_push_no_bcs with push variant 0):
static int g_reference_quirks = 1;
EMUL_API void emul_set_reference_quirks(int on) { g_reference_quirks = on; }
static int g_push_generic = 0;
EMUL_API void emul_set_push_generic(int on) { g_push_generic = on; }
EMUL_API int emul_ghost_cells() { return NG; }

// the radial tables of particles.cu::build_tables, the PushConst of push_species, k_push_v0<M> over the list
// and k_r_min_final.  fields: the six E/B mode arrays, jx/jr/jt: zeroed J arrays to deposit into (complex,
// with ghosts); soa: 7 arrays of n doubles, updated in place.
EMUL_API int emul_push_v0(int nx, int ny, int M, const void* const* fields6, void* const* j3, double* const* soa,
                          int64_t n, double charge, double mass, int zero_current, int hc_push, double dt, double dx,
                          double dy, double x_grid_min_local, double y_grid_min_local, double taylor_switch) {
  Geom g;
  g.nx = nx; g.ny = ny; g.M = M;
  g.SX = nx + 2 * NG; g.SY = ny + 2 * NG;
  g.plane = (size_t)g.SX * g.SY;
  const int ntab = ny + 2 * JNG + 1;
  std::vector<double> t(4 * (size_t)ntab, 0.0);
  {   // particles.F90:190-217, r_low accumulated by repeated addition
    double* rt = t.data();
    double* xt = rt + ntab;
    double* vol = xt + ntab;
    double* ratio = vol + ntab;
    double r_low = y_grid_min_local - (double)JNG * dy;
    xt[0] = 1.0 / (2.0 * PI * std::fabs(r_low) * dx);
    for (int iy = 1 - JNG; iy <= ny + JNG; ++iy) {
      if (std::lround(2.0 * r_low / dy) == -1) rt[iy + JNG] = 1.0 / (PI * ((0.5 * dy) * (0.5 * dy)));
      else rt[iy + JNG] = 1.0 / (PI * std::fabs((r_low + dy) * (r_low + dy) - r_low * r_low));
      xt[iy + JNG] = 1.0 / (2.0 * PI * std::fabs(r_low + dy) * dx);
      r_low = r_low + dy;
    }
    for (int iy = 1 - JNG; iy <= ny + JNG; ++iy) {
      vol[iy + JNG] = rt[iy + JNG] / dx;
      ratio[iy + JNG] = xt[iy + JNG] / xt[iy + JNG - 1];
    }
  }
  const double fac = SHAPE_FAC;   // particles.F90:145-153
  PushConst P;
  P.g = g;
  P.exm = (const cplx*)fields6[0]; P.erm = (const cplx*)fields6[1]; P.etm = (const cplx*)fields6[2];
  P.bxm = (const cplx*)fields6[3]; P.brm = (const cplx*)fields6[4]; P.btm = (const cplx*)fields6[5];
  P.jx = (double*)j3[0]; P.jr = (double*)j3[1]; P.jt = (double*)j3[2];
  P.tab = t.data(); P.ntab = ntab;
  P.x_grid_min_local = x_grid_min_local;
  P.y_grid_min_local = y_grid_min_local;
  P.idx = 1.0 / dx; P.idy = 1.0 / dy; P.idt = 1.0 / dt;
  P.dtco2 = C_LIGHT * (dt / 2.0);
  const double dtfac = 0.5 * dt * fac;
  P.part_mc = C_LIGHT * mass;
  P.ipart_mc = 1.0 / P.part_mc;
  P.cmratio = charge * dtfac * P.ipart_mc;
  P.ccmratio = C_LIGHT * P.cmratio;
  P.q_fac = charge * fac;
  P.deposit = zero_current ? 0 : 1;
  P.hc_push = hc_push ? 1 : 0;
  P.taylor_switch = taylor_switch;   // 1.0e-4 (particles.F90:593) unless the conditioning test moves it
  P.hc_alpha = 0.5 * charge * dt / mass;
  const dim3 grid((unsigned)((n + 127) / 128)), block(128);
  // push variant 4 (push_shapes.cuh) where asked for, and always in the builds of the other particle shapes
#define EMUL_V0(MM)                                                                                                   \
  do {                                                                                                                \
    if (CYL_SHAPE != 0 || g_push_generic)                                                                             \
      emul_launch(k_push_generic<MM>, grid, block, P, soa[0], soa[1], soa[2], soa[3], soa[4], soa[5], (const double*)soa[6], n); \
    else                                                                                                              \
      emul_launch(k_push_v0<MM>, grid, block, P, soa[0], soa[1], soa[2], soa[3], soa[4], soa[5], (const double*)soa[6], n); \
  } while (0)
  switch (M) {
    case 1: EMUL_V0(1); break;
    case 2: EMUL_V0(2); break;
    case 3: EMUL_V0(3); break;
    case 4: EMUL_V0(4); break;
    case 5: EMUL_V0(5); break;
    case 6: EMUL_V0(6); break;
    default: return 2;
  }
#undef EMUL_V0
  return 0;
}

// current_bcs_r_min_final (bcs.cu::do_r_min_final) on the three J arrays
EMUL_API void emul_r_min_final(int nx, int ny, int M, void* const* j3) {
  Geom g;
  g.nx = nx; g.ny = ny; g.M = M;
  g.SX = nx + 2 * NG; g.SY = ny + 2 * NG;
  g.plane = (size_t)g.SX * g.SY;
  emul_launch(k_r_min_final, dim3((g.SX + 127) / 128, M), dim3(128), g, (cplx*)j3[0], (cplx*)j3[1], (cplx*)j3[2]);
}

// efield_bcs (which = 0) / bfield_bcs (which = 1, mpi_only = 0) on ONE slab: bcs.cu::do_efield_bcs / do_bfield_bcs
// -- the x halo with the one-row shift of the r-staggered arrays, then edge_bcs with its operator tables
// (conducting walls first, then clamp / zero-gradient by boundary kind), restated here.
EMUL_API void emul_field_bcs(int which, int nx, int ny, int M, void* const* f3, const int32_t* bc_field) {
  Geom g;
  g.nx = nx; g.ny = ny; g.M = M;
  g.SX = nx + 2 * NG; g.SY = ny + 2 * NG;
  g.plane = (size_t)g.SX * g.SY;
  static const int STAG_X[6] = {0, 1, 1, 1, 0, 0}, STAG_Y[6] = {1, 0, 1, 0, 1, 0};   // setup.F90:126-136
  const int base = which ? 3 : 0;
  Halo3 h;
  for (int k = 0; k < 3; ++k) h.f[k] = (cplx*)f3[k];
  h.skip[0] = which ? 0 : 1; h.skip[1] = which ? 1 : 0; h.skip[2] = which ? 0 : 1;
  if (bc_field[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC) {   // the slab is its own neighbour: ghosts <- interior edges
    const size_t elems = (size_t)3 * g.M * g.SY * NG;
    std::vector<cplx> sl(elems), sr(elems);
    const dim3 grd((g.SY * NG + 127) / 128, g.M, 3);
    emul_launch(k_halo_pack, grd, dim3(128), g, h, sl.data(), sr.data(), 0, elems);
    emul_launch(k_halo_unpack, grd, dim3(128), g, h, (const cplx*)sr.data(), (const cplx*)sl.data(), 0, elems);
  }
  Tri t;
  for (int k = 0; k < 3; ++k) t.f[k] = (cplx*)f3[k];
  auto apply = [&](int bd, const int* op) {
    if (op[0] == OP_NONE && op[1] == OP_NONE && op[2] == OP_NONE) return;
    if (bc_field[bd] == CYLGPU_BC_PERIODIC) return;
    for (int k = 0; k < 3; ++k) {
      t.op[k] = op[k];
      t.stag[k] = (bd == CYLGPU_BD_Y_MAX) ? STAG_Y[base + k] : STAG_X[base + k];
    }
    if (bd == CYLGPU_BD_Y_MAX) emul_launch(k_edge_y, dim3((g.SX + 127) / 128, g.M, 3), dim3(128), g, t);
    else emul_launch(k_edge_x, dim3((g.SY + 127) / 128, g.M, 3), dim3(128), g, t, bd);
  };
  const int ecx[3] = {OP_CLAMP, OP_ZEROGRAD, OP_ZEROGRAD}, ecy[3] = {OP_ZEROGRAD, OP_CLAMP, OP_ZEROGRAD};
  const int bcx[3] = {OP_ZEROGRAD, OP_CLAMP, OP_CLAMP}, bcy[3] = {OP_CLAMP, OP_ZEROGRAD, OP_CLAMP};
  const int* cx = which ? bcx : ecx;
  const int* cy = which ? bcy : ecy;
  auto conduct = [&](int bc, const int* c3, int* op) {
    op[0] = op[1] = op[2] = OP_NONE;
    if (bc == CYLGPU_BC_CONDUCT) { op[0] = c3[0]; op[1] = c3[1]; op[2] = c3[2]; }
  };
  auto general = [&](int bc, int* op) {
    op[0] = op[1] = op[2] = OP_NONE;
    if (bc == CYLGPU_BC_CLAMP || bc == CYLGPU_BC_SIMPLE_LASER || bc == CYLGPU_BC_SIMPLE_OUTFLOW) op[0] = op[1] = op[2] = OP_CLAMP;
    if (bc == CYLGPU_BC_ZERO_GRADIENT || bc == CYLGPU_BC_CPML_LASER || bc == CYLGPU_BC_CPML_OUTFLOW) op[0] = op[1] = op[2] = OP_ZEROGRAD;
  };
  int op[3];
  for (int bd : {(int)CYLGPU_BD_X_MIN, (int)CYLGPU_BD_X_MAX}) { conduct(bc_field[bd], cx, op); apply(bd, op); }
  conduct(bc_field[CYLGPU_BD_Y_MAX], cy, op);
  apply(CYLGPU_BD_Y_MAX, op);
  for (int bd : {(int)CYLGPU_BD_X_MIN, (int)CYLGPU_BD_X_MAX, (int)CYLGPU_BD_Y_MAX}) { general(bc_field[bd], op); apply(bd, op); }
}

// bfield_bcs(mpi_only = true) on one periodic slab: the halo alone (Bxm, Brm with its row shift, Btm)
EMUL_API void emul_bfield_halo(int nx, int ny, int M, void* const* b3) {
  Geom g;
  g.nx = nx; g.ny = ny; g.M = M;
  g.SX = nx + 2 * NG; g.SY = ny + 2 * NG;
  g.plane = (size_t)g.SX * g.SY;
  Halo3 h;
  for (int k = 0; k < 3; ++k) h.f[k] = (cplx*)b3[k];
  h.skip[0] = 0; h.skip[1] = 1; h.skip[2] = 0;
  const size_t elems = (size_t)3 * g.M * g.SY * NG;
  std::vector<cplx> sl(elems), sr(elems);
  const dim3 hg((g.SY * NG + 127) / 128, g.M, 3);
  emul_launch(k_halo_pack, hg, dim3(128), g, h, sl.data(), sr.data(), 0, elems);
  emul_launch(k_halo_unpack, hg, dim3(128), g, h, (const cplx*)sr.data(), (const cplx*)sl.data(), 0, elems);
}

// bfield_final_bcs on ONE slab (bcs.cu::do_bfield_final_bcs_device): bfield_bcs, the laser / outflow line
// updates on x_min, x_max and r_max (or zero_b), then the halo.  f15: the 15 mode arrays in field-id order;
// snaps12: the boundary snapshots in snapshot-id order; src4: source1/2 of x_min then x_max, ny + 1 values each.
extern "C" void emul_field_bcs(int which, int nx, int ny, int M, void* const* f3, const int32_t* bc_field);
EMUL_API void emul_bfield_final_bcs(int nx, int ny, int M, void* const* f15, const void* const* snaps12,
                                    const double* const* src4, const int32_t* bc_field, double dx, double dy,
                                    double dt, double y_grid_min_local) {
  Geom g;
  g.nx = nx; g.ny = ny; g.M = M;
  g.SX = nx + 2 * NG; g.SY = ny + 2 * NG;
  g.plane = (size_t)g.SX * g.SY;
  emul_field_bcs(1, nx, ny, M, f15 + 3, bc_field);
  FieldSet F;
  F.exm = (cplx*)f15[0]; F.erm = (cplx*)f15[1]; F.etm = (cplx*)f15[2];
  F.bxm = (cplx*)f15[3]; F.brm = (cplx*)f15[4]; F.btm = (cplx*)f15[5];
  F.jxm = (cplx*)f15[6]; F.jrm = (cplx*)f15[7]; F.jtm = (cplx*)f15[8];
  F.bxo = (const cplx*)f15[9]; F.bro = (const cplx*)f15[10]; F.bto = (const cplx*)f15[11];
  F.jxo = (const cplx*)f15[12]; F.jro = (const cplx*)f15[13]; F.jto = (const cplx*)f15[14];
  const dim3 grd((g.ny + 1 + 127) / 128, g.M);
  auto snap = [&](int k) { return (const cplx*)snaps12[k]; };
  int b = bc_field[CYLGPU_BD_X_MIN];
  if (b == CYLGPU_BC_SIMPLE_LASER || b == CYLGPU_BC_SIMPLE_OUTFLOW)
    emul_launch(k_outflow_x, grd, dim3(128), g, F, snap(1), snap(2), snap(3), snap(4), snap(5), src4[0], src4[1], 0, dx, dy,
                dt, y_grid_min_local, g_reference_quirks);
  b = bc_field[CYLGPU_BD_X_MAX];
  if (b == CYLGPU_BC_SIMPLE_LASER || b == CYLGPU_BC_SIMPLE_OUTFLOW)
    emul_launch(k_outflow_x, grd, dim3(128), g, F, snap(7), snap(8), snap(9), snap(10), snap(11), src4[2], src4[3], 1, dx,
                dy, dt, y_grid_min_local, g_reference_quirks);
  if (bc_field[CYLGPU_BD_Y_MAX] == CYLGPU_BC_SIMPLE_OUTFLOW) {
    emul_launch(k_outflow_r_max, dim3((g.nx + 1 + 127) / 128, g.M), dim3(128), g, F, 1, g.nx - 1, 1, g.nx, 0, dx, dy, dt,
                y_grid_min_local, g_reference_quirks);
  } else if (bc_field[CYLGPU_BD_Y_MAX] == CYLGPU_BC_ZERO_B) {
    emul_launch(k_zero_b_rmax, dim3((g.SX + 127) / 128, g.M), dim3(128), g, F.bxm, F.brm, F.btm);
  }
  if (bc_field[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC) {   // bfield_bcs(mpi_only): the halo alone
    Halo3 h;
    h.f[0] = F.bxm; h.f[1] = F.brm; h.f[2] = F.btm;
    h.skip[0] = 0; h.skip[1] = 1; h.skip[2] = 0;
    const size_t elems = (size_t)3 * g.M * g.SY * NG;
    std::vector<cplx> sl(elems), sr(elems);
    const dim3 hg((g.SY * NG + 127) / 128, g.M, 3);
    emul_launch(k_halo_pack, hg, dim3(128), g, h, sl.data(), sr.data(), 0, elems);
    emul_launch(k_halo_unpack, hg, dim3(128), g, h, (const cplx*)sr.data(), (const cplx*)sl.data(), 0, elems);
  }
}

// current_finish without smoothing on ONE slab (bcs.cu::current_bcs_impl + the J halo): reflection of the ghost
// currents at reflecting walls, the additive ghost exchange and the halo -- as two messages (modes 1 then 0) or as
// the product's merged single message (mode 2).  periodic: the slab is its own x neighbour.
EMUL_API void emul_current_finish(int nx, int ny, int M, void* const* j3, const int32_t* bca, const int32_t* bc_field,
                                  double dy, double y_grid_min_local, int merged) {
  Geom g;
  g.nx = nx; g.ny = ny; g.M = M;
  g.SX = nx + 2 * NG; g.SY = ny + 2 * NG;
  g.plane = (size_t)g.SX * g.SY;
  cplx *jx = (cplx*)j3[0], *jr = (cplx*)j3[1], *jt = (cplx*)j3[2];
  if (bca[CYLGPU_BD_X_MIN] == CYLGPU_BC_REFLECT) emul_launch(k_jreflect_x, dim3((g.SY + 127) / 128, g.M, 3), dim3(128), g, jx, jr, jt, 0);
  if (bca[CYLGPU_BD_X_MAX] == CYLGPU_BC_REFLECT) emul_launch(k_jreflect_x, dim3((g.SY + 127) / 128, g.M, 3), dim3(128), g, jx, jr, jt, 1);
  if (bca[CYLGPU_BD_Y_MAX] == CYLGPU_BC_REFLECT)
    emul_launch(k_jreflect_y, dim3((g.SX + 127) / 128, g.M, 3), dim3(128), g, jx, jr, jt, dy, y_grid_min_local);
  const bool ring = bc_field[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC || bca[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC;
  const bool to = ring && bca[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC;            // to_l == to_r on one slab
  const bool fill = ring && bc_field[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC;     // fill_l == fill_r
  Halo3 h;
  h.f[0] = jx; h.f[1] = jr; h.f[2] = jt;
  h.skip[0] = h.skip[1] = h.skip[2] = 0;
  const size_t elems = (size_t)3 * g.M * g.SY * NG;
  const dim3 grd((g.SY * NG + 127) / 128, g.M, 3);
  auto exchange = [&](int mode) {
    std::vector<cplx> sl(2 * elems), sr(2 * elems);
    emul_launch(k_halo_pack, grd, dim3(128), g, h, sl.data(), sr.data(), mode, elems);
    emul_launch(k_halo_unpack, grd, dim3(128), g, h, (const cplx*)sr.data(), (const cplx*)sl.data(), mode, elems);
  };
  if (merged && to && fill) { exchange(2); return; }
  if (to) exchange(1);
  if (fill) exchange(0);
}

// setup_field_boundaries (bcs.cu::do_snapshot): the twelve x_min / x_max boundary snapshots of E and B
EMUL_API void emul_snapshot(int nx, int ny, int M, void* const* f15, void* const* snaps12) {
  Geom g;
  g.nx = nx; g.ny = ny; g.M = M;
  g.SX = nx + 2 * NG; g.SY = ny + 2 * NG;
  g.plane = (size_t)g.SX * g.SY;
  FieldSet F;
  F.exm = (cplx*)f15[0]; F.erm = (cplx*)f15[1]; F.etm = (cplx*)f15[2];
  F.bxm = (cplx*)f15[3]; F.brm = (cplx*)f15[4]; F.btm = (cplx*)f15[5];
  F.jxm = (cplx*)f15[6]; F.jrm = (cplx*)f15[7]; F.jtm = (cplx*)f15[8];
  F.bxo = (const cplx*)f15[9]; F.bro = (const cplx*)f15[10]; F.bto = (const cplx*)f15[11];
  F.jxo = (const cplx*)f15[12]; F.jro = (const cplx*)f15[13]; F.jto = (const cplx*)f15[14];
  Snaps S;
  for (int k = 0; k < CYLGPU_NSNAPS; ++k) S.s[k] = (cplx*)snaps12[k];
  emul_launch(k_snapshot, dim3((g.SY + 127) / 128, g.M), dim3(128), g, F, S);
}

// shift_fields of the moving window on ONE non-periodic slab (bcs.cu::do_shift_fields): the nine arrays move one
// cell towards x_min (out of place), the x_max columns are refilled from the snapshots
EMUL_API void emul_shift_fields(int nx, int ny, int M, void* const* f15, void* const* snaps12) {
  Geom g;
  g.nx = nx; g.ny = ny; g.M = M;
  g.SX = nx + 2 * NG; g.SY = ny + 2 * NG;
  g.plane = (size_t)g.SX * g.SY;
  const size_t nel = g.plane * g.M;
  std::vector<cplx> spare(nel);
  const int blocks = (int)((nel + 255) / 256 < 148 * 16 ? (nel + 255) / 256 : 148 * 16);
  for (int k = 0; k < 9; ++k) {
    emul_launch(k_shift_x, dim3((unsigned)blocks), dim3(256), g, (const cplx*)f15[k], spare.data());
    memcpy(f15[k], spare.data(), nel * sizeof(cplx));     // the product swaps the pointers instead
  }
  FieldSet F;
  F.exm = (cplx*)f15[0]; F.erm = (cplx*)f15[1]; F.etm = (cplx*)f15[2];
  F.bxm = (cplx*)f15[3]; F.brm = (cplx*)f15[4]; F.btm = (cplx*)f15[5];
  F.jxm = (cplx*)f15[6]; F.jrm = (cplx*)f15[7]; F.jtm = (cplx*)f15[8];
  F.bxo = (const cplx*)f15[9]; F.bro = (const cplx*)f15[10]; F.bto = (const cplx*)f15[11];
  F.jxo = (const cplx*)f15[12]; F.jro = (const cplx*)f15[13]; F.jto = (const cplx*)f15[14];
  Snaps S;
  for (int k = 0; k < CYLGPU_NSNAPS; ++k) S.s[k] = (cplx*)snaps12[k];
  emul_launch(k_window_fill_xmax, dim3((g.SY + 127) / 128, g.M), dim3(128), g, F, S);
}

// smooth_current on ONE slab (bcs.cu::do_smooth_current): the strided compensated binomial filter on the three J
// arrays with two ping-pong work sets, the halo of the work set before every pass (self halo when periodic).
EMUL_API void emul_smooth_current(int nx, int ny, int M, void* const* j3, int its, int comp_its, int nstrides,
                                  const int32_t* strides_in, int periodic) {
  Geom g;
  g.nx = nx; g.ny = ny; g.M = M;
  g.SX = nx + 2 * NG; g.SY = ny + 2 * NG;
  g.plane = (size_t)g.SX * g.SY;
  std::vector<int> strides(strides_in, strides_in + nstrides);
  if (strides.empty()) strides.push_back(1);
  const size_t nel = g.plane * g.M;
  cplx* J[3] = {(cplx*)j3[0], (cplx*)j3[1], (cplx*)j3[2]};
  std::vector<cplx> wk[2][3];
  for (int s = 0; s < 2; ++s)
    for (int k = 0; k < 3; ++k) wk[s][k].assign(J[k], J[k] + nel);
  double alpha = 0.5;
  const double beta = (1.0 - alpha) * 0.25;
  int cur = 0;
  const dim3 grd((g.nx + 127) / 128, g.ny, 3 * g.M);
  for (int it = 1; it <= its + comp_its; ++it) {
    for (int stride : strides) {
      cplx* W[3] = {wk[cur][0].data(), wk[cur][1].data(), wk[cur][2].data()};
      cplx* D[3] = {wk[cur ^ 1][0].data(), wk[cur ^ 1][1].data(), wk[cur ^ 1][2].data()};
      if (periodic) {
        Halo3 h;
        for (int k = 0; k < 3; ++k) { h.f[k] = W[k]; h.skip[k] = 0; }
        const size_t elems = (size_t)3 * g.M * g.SY * NG;
        std::vector<cplx> sl(elems), sr(elems);
        const dim3 hg((g.SY * NG + 127) / 128, g.M, 3);
        emul_launch(k_halo_pack, hg, dim3(128), g, h, sl.data(), sr.data(), 0, elems);
        emul_launch(k_halo_unpack, hg, dim3(128), g, h, (const cplx*)sr.data(), (const cplx*)sl.data(), 0, elems);
      }
      Tri3 src{{W[0], W[1], W[2]}}, dst{{D[0], D[1], D[2]}};
      emul_launch(k_smooth, grd, dim3(128), g, src, dst, alpha, beta, stride);
      cur ^= 1;
    }
    if (it > its) alpha = (double)its * 0.5 + 1.0;
  }
  Tri3 src{{wk[cur][0].data(), wk[cur][1].data(), wk[cur][2].data()}}, dst{{J[0], J[1], J[2]}};
  emul_launch(k_copy_interior, grd, dim3(128), g, src, dst);
}

// update_e_field / update_b_field of one slab without boundary conditions (fields.cu::launch_update_e / _b).
// f9: exm erm etm bxm brm btm jxm jrm jtm, complex with ghosts, updated in place.
EMUL_API void emul_update_field(int which, int nx, int ny, int M, void* const* f9, double dx, double dy, double dt,
                                double y_grid_min_local, void* const* bold3) {
  Geom g;
  g.nx = nx; g.ny = ny; g.M = M;
  g.SX = nx + 2 * NG; g.SY = ny + 2 * NG;
  g.plane = (size_t)g.SX * g.SY;
  cplx* f[9];
  for (int k = 0; k < 9; ++k) f[k] = (cplx*)f9[k];
  const dim3 blk(128);
  FieldRecips R;   // fields.cu::field_recips
  R.idx = 1.0 / dx; R.idy = 1.0 / dy; R.ieps0 = 1.0 / EPSILON0;
  if (which == 0) {
    emul_launch(k_update_e_bulk, dim3((g.nx + 1 + 127) / 128, g.ny, g.M), blk, g, f[0], f[1], f[2], (const cplx*)f[3],
                (const cplx*)f[4], (const cplx*)f[5], (const cplx*)f[6], (const cplx*)f[7], (const cplx*)f[8], R, dy, dt,
                y_grid_min_local, 0, g.nx);
    emul_launch(k_update_e_axis, dim3((g.SX + 127) / 128, g.M, NG), blk, g, f[0], f[1], f[2], (const cplx*)f[5],
                (const cplx*)f[6], dy, dt);
  } else {
    // which == 2: fields.cu::launch_update_b(save_old = true), the b*_old = b* copies of update_eb_fields_half fused
    cplx* bo[3] = {nullptr, nullptr, nullptr};
    if (which == 2) {
      for (int k = 0; k < 3; ++k) bo[k] = (cplx*)bold3[k];
      emul_launch(k_copy_b_old_rim, dim3((g.SX + 127) / 128, g.SY, g.M), blk, g, (const cplx*)f[3], (const cplx*)f[4],
                  (const cplx*)f[5], bo[0], bo[1], bo[2], 0, g.nx);
    }
    if (g.ny > 1) {
      const dim3 grd((g.nx + 1 + 127) / 128, g.ny - 1, g.M);
      if (which == 2)
        emul_launch(k_update_b_bulk<true>, grd, blk, g, f[3], f[4], f[5], (const cplx*)f[0], (const cplx*)f[1],
                    (const cplx*)f[2], bo[0], bo[1], bo[2], R, dy, dt, y_grid_min_local, 0, g.nx);
      else
        emul_launch(k_update_b_bulk<false>, grd, blk, g, f[3], f[4], f[5], (const cplx*)f[0], (const cplx*)f[1],
                    (const cplx*)f[2], bo[0], bo[1], bo[2], R, dy, dt, y_grid_min_local, 0, g.nx);
    }
    emul_launch(k_update_b_axis, dim3((g.SX + 127) / 128, g.M, NG), blk, g, f[3], f[4], f[5], (const cplx*)f[0],
                (const cplx*)f[2], dx, dy, dt);
  }
}

// particle_bcs of one species on one slab: particles.cu::make_bcs_const + k_pbcs_classify.  The SoA arrays
// are modified in place (reflection, periodic shift); hole_list / hole_flag (capacity n) receive the leavers
// and their fate (1 left, 2 right, 3 gone); counts[0..3] = holes, left, right, gone.
EMUL_API void emul_pbcs_classify(double* const* soa, int64_t n, const int32_t* bc_particle, double x_min, double x_max,
                                 double x_min_local, double x_max_local, double y_max, double dx, double dy,
                                 int x_min_boundary, int x_max_boundary, uint32_t* hole_list, uint8_t* hole_flag,
                                 unsigned long long* counts) {
  BcsConst B;
  B.x_min = x_min; B.x_max = x_max;
  B.x_min_local = x_min_local; B.x_max_local = x_max_local;
  B.y_max = y_max;
  double boundary_shift = dx * (double)((1 + PNG + 0) / 2);   // boundary.F90:1561-1563
  B.x_min_outer = B.x_min - boundary_shift;
  B.x_max_outer = B.x_max + boundary_shift;
  boundary_shift = dy * (double)((1 + PNG + 0) / 2);
  B.y_max_outer = B.y_max + boundary_shift;
  B.x_shift = B.x_max - B.x_min;
  B.y_max2_inside = B.y_max * B.y_max * (1.0 - 1.0e-14);
  B.x_min_boundary = x_min_boundary;
  B.x_max_boundary = x_max_boundary;
  B.remove_x = -1.0e300;
  B.x_only = 0;
  for (int k = 0; k < 4; ++k) B.bc[k] = bc_particle[k];
  unsigned long long cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (n > 0)
    emul_launch(k_pbcs_classify, dim3((unsigned)((n + 255) / 256)), dim3(256), B, soa[0], soa[1], soa[2], soa[3], soa[4],
                soa[5], hole_list, hole_flag, (unsigned long long*)cnt, n);
  for (int k = 0; k < 4; ++k) counts[k] = cnt[k];
}

// window_insert.cu::do_insert_particles_device for the x_max slab: returns the number of particles
// written (7 SoA arrays of capacity cap each), or -1 if cap is too small.
EMUL_API int64_t emul_insert_column(int ny, int isp, double x_grid_max, double npart_per_cell_real, const double* density_in,
                           const double* temperature, const double* drift, double dmin, double dmax, uint64_t seed,
                           uint64_t column, double dx, double dy, double y_grid_min_local, double mass, int64_t cap,
                           double* const* soa) {
  const int nrow = ny + 2;
  const int64_t npart_per_cell = (int64_t)std::floor(npart_per_cell_real);
  const double npart_frac = npart_per_cell_real - (double)npart_per_cell;
  const ColumnStream rs = column_stream(seed, isp, column);
  std::vector<double> stage((size_t)7 * nrow + (size_t)(ny + 1));
  double* prof = stage.data();
  int64_t* row_start = reinterpret_cast<int64_t*>(stage.data() + (size_t)7 * nrow);
  for (int iy = 0; iy < nrow; ++iy) {
    double d = density_in[iy];
    if (d > dmax) d = dmax;
    if (d < dmin) d = 0.0;
    prof[iy] = d;
  }
  for (int i = 0; i < 3 * nrow; ++i) { prof[nrow + i] = temperature[i]; prof[4 * nrow + i] = drift[i]; }
  int64_t total = 0;
  row_start[0] = 0;
  for (int iy = 1; iy <= ny; ++iy) {
    int64_t ncell = 0;
    if (!(prof[iy] < dmin)) {
      int64_t n_frac = 0;
      if (npart_frac > 0.0 && rs.cell_uniform((uint32_t)iy) < npart_frac) n_frac = 1;
      ncell = npart_per_cell + n_frac;
    }
    total += ncell;
    row_start[iy] = total;
  }
  if (total > cap) return -1;
  ColumnArgs a;
  a.rs = rs;
  a.prof = prof;
  a.row_start = row_start;
  a.x = soa[0]; a.y = soa[1]; a.z = soa[2]; a.px = soa[3]; a.py = soa[4]; a.pz = soa[5]; a.w = soa[6];
  a.base = 0;
  a.base_dev = nullptr;
  a.ny = ny;
  a.iy_global_offset = 0;
  a.dx = dx; a.dy = dy;
  a.x0 = x_grid_max + 0.5 * dx;
  a.y_grid_min_local = y_grid_min_local;
  a.mass = mass;
  emul_launch(k_insert_column, dim3((unsigned)ny), dim3(128), a);
  return total;
}

// particle_bcs with device-resident counts for a chain (or periodic ring) of `nslab` x-slabs of one species:
// particles.cu::pbcs_species_fast restated per slab -- k_pbcs_classify_dev, compact_dev's five kernels, the
// fixed-size message [7-double header][xcap slots] handed to the neighbour, k_unpack_dev, k_bump_count -- with
// every count living in "device" memory (n_dev, the plan, the statistics).  soa: nslab x 7 arrays of capacity cap;
// n: particles per slab in / out; bounds: x_min_local, x_max_local per slab; pstats: nslab x PST_N out.
EMUL_API int emul_pbcs_fast(int nslab, double* const* soa, int64_t* n, int64_t cap, const int32_t* bc_particle,
                            double x_min, double x_max, const double* x_min_local, const double* x_max_local,
                            double y_max, double dx, double dy, int periodic, long long xcap, int64_t* pstats_out,
                            double remove_x, int x_only) {
  const size_t msg = (size_t)(7 * xcap + XHDR);
  std::vector<std::vector<double>> send_l(nslab, std::vector<double>(msg, -7.0)), send_r(nslab, std::vector<double>(msg, -7.0));
  std::vector<std::vector<int64_t>> ndev(nslab, std::vector<int64_t>(1 + PST_N, 0));
  const dim3 lg(LEAVER_GRID > 8 ? 8 : LEAVER_GRID), lb(256);   // (a smaller grid: the loops are grid-stride)
  for (int k = 0; k < nslab; ++k) {
    BcsConst B;
    B.x_min = x_min; B.x_max = x_max;
    B.x_min_local = x_min_local[k]; B.x_max_local = x_max_local[k];
    B.y_max = y_max;
    double boundary_shift = dx * (double)((1 + PNG + 0) / 2);
    B.x_min_outer = B.x_min - boundary_shift;
    B.x_max_outer = B.x_max + boundary_shift;
    boundary_shift = dy * (double)((1 + PNG + 0) / 2);
    B.y_max_outer = B.y_max + boundary_shift;
    B.x_shift = B.x_max - B.x_min;
    B.y_max2_inside = B.y_max * B.y_max * (1.0 - 1.0e-14);
    B.x_min_boundary = (k == 0);
    B.x_max_boundary = (k == nslab - 1);
    B.remove_x = (k == 0) ? remove_x : -1.0e300;   // remove_particles riding on the classification (x_min slab)
    B.x_only = x_only;
    for (int q = 0; q < 4; ++q) B.bc[q] = bc_particle[q];
    const bool has_l = k > 0 || periodic, has_r = k < nslab - 1 || periodic;
    Soa s;
    for (int q = 0; q < 7; ++q) s.d[q] = soa[7 * k + q];
    // the launches cover an upper bound of the count, as the host's do
    const int64_t bound = n[k] + 100;
    std::vector<uint32_t> hole_list((size_t)bound), lowhole((size_t)bound), hightail((size_t)bound);
    std::vector<uint8_t> flag((size_t)bound), tailmark((size_t)bound, 1);
    unsigned long long cnt[16] = {0};
    CompactPlan plan;
    int64_t* n_dev = ndev[k].data();
    int64_t* pst = n_dev + 1;
    *n_dev = n[k];
    emul_launch(k_pbcs_classify_dev, dim3((unsigned)((bound + 255) / 256)), dim3(256), B, s.d[0], s.d[1], s.d[2], s.d[3],
                s.d[4], s.d[5], hole_list.data(), flag.data(), (unsigned long long*)cnt, (const int64_t*)n_dev);
    emul_launch(k_plan_compact, dim3(1), dim3(1), (const unsigned long long*)cnt, cnt + 8, n_dev, &plan, pst, xcap, 0,
                has_l ? send_l[k].data() : (double*)nullptr, has_r ? send_r[k].data() : (double*)nullptr);
    emul_launch(k_clear_tail, lg, lb, tailmark.data(), (const CompactPlan*)&plan);
    emul_launch(k_collect_dev, lg, lb, s, (const uint32_t*)hole_list.data(), (const uint8_t*)flag.data(),
                (const CompactPlan*)&plan, send_l[k].data() + XHDR, send_r[k].data() + XHDR,
                (has_l || has_r) ? xcap : 0LL, lowhole.data(), tailmark.data(), cnt + 8);
    emul_launch(k_tail_keepers_dev, lg, lb, (const uint8_t*)tailmark.data(), (const CompactPlan*)&plan, hightail.data(),
                cnt + 8);
    emul_launch(k_fill_holes_dev, lg, lb, s, (const uint32_t*)lowhole.data(), (const uint32_t*)hightail.data(),
                (const CompactPlan*)&plan, (const unsigned long long*)(cnt + 8));
  }
  // the exchange: what a slab sends left is what its left neighbour receives "from the right", and vice versa
  for (int k = 0; k < nslab; ++k) {
    const int left = k > 0 ? k - 1 : (periodic ? nslab - 1 : -1), right = k < nslab - 1 ? k + 1 : (periodic ? 0 : -1);
    const double* rr = right >= 0 ? send_l[right].data() : nullptr;
    const double* rl = left >= 0 ? send_r[left].data() : nullptr;
    Soa s;
    for (int q = 0; q < 7; ++q) s.d[q] = soa[7 * k + q];
    int64_t* n_dev = ndev[k].data();
    const long long arriving = (rr ? *reinterpret_cast<const long long*>(rr) : 0) + (rl ? *reinterpret_cast<const long long*>(rl) : 0);
    if (*n_dev + arriving > cap) return -1;
    emul_launch(k_unpack_dev, lg, lb, s, (const int64_t*)n_dev, rr, rl);
    emul_launch(k_bump_count, dim3(1), dim3(1), n_dev, n_dev + 1, rr, rl);
    n[k] = *n_dev;
    for (int q = 0; q < PST_N; ++q) pstats_out[k * PST_N + q] = n_dev[1 + q];
  }
  return 0;
}

}  // extern "C"


// ------------------------------------------------------------------------------------------------------
// update_eb_fields_half (phase 0) / update_eb_fields_final (phase 1) of a ROW OF SLABS, in the launch order of
// api.cu (fields_half_body / cylgpu_fields_final) with the halos between the slabs done by the product's pack /
// unpack kernels.  wide = 1: the communication-avoiding order (field_ranges.cuh: ghost columns advanced by the
// slab itself, one closing exchange of E and B; wide = 2 with phase 1: the final phase right before the window's
// shift_fields, no exchange at all); wide = 0: the reference's five exchanges.
// f15: nranks x 15 mode arrays in field-id order; snaps12: nranks x 12 boundary snapshots; src4: source1/2 of
// x_min then x_max.
// ------------------------------------------------------------------------------------------------------
namespace {
struct Row {
  int n, ny, M;
  std::vector<Geom> g;
  std::vector<cplx*> f;   // [rank * 15 + id]
  bool periodic;
  int left(int k) const { return k > 0 ? k - 1 : (periodic ? n - 1 : -1); }
  int right(int k) const { return k < n - 1 ? k + 1 : (periodic ? 0 : -1); }
  bool fill_l(int k) const { return left(k) >= 0; }    // (a non-periodic domain boundary has no neighbour)
  bool fill_r(int k) const { return right(k) >= 0; }
};

// field_mode_bc on `narr` groups of three arrays of every slab (base ids in `bases`, row skips per group)
void row_halo(const Row& R, const int* bases, const int (*skips)[3], int ngroups) {
  const size_t he = (size_t)3 * R.M * (R.ny + 2 * NG) * NG;
  std::vector<std::vector<cplx>> sl(R.n), sr(R.n);
  for (int k = 0; k < R.n; ++k) {
    sl[k].assign(ngroups * he, cplx{0.0, 0.0});
    sr[k].assign(ngroups * he, cplx{0.0, 0.0});
    const dim3 grd((R.g[k].SY * NG + 127) / 128, R.M, 3);
    for (int q = 0; q < ngroups; ++q) {
      Halo3 h;
      for (int a = 0; a < 3; ++a) { h.f[a] = R.f[k * 15 + bases[q] + a]; h.skip[a] = skips[q][a]; }
      emul_launch(k_halo_pack, grd, dim3(128), R.g[k], h, R.fill_l(k) ? sl[k].data() + q * he : (cplx*)nullptr,
                  R.fill_r(k) ? sr[k].data() + q * he : (cplx*)nullptr, 0, he);
    }
  }
  for (int k = 0; k < R.n; ++k) {
    const dim3 grd((R.g[k].SY * NG + 127) / 128, R.M, 3);
    for (int q = 0; q < ngroups; ++q) {
      Halo3 h;
      for (int a = 0; a < 3; ++a) { h.f[a] = R.f[k * 15 + bases[q] + a]; h.skip[a] = skips[q][a]; }
      const cplx* rl = R.fill_l(k) ? sr[R.left(k)].data() + q * he : nullptr;    // my left neighbour's right-going block
      const cplx* rr = R.fill_r(k) ? sl[R.right(k)].data() + q * he : nullptr;
      emul_launch(k_halo_unpack, grd, dim3(128), R.g[k], h, rl, rr, 0, he);
    }
  }
}

// bcs.cu::edge_bcs of slab k (only on the domain boundaries it owns)
void row_edges(const Row& R, int k, int which, const int32_t* bc_field) {
  static const int STAG_X[6] = {0, 1, 1, 1, 0, 0}, STAG_Y[6] = {1, 0, 1, 0, 1, 0};
  const Geom& g = R.g[k];
  const int base = which ? 3 : 0;
  Tri t;
  for (int a = 0; a < 3; ++a) t.f[a] = R.f[k * 15 + base + a];
  auto apply = [&](int bd, const int* op) {
    if (op[0] == OP_NONE && op[1] == OP_NONE && op[2] == OP_NONE) return;
    if (bc_field[bd] == CYLGPU_BC_PERIODIC) return;
    if (bd == CYLGPU_BD_X_MIN && k != 0) return;
    if (bd == CYLGPU_BD_X_MAX && k != R.n - 1) return;
    for (int a = 0; a < 3; ++a) {
      t.op[a] = op[a];
      t.stag[a] = (bd == CYLGPU_BD_Y_MAX) ? STAG_Y[base + a] : STAG_X[base + a];
    }
    if (bd == CYLGPU_BD_Y_MAX) emul_launch(k_edge_y, dim3((g.SX + 127) / 128, g.M, 3), dim3(128), g, t);
    else emul_launch(k_edge_x, dim3((g.SY + 127) / 128, g.M, 3), dim3(128), g, t, bd);
  };
  const int ecx[3] = {OP_CLAMP, OP_ZEROGRAD, OP_ZEROGRAD}, ecy[3] = {OP_ZEROGRAD, OP_CLAMP, OP_ZEROGRAD};
  const int bcx[3] = {OP_ZEROGRAD, OP_CLAMP, OP_CLAMP}, bcy[3] = {OP_CLAMP, OP_ZEROGRAD, OP_CLAMP};
  const int* cx = which ? bcx : ecx;
  const int* cy = which ? bcy : ecy;
  auto conduct = [&](int bc, const int* c3, int* op) {
    op[0] = op[1] = op[2] = OP_NONE;
    if (bc == CYLGPU_BC_CONDUCT) { op[0] = c3[0]; op[1] = c3[1]; op[2] = c3[2]; }
  };
  auto general = [&](int bc, int* op) {
    op[0] = op[1] = op[2] = OP_NONE;
    if (bc == CYLGPU_BC_CLAMP || bc == CYLGPU_BC_SIMPLE_LASER || bc == CYLGPU_BC_SIMPLE_OUTFLOW) op[0] = op[1] = op[2] = OP_CLAMP;
    if (bc == CYLGPU_BC_ZERO_GRADIENT || bc == CYLGPU_BC_CPML_LASER || bc == CYLGPU_BC_CPML_OUTFLOW) op[0] = op[1] = op[2] = OP_ZEROGRAD;
  };
  int op[3];
  for (int bd : {(int)CYLGPU_BD_X_MIN, (int)CYLGPU_BD_X_MAX}) { conduct(bc_field[bd], cx, op); apply(bd, op); }
  conduct(bc_field[CYLGPU_BD_Y_MAX], cy, op);
  apply(CYLGPU_BD_Y_MAX, op);
  for (int bd : {(int)CYLGPU_BD_X_MIN, (int)CYLGPU_BD_X_MAX, (int)CYLGPU_BD_Y_MAX}) { general(bc_field[bd], op); apply(bd, op); }
}
}  // namespace

extern "C" EMUL_API void emul_field_phase_slabs(int phase, int wide, int nranks, const int* nx_each, int ny, int M,
                                     void* const* f15, const void* const* snaps12, const double* const* src4,
                                     const int32_t* bc_field, double dx, double dy, double dt,
                                     double y_grid_min_local) {
  Row R;
  R.n = nranks; R.ny = ny; R.M = M;
  R.periodic = bc_field[CYLGPU_BD_X_MIN] == CYLGPU_BC_PERIODIC;
  for (int k = 0; k < nranks; ++k) {
    Geom g;
    g.nx = nx_each[k]; g.ny = ny; g.M = M;
    g.SX = g.nx + 2 * NG; g.SY = ny + 2 * NG;
    g.plane = (size_t)g.SX * g.SY;
    R.g.push_back(g);
    for (int a = 0; a < 15; ++a) R.f.push_back((cplx*)f15[k * 15 + a]);
  }
  FieldRecips Rc;
  Rc.idx = 1.0 / dx; Rc.idy = 1.0 / dy; Rc.ieps0 = 1.0 / EPSILON0;
  const dim3 blk(128);
  auto F = [&](int k, int id) { return R.f[k * 15 + id]; };
  auto ranges = [&](int k) {
    const bool w = wide != 0 && nranks > 1 && R.g[k].nx >= 2 * NG && (R.fill_l(k) || R.fill_r(k));   // bcs.cu::wide_fields
    return field_ranges(phase == 1 && wide == 2 ? 2 : phase, w, R.fill_l(k), R.fill_r(k), k == 0, k == nranks - 1, R.g[k].nx);
  };
  auto sweep_e = [&](int k, int lo, int hi) {   // fields.cu::launch_update_e
    const Geom& g = R.g[k];
    emul_launch(k_update_e_bulk, dim3((hi - lo + 1 + 127) / 128, g.ny, g.M), blk, g, F(k, 0), F(k, 1), F(k, 2),
                (const cplx*)F(k, 3), (const cplx*)F(k, 4), (const cplx*)F(k, 5), (const cplx*)F(k, 6),
                (const cplx*)F(k, 7), (const cplx*)F(k, 8), Rc, dy, dt, y_grid_min_local, lo, hi);
    emul_launch(k_update_e_axis, dim3((g.SX + 127) / 128, g.M, NG), blk, g, F(k, 0), F(k, 1), F(k, 2),
                (const cplx*)F(k, 5), (const cplx*)F(k, 6), dy, dt);
  };
  auto sweep_b = [&](int k, bool save_old, int lo, int hi) {   // fields.cu::launch_update_b
    const Geom& g = R.g[k];
    if (save_old)
      emul_launch(k_copy_b_old_rim, dim3((g.SX + 127) / 128, g.SY, g.M), blk, g, (const cplx*)F(k, 3), (const cplx*)F(k, 4),
                  (const cplx*)F(k, 5), F(k, 9), F(k, 10), F(k, 11), lo, hi);
    if (g.ny > 1) {
      const dim3 grd((hi - lo + 1 + 127) / 128, g.ny - 1, g.M);
      if (save_old)
        emul_launch(k_update_b_bulk<true>, grd, blk, g, F(k, 3), F(k, 4), F(k, 5), (const cplx*)F(k, 0), (const cplx*)F(k, 1),
                    (const cplx*)F(k, 2), F(k, 9), F(k, 10), F(k, 11), Rc, dy, dt, y_grid_min_local, lo, hi);
      else
        emul_launch(k_update_b_bulk<false>, grd, blk, g, F(k, 3), F(k, 4), F(k, 5), (const cplx*)F(k, 0), (const cplx*)F(k, 1),
                    (const cplx*)F(k, 2), F(k, 9), F(k, 10), F(k, 11), Rc, dy, dt, y_grid_min_local, lo, hi);
    }
    emul_launch(k_update_b_axis, dim3((g.SX + 127) / 128, g.M, NG), blk, g, F(k, 3), F(k, 4), F(k, 5), (const cplx*)F(k, 0),
                (const cplx*)F(k, 2), dx, dy, dt);
  };
  auto outflow = [&](int k, const FieldRanges& X) {   // bcs.cu::do_bfield_final_bcs_device between its two halos
    const Geom& g = R.g[k];
    FieldSet S;
    S.exm = F(k, 0); S.erm = F(k, 1); S.etm = F(k, 2); S.bxm = F(k, 3); S.brm = F(k, 4); S.btm = F(k, 5);
    S.jxm = F(k, 6); S.jrm = F(k, 7); S.jtm = F(k, 8); S.bxo = F(k, 9); S.bro = F(k, 10); S.bto = F(k, 11);
    S.jxo = F(k, 12); S.jro = F(k, 13); S.jto = F(k, 14);
    const dim3 grd((g.ny + 1 + 127) / 128, g.M);
    auto snap = [&](int q) { return (const cplx*)snaps12[k * 12 + q]; };
    if (k == 0) {
      const int b = bc_field[CYLGPU_BD_X_MIN];
      if (b == CYLGPU_BC_SIMPLE_LASER || b == CYLGPU_BC_SIMPLE_OUTFLOW)
        emul_launch(k_outflow_x, grd, dim3(128), g, S, snap(1), snap(2), snap(3), snap(4), snap(5), src4[0], src4[1], 0, dx,
                    dy, dt, y_grid_min_local, g_reference_quirks);
    }
    if (k == nranks - 1) {
      const int b = bc_field[CYLGPU_BD_X_MAX];
      if (b == CYLGPU_BC_SIMPLE_LASER || b == CYLGPU_BC_SIMPLE_OUTFLOW)
        emul_launch(k_outflow_x, grd, dim3(128), g, S, snap(7), snap(8), snap(9), snap(10), snap(11), src4[2], src4[3], 1,
                    dx, dy, dt, y_grid_min_local, g_reference_quirks);
    }
    if (bc_field[CYLGPU_BD_Y_MAX] == CYLGPU_BC_SIMPLE_OUTFLOW) {
      const int ix0 = X.obx_lo < X.obt_lo ? X.obx_lo : X.obt_lo;
      const int ix1 = X.obx_hi > X.obt_hi ? X.obx_hi : X.obt_hi;
      emul_launch(k_outflow_r_max, dim3((ix1 - ix0 + 1 + 127) / 128, g.M), dim3(128), g, S, X.obx_lo, X.obx_hi, X.obt_lo,
                  X.obt_hi, ix0, dx, dy, dt, y_grid_min_local, g_reference_quirks);
    } else if (bc_field[CYLGPU_BD_Y_MAX] == CYLGPU_BC_ZERO_B) {
      emul_launch(k_zero_b_rmax, dim3((g.SX + 127) / 128, g.M), dim3(128), g, S.bxm, S.brm, S.btm);
    }
  };
  const int baseE[1] = {0}, baseB[1] = {3}, baseEB[2] = {0, 3};
  const int skipE[1][3] = {{1, 0, 1}}, skipB[1][3] = {{0, 1, 0}}, skipEB[2][3] = {{1, 0, 1}, {0, 1, 0}};
  if (wide) {
    if (phase == 0) {
      for (int k = 0; k < nranks; ++k) {
        const FieldRanges X = ranges(k);
        sweep_e(k, X.e_lo, X.e_hi);
        row_edges(R, k, 0, bc_field);
        sweep_b(k, true, X.b_lo, X.b_hi);
      }
      row_halo(R, baseEB, skipEB, 2);
    } else {
      for (int k = 0; k < nranks; ++k) {
        const FieldRanges X = ranges(k);
        sweep_b(k, false, X.b_lo, X.b_hi);
        row_edges(R, k, 1, bc_field);
        outflow(k, X);
        sweep_e(k, X.e_lo, X.e_hi);
        row_edges(R, k, 0, bc_field);
      }
      if (wide != 2) row_halo(R, baseEB, skipEB, 2);   // wide == 2: shift_fields follows and exchanges (api.cu to_shift)
    }
    return;
  }
  // the reference's order: every slab sweeps, the row exchanges, every slab fills its domain edges
  auto all = [&](auto fn) { for (int k = 0; k < nranks; ++k) fn(k); };
  if (phase == 0) {
    all([&](int k) { sweep_e(k, 0, R.g[k].nx); });
    row_halo(R, baseE, skipE, 1);
    all([&](int k) { row_edges(R, k, 0, bc_field); });
    all([&](int k) { sweep_b(k, true, 0, R.g[k].nx); });
    row_halo(R, baseB, skipB, 1);
  } else {
    all([&](int k) { sweep_b(k, false, 0, R.g[k].nx); });
    row_halo(R, baseB, skipB, 1);
    all([&](int k) { row_edges(R, k, 1, bc_field); });
    all([&](int k) { outflow(k, ranges(k)); });
    row_halo(R, baseB, skipB, 1);
    all([&](int k) { sweep_e(k, 0, R.g[k].nx); });
    row_halo(R, baseE, skipE, 1);
    all([&](int k) { row_edges(R, k, 0, bc_field); });
  }
}
