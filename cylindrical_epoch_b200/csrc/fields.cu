// fields.cu -- per-mode Yee-type FDTD half-step updates with the on-axis treatment.
// Replaces update_e_field (fields.f90:53-182) and update_b_field (fields.f90:186-312).
//
// Layout: complex128 (double2) arrays, x fastest, then r, then mode -- identical to the
// Fortran arrays, so x-adjacent threads read adjacent 16-byte elements (coalesced 512 B
// per warp and component).  The stencils only reach +-1 in x and r; the second read of a
// neighbour is served by L1/L2, so the HBM traffic is the algorithmic 12 (E) / 9 (B)
// arrays per sweep.
#include "ctx.cuh"

namespace cylgpu {

#include "field_kernels.cuh"

static FieldRecips field_recips(const cylgpu_ctx* c) {
  FieldRecips R;
  R.idx = 1.0 / c->cfg.dx; R.idy = 1.0 / c->cfg.dy; R.ieps0 = 1.0 / EPSILON0;
  return R;
}

// ix_lo..ix_hi: the columns of the bulk sweep (the reference's 0..nx, or field_ranges.cuh); the axis rows and
// the mirror rows always run over the full extent, as in the reference
int launch_update_e(cylgpu_ctx* c, int ix_lo, int ix_hi) {
  const Geom& g = c->g;
  dim3 blk(128), grd((ix_hi - ix_lo + 1 + 127) / 128, g.ny, g.M);
  k_update_e_bulk<<<grd, blk, 0, c->stream>>>(g, c->f[CYLGPU_EXM], c->f[CYLGPU_ERM], c->f[CYLGPU_ETM],
                                              c->f[CYLGPU_BXM], c->f[CYLGPU_BRM], c->f[CYLGPU_BTM],
                                              c->f[CYLGPU_JXM], c->f[CYLGPU_JRM], c->f[CYLGPU_JTM], field_recips(c),
                                              c->cfg.dy, c->dt, c->cfg.y_grid_min_local, ix_lo, ix_hi);
  k_update_e_axis<<<dim3((g.SX + 127) / 128, g.M, NG), 128, 0, c->stream>>>(g, c->f[CYLGPU_EXM], c->f[CYLGPU_ERM],
                                                             c->f[CYLGPU_ETM], c->f[CYLGPU_BTM],
                                                             c->f[CYLGPU_JXM], c->cfg.dy, c->dt);
  c->stats.kernel_launches += 2;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// save_old: b*_old = b* (fields.f90:326-328) fused into the sweep -- one extra write stream instead of a
// separate read + write pass over the three arrays
int launch_update_b(cylgpu_ctx* c, bool save_old, int ix_lo, int ix_hi) {
  const Geom& g = c->g;
  cplx *bxo = c->f[CYLGPU_BXM_OLD], *bro = c->f[CYLGPU_BRM_OLD], *bto = c->f[CYLGPU_BTM_OLD];
  if (save_old) {
    k_copy_b_old_rim<<<dim3((g.SX + 127) / 128, g.SY, g.M), 128, 0, c->stream>>>(
        g, c->f[CYLGPU_BXM], c->f[CYLGPU_BRM], c->f[CYLGPU_BTM], bxo, bro, bto, ix_lo, ix_hi);
    c->stats.kernel_launches += 1;
  }
  if (g.ny > 1) {
    dim3 blk(128), grd((ix_hi - ix_lo + 1 + 127) / 128, g.ny - 1, g.M);
    if (save_old)
      k_update_b_bulk<true><<<grd, blk, 0, c->stream>>>(g, c->f[CYLGPU_BXM], c->f[CYLGPU_BRM], c->f[CYLGPU_BTM],
                                                        c->f[CYLGPU_EXM], c->f[CYLGPU_ERM], c->f[CYLGPU_ETM], bxo, bro,
                                                        bto, field_recips(c), c->cfg.dy, c->dt, c->cfg.y_grid_min_local, ix_lo,
                                                        ix_hi);
    else
      k_update_b_bulk<false><<<grd, blk, 0, c->stream>>>(g, c->f[CYLGPU_BXM], c->f[CYLGPU_BRM], c->f[CYLGPU_BTM],
                                                         c->f[CYLGPU_EXM], c->f[CYLGPU_ERM], c->f[CYLGPU_ETM], bxo, bro,
                                                         bto, field_recips(c), c->cfg.dy, c->dt, c->cfg.y_grid_min_local, ix_lo,
                                                        ix_hi);
    c->stats.kernel_launches += 1;
  } else if (save_old) {
    const size_t bytes = g.plane * g.M * sizeof(cplx);   // nothing swept: plain copies
    CUDA_TRY(cudaMemcpyAsync(bxo, c->f[CYLGPU_BXM], bytes, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(bro, c->f[CYLGPU_BRM], bytes, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(bto, c->f[CYLGPU_BTM], bytes, cudaMemcpyDeviceToDevice, c->stream));
  }
  k_update_b_axis<<<dim3((g.SX + 127) / 128, g.M, NG), 128, 0, c->stream>>>(g, c->f[CYLGPU_BXM], c->f[CYLGPU_BRM],
                                                             c->f[CYLGPU_BTM], c->f[CYLGPU_EXM],
                                                             c->f[CYLGPU_ETM], c->cfg.dx, c->cfg.dy, c->dt);
  c->stats.kernel_launches += 1;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_update_e(cylgpu_ctx* c) { return launch_update_e(c, 0, c->g.nx); }
int launch_update_b(cylgpu_ctx* c, bool save_old) { return launch_update_b(c, save_old, 0, c->g.nx); }

}  // namespace cylgpu
