// push_v0.cuh -- push variant 0: one thread per particle, gather through L1/L2, per-particle REDs
// (particles.F90:296-665 with nothing but push.cuh's arithmetic around it).  The simple variant the
// parity tests and the tuned kernels are measured against; barrier-free, so the kernel-emulation tests
// (tests/emul/) also run it on the CPU against the oracle.  Kernel-only header, included inside
// namespace cylgpu after push.cuh.  Product code: no oracle here.
#pragma once

// ------------------------------------------------------------------------------------------
// deposit, particles.F90:584-665, one particle, straight into HBM with FP64 reductions
// (RED.E.ADD.F64 at L2).  Imaginary parts of mode 0 are identically zero and are skipped.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add(double* a, size_t o, cplx v, bool with_imag) {
  atomicAdd(a + 2 * o, v.x);
  if (with_imag) atomicAdd(a + 2 * o + 1, v.y);
}

__device__ __forceinline__ void deposit_global(const PushConst& P, const DepositIn& D) {
  const Geom& g = P.g;
  const double third = 1.0 / 3.0;
  const double* inv_area_rt = P.tab + JNG;              // index by cy directly
  const double* inv_area_xt = P.tab + P.ntab + JNG;
  const double* inv_volume = P.tab + 2 * P.ntab + JNG;
  const double* ratio_area_xt = P.tab + 3 * P.ntab + JNG;
  cplx exp_imtheta0 = C(1.0, 0.0), exp_imdtheta = C(1.0, 0.0);
  for (int im = 0; im < g.M; ++im) {
    ModeFac mf;
    if (im > 0) {
      exp_imtheta0 = exp_imtheta0 * D.exp_itheta_05;
      exp_imdtheta = exp_imdtheta * D.exp_idtheta;
      mf = mode_factors(im, D.dtheta, exp_imtheta0, exp_imdtheta, P.taylor_switch);
    }
    cplx jyh[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) jyh[k] = C(0.0, 0.0);
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
      const int iy = ky - 2;
      if (iy < D.ymin || iy > D.ymax) continue;
      const int cy = D.cell_y2 + iy;
      cplx w_rt, ym_fac_1;
      if (im == 0) {
        w_rt = C(D.gy[ky] + 0.5 * D.hy[ky], 0.0);
        ym_fac_1 = C(0.5 * D.gy[ky] + third * D.hy[ky], 0.0);
      } else {
        w_rt = mf.f2 * D.gy[ky] + mf.f3 * D.hy[ky];
        ym_fac_1 = mf.f3 * D.gy[ky] + mf.f4 * D.hy[ky];
      }
      const double fjx = D.fcx * __ldg(&inv_area_rt[cy]);
      const double fjy = D.fcx * D.hy[ky] * __ldg(&inv_area_xt[cy]);
      const double fjz = D.fcz * __ldg(&inv_volume[cy]);
      const double ratio = __ldg(&ratio_area_xt[cy]);
      cplx jxh = C(0.0, 0.0);
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
        const int ix = kx - 2;
        if (ix < D.xmin || ix > D.xmax) continue;
        const int cx = D.cell_x2 + ix;
        cplx w_xt;
        if (im == 0) w_xt = C(D.gx[kx] + 0.5 * D.hx[kx], 0.0);
        else w_xt = mf.f2 * D.gx[kx] + mf.f3 * D.hx[kx];
        const cplx w_xr = D.gx[kx] * w_rt + D.hx[kx] * ym_fac_1;
        jxh = jxh - (fjx * D.hx[kx]) * w_rt;
        jyh[kx] = jyh[kx] * ratio - fjy * w_xt;
        const cplx jzh = fjz * w_xr;
        const size_t o = g.at(cx, cy, im);
        red_add(P.jx, o + 1, jxh, im > 0);
        red_add(P.jr, o + g.SX, jyh[kx], im > 0);
        red_add(P.jt, o, jzh, im > 0);
      }
    }
  }
}

// variant 0: one thread per particle, everything through L1/L2
template <int M>
__global__ void __launch_bounds__(128) k_push_v0(PushConst P, double* __restrict__ x, double* __restrict__ y,
                                                 double* __restrict__ z, double* __restrict__ px,
                                                 double* __restrict__ py, double* __restrict__ pz,
                                                 const double* __restrict__ w, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double X = x[i], Y = y[i], Z = z[i], PX = px[i], PY = py[i], PZ = pz[i];
  const double W = w[i];
  DepositIn D;
  push_one<M>(P, X, Y, Z, PX, PY, PZ, W, D);
  x[i] = X; y[i] = Y; z[i] = Z;
  px[i] = PX; py[i] = PY; pz[i] = PZ;
  if (P.deposit) deposit_global(P, D);
}
