// microbench.cu -- instruction-throughput probes on B200 that decide the deposit design:
// DFMA, DMMA m8n8k4 (FP64 tensor), SHFL.BFLY, FSEL, LDS.128 broadcast, and DFMA+DMMA together.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
__global__ void k_dfma(double* out, double a, double b) {
  double x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  double s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k_dmma(double* out, double a, double b) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
  double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_both(double* out, double a, double b) {
  double c[4][2], x[8];
  for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0; for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_shfl(double* out) {
  int x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1 + (i & 3)) + 1;
  int s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fsel(double* out, int p) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  const bool h = (threadIdx.x & p) != 0;
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = h ? x[(i + 1) & 7] : x[(i + 3) & 7];
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// LDS.128: each quarter-warp group reads one 16-byte word (4 distinct addresses = broadcast)
__global__ void k_lds128(double* out, int stride) {
  __shared__ double2 sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_double2(i, 1);
  __syncthreads();
  int idx = (threadIdx.x >> 3) & 3;
  double s = 0;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double2 v = sm[(idx + i * stride) & 2047];
      s += v.x; idx += (int)v.y;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// LDS.64 with 32 distinct addresses (fragment-style load: 2 wavefronts)
__global__ void k_lds64(double* out, int stride) {
  __shared__ double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1;
  __syncthreads();
  int idx = (threadIdx.x & 31);
  double s = 0;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double v = sm[(idx + i * stride) & 4095];
      s += v; idx += (int)v;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 8 * 1024 * 8);
  const int nb = 148 * 4, nt = 512;   // 64 warps per SM
  const double warps = (double)nb * nt / 32;
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const double clk = pr.clockRate * 1e3;
  auto rep = [&](const char* nm, float ms, double inst_per_warp) {
    const double per_sm_clk = warps * inst_per_warp / 148.0 / (ms * 1e-3 * clk);
    printf("%-28s %8.3f ms  %6.2f warp-inst/clk/SM (at %.0f MHz nominal)\n", nm, ms, per_sm_clk, clk / 1e6);
  };
  rep("DFMA", timeit([&] { k_dfma<<<nb, nt>>>(out, 1.0000001, 1e-9); }), 8.0 * ITER);
  rep("DMMA m8n8k4", timeit([&] { k_dmma<<<nb, nt>>>(out, 1.0000001, 1e-9); }), 8.0 * ITER);
  rep("DMMA(4)+DFMA(8) [12/iter]", timeit([&] { k_both<<<nb, nt>>>(out, 1.0000001, 1e-9); }), 12.0 * ITER);
  rep("SHFL.BFLY 32b", timeit([&] { k_shfl<<<nb, nt>>>(out); }), 8.0 * ITER);
  rep("FSEL", timeit([&] { k_fsel<<<nb, nt>>>(out, 16); }), 8.0 * ITER);
  rep("LDS.128 4-addr bcast", timeit([&] { k_lds128<<<nb, nt>>>(out, 5); }), 8.0 * ITER);
  rep("LDS.64 32-addr", timeit([&] { k_lds64<<<nb, nt>>>(out, 33); }), 8.0 * ITER);
  // 16 warps per SM, as the push kernel runs
  const int nb2 = 148 * 4, nt2 = 128;
  const double w2 = (double)nb2 * nt2 / 32;
  float ms = timeit([&] { k_dmma<<<nb2, nt2>>>(out, 1.0000001, 1e-9); });
  printf("DMMA 16 warps/SM             %8.3f ms  %6.2f warp-inst/clk/SM\n", ms, w2 * 8.0 * ITER / 148.0 / (ms * 1e-3 * clk));
  ms = timeit([&] { k_dfma<<<nb2, nt2>>>(out, 1.0000001, 1e-9); });
  printf("DFMA 16 warps/SM             %8.3f ms  %6.2f warp-inst/clk/SM\n", ms, w2 * 8.0 * ITER / 148.0 / (ms * 1e-3 * clk));
  return 0;
}
