// microbench2.cu -- dependent-chain latency of DFMA / DMMA / LDS.64 and scaling with warps per SM
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 8192
template <int ILP>
__global__ void k_dfma(double* out, double a, double b) {
  double x[ILP];
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  double s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_dmma(double* out, double a, double b) {
  double c[ILP][2];
  for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < ITER; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  double s = 0; for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FP32 for comparison
__global__ void k_ffma(float* out, float a, float b) {
  float x = threadIdx.x;
  for (int it = 0; it < ITER; ++it) x = fmaf(x, a, b);
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
__global__ void k_lds(double* out) {
  __shared__ int sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i + 32) & 1023;
  __syncthreads();
  int idx = threadIdx.x & 1023;
  for (int it = 0; it < ITER; ++it) idx = sm[idx];
  out[blockIdx.x * blockDim.x + threadIdx.x] = idx;
}
template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 2048 * 8);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const double clk = pr.clockRate * 1e3;
  // one CTA per SM, w warps: cycles per dependent op = time * clk / ITER (when not throughput-bound)
  for (int w : {1, 4, 8, 12, 16, 20, 32}) {
    float a = timeit([&] { k_dfma<1><<<148, 32 * w>>>(out, 1.0000001, 1e-9); });
    float a4 = timeit([&] { k_dfma<4><<<148, 32 * w>>>(out, 1.0000001, 1e-9); });
    float d = timeit([&] { k_dmma<1><<<148, 32 * w>>>(out, 1.0000001, 1e-9); });
    float d4 = timeit([&] { k_dmma<4><<<148, 32 * w>>>(out, 1.0000001, 1e-9); });
    float f = timeit([&] { k_ffma<<<148, 32 * w>>>((float*)out, 1.0000001f, 1e-9f); });
    float l = timeit([&] { k_lds<<<148, 32 * w>>>(out); });
    printf("warps/SM %2d: DFMA chain %6.1f clk/op (ILP4: %5.2f inst/clk/SM)  DMMA chain %6.1f clk/op (ILP4 %5.3f inst/clk/SM)  FFMA chain %5.1f  LDS chain %5.1f\n",
           w, a * 1e-3 * clk / ITER, w * 4.0 * ITER / (a4 * 1e-3 * clk), d * 1e-3 * clk / ITER,
           w * 4.0 * ITER / (d4 * 1e-3 * clk), f * 1e-3 * clk / ITER, l * 1e-3 * clk / ITER);
  }
  return 0;
}
