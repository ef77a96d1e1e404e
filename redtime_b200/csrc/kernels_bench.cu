// Roofline denominators for the FP64 pipe, timed with CUDA events (MEASURED_PEAKS.json holds no FP64
// figure, so bench.py measures them live and quotes the nominal 148 SM x 64 FMA/clk x 2 x f_SM
// beside them):
//   rtrg_bench_dmma  register-resident DMMA.8x8x4 loop (mma.sync.m8n8k4.f64), 12 independent
//                    accumulator tiles per warp -- the instruction k_bilinear runs on; reaches the
//                    nominal rate (37.2 TFLOP/s at 1.965 GHz)
//   rtrg_bench_dfma  register-resident scalar DFMA loop, 8 independent chains per thread (34.3)
#include <cuda_runtime.h>

#include "../../include/redtime_b200.h"

namespace {
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 64; u++) {
      x0 = fma(x0, a, b), x1 = fma(x1, a, b), x2 = fma(x2, a, b), x3 = fma(x3, a, b);
      x4 = fma(x4, a, b), x5 = fma(x5, a, b), x6 = fma(x6, a, b), x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true: keeps the loop
}
__global__ void __launch_bounds__(256) k_dmma_peak(double *out, int iters, double a0, double b0) {
  double c[12][2];
#pragma unroll
  for (int i = 0; i < 12; i++) c[i][0] = threadIdx.x + i, c[i][1] = i;
  const double a = a0 + threadIdx.x * 1e-12, b = b0;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int i = 0; i < 12; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1])
                     : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true: keeps the loop
}
}  // namespace

extern "C" int rtrg_bench_dfma(int device, double seconds, double *tflops) {
  if (!tflops) return RTRG_EINVAL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return RTRG_ENOGPU;
  cudaSetDevice(device);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  double *out = nullptr;
  const int ctas = prop.multiProcessorCount * 8, tpb = 256, iters = 512;
  if (cudaMalloc(&out, (size_t)ctas * tpb * sizeof(double)) != cudaSuccess) return RTRG_ENOMEM;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const double flop = 2.0 * 512.0 * iters * (double)ctas * tpb;
  k_dfma_peak<<<ctas, tpb>>>(out, iters, 0.999999, 1e-9);  // warm-up
  cudaDeviceSynchronize();
  double best = 0, spent = 0;
  for (int rep = 0; rep < 200 && spent < seconds; rep++) {
    cudaEventRecord(e0);
    k_dfma_peak<<<ctas, tpb>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    spent += ms * 1e-3;
    const double tf = flop / (ms * 1e-3) * 1e-12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return cudaGetLastError() == cudaSuccess ? RTRG_OK : RTRG_ECUDA;
}

extern "C" int rtrg_bench_dmma(int device, double seconds, double *tflops) {
  if (!tflops) return RTRG_EINVAL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return RTRG_ENOGPU;
  cudaSetDevice(device);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  double *out = nullptr;
  const int ctas = prop.multiProcessorCount * 4, tpb = 256, iters = 1024;
  if (cudaMalloc(&out, (size_t)ctas * tpb * sizeof(double)) != cudaSuccess) return RTRG_ENOMEM;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  // one m8n8k4 DMMA = 8 x 8 x 4 FMAs per warp
  const double flop = 2.0 * 256.0 * 8 * 12 * iters * (double)ctas * (tpb / 32);
  k_dmma_peak<<<ctas, tpb>>>(out, iters, 0.999999, 1e-9);  // warm-up
  cudaDeviceSynchronize();
  double best = 0, spent = 0;
  for (int rep = 0; rep < 200 && spent < seconds; rep++) {
    cudaEventRecord(e0);
    k_dmma_peak<<<ctas, tpb>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    spent += ms * 1e-3;
    const double tf = flop / (ms * 1e-3) * 1e-12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return cudaGetLastError() == cudaSuccess ? RTRG_OK : RTRG_ECUDA;
}
