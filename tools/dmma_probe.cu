// Micro-benchmark: FP64 mma.sync (DMMA) throughput on sm_100a next to the DFMA loop that is the
// roofline denominator of k_bilinear.  Shapes m8n8k4, m16n8k4, m16n8k8, m16n8k16; NT independent
// accumulator tiles per warp.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma_probe
#include <cuda_runtime.h>

#include <cstdio>

__device__ __forceinline__ void mma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
      "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]),
        "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int SHAPE, int NT>
__global__ void __launch_bounds__(256) k_dmma(double *out, int iters, double a0, double b0) {
  double c[NT][4];
#pragma unroll
  for (int t = 0; t < NT; t++)
#pragma unroll
    for (int i = 0; i < 4; i++) c[t][i] = threadIdx.x + t + i;
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = a0 + i * 1e-9 + threadIdx.x * 1e-12;
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = b0 + i * 1e-9;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int t = 0; t < NT; t++) {
        if (SHAPE == 0) {
          double cc[2] = {c[t][0], c[t][1]};
          mma884(cc, a[0], b[0]);
          c[t][0] = cc[0], c[t][1] = cc[1];
        } else if (SHAPE == 1) {
          double aa[2] = {a[0], a[1]};
          mma1684(c[t], aa, b[0]);
        } else if (SHAPE == 2) {
          double aa[4] = {a[0], a[1], a[2], a[3]}, bb[2] = {b[0], b[1]};
          mma1688(c[t], aa, bb);
        } else {
          mma16816(c[t], a, b);
        }
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < NT; t++)
#pragma unroll
    for (int i = 0; i < 4; i++) s += c[t][i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// The k_bilinear shape: 4 M-tiles (32 alpha-side lags) x 3 N-tiles (slots, 8 rows each) per warp,
// A fragments change every k-step (rotated registers stand in for the T stream), B fragments too.
template <int MT>
__global__ void __launch_bounds__(256) k_dmma_tile(double *out, int iters, double a0, double b0) {
  double c[MT][3][2];
#pragma unroll
  for (int m = 0; m < MT; m++)
#pragma unroll
    for (int q = 0; q < 3; q++) c[m][q][0] = threadIdx.x + m, c[m][q][1] = q;
  double a[MT][2], b[3][2];
#pragma unroll
  for (int m = 0; m < MT; m++) a[m][0] = a0 + m * 1e-9 + threadIdx.x * 1e-12, a[m][1] = a0 - m * 1e-9;
#pragma unroll
  for (int q = 0; q < 3; q++) b[q][0] = b0 + q * 1e-9, b[q][1] = b0 - q * 1e-9;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int h = 0; h < 2; h++)
#pragma unroll
        for (int m = 0; m < MT; m++)
#pragma unroll
          for (int q = 0; q < 3; q++) mma884(c[m][q], a[m][h], b[q][h]);
      // rotate so that nothing is loop invariant
      const double t0 = a[0][0];
#pragma unroll
      for (int m = 0; m < MT - 1; m++) a[m][0] = a[m + 1][0];
      a[MT - 1][0] = t0;
      const double t1 = b[0][1];
      b[0][1] = b[1][1], b[1][1] = b[2][1], b[2][1] = t1;
    }
  }
  double s = 0;
#pragma unroll
  for (int m = 0; m < MT; m++)
#pragma unroll
    for (int q = 0; q < 3; q++) s += c[m][q][0] + c[m][q][1];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_peak(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 64; u++) {
      x0 = fma(x0, a, b), x1 = fma(x1, a, b), x2 = fma(x2, a, b), x3 = fma(x3, a, b);
      x4 = fma(x4, a, b), x5 = fma(x5, a, b), x6 = fma(x6, a, b), x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// DMMA stream with KD scalar DFMAs (independent chains) after every 24 DMMAs: what a scalar FP64
// instruction costs the pipe when it is interleaved with DMMAs.
template <int KD>
__global__ void __launch_bounds__(256) k_dmma_mix(double *out, int iters, double a0, double b0) {
  double c[4][3][2];
#pragma unroll
  for (int m = 0; m < 4; m++)
#pragma unroll
    for (int q = 0; q < 3; q++) c[m][q][0] = threadIdx.x + m, c[m][q][1] = q;
  double a[4][2], b[3][2], x[24];
#pragma unroll
  for (int i = 0; i < 24; i++) x[i] = threadIdx.x + i;
#pragma unroll
  for (int m = 0; m < 4; m++) a[m][0] = a0 + m * 1e-9 + threadIdx.x * 1e-12, a[m][1] = a0 - m * 1e-9;
#pragma unroll
  for (int q = 0; q < 3; q++) b[q][0] = b0 + q * 1e-9, b[q][1] = b0 - q * 1e-9;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int h = 0; h < 2; h++)
#pragma unroll
        for (int m = 0; m < 4; m++)
#pragma unroll
          for (int q = 0; q < 3; q++) mma884(c[m][q], a[m][h], b[q][h]);
#pragma unroll
      for (int i = 0; i < KD; i++) x[i] = fma(x[i], a0, b0);
      const double t0 = a[0][0];
#pragma unroll
      for (int m = 0; m < 3; m++) a[m][0] = a[m + 1][0];
      a[3][0] = t0;
      const double t1 = b[0][1];
      b[0][1] = b[1][1], b[1][1] = b[2][1], b[2][1] = t1;
    }
  }
  double s = 0;
#pragma unroll
  for (int m = 0; m < 4; m++)
#pragma unroll
    for (int q = 0; q < 3; q++) s += c[m][q][0] + c[m][q][1];
#pragma unroll
  for (int i = 0; i < 24; i++) s += x[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F>
static double time_ms(F f);
template <int KD>
static void run_mix(double *out, int sms) {
  const int ctas = sms * 2, iters = 2000;
  const double ms = time_ms([&] { k_dmma_mix<KD><<<ctas, 256>>>(out, iters, 0.999999, 1e-9); });
  const double flop = 512.0 * 4 * 24 * iters * (double)ctas * 8;
  std::printf("24 DMMA + %2d DFMA, 16 warps/SM: %7.2f TFLOP/s in DMMAs, %.3f ms\n", KD, flop / ms * 1e-9, ms);
}

template <class F>
static double time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

template <int SHAPE, int NT>
static void run(double *out, int sms, const char *name, double flop_per_mma) {
  for (int cps : {1, 2, 4, 8}) {
    const int ctas = sms * cps, iters = 2000;
    const double ms = time_ms([&] { k_dmma<SHAPE, NT><<<ctas, 256>>>(out, iters, 0.999999, 1e-9); });
    const double flop = flop_per_mma * 8.0 * NT * iters * (double)ctas * 8;
    std::printf("%-10s NT=%2d  %d CTA/SM (%2d warps/SM): %7.2f TFLOP/s\n", name, NT, cps, cps * 8, flop / ms * 1e-9);
  }
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double *out;
  cudaMalloc(&out, (size_t)sms * 8 * 256 * sizeof(double));
  {
    const int ctas = sms * 8, iters = 512;
    const double ms = time_ms([&] { k_peak<<<ctas, 256>>>(out, iters, 0.999999, 1e-9); });
    std::printf("DFMA loop: %7.2f TFLOP/s\n", 2.0 * 512.0 * iters * (double)ctas * 256 / ms * 1e-9);
  }
  run<0, 4>(out, sms, "m8n8k4", 512);
  run<0, 12>(out, sms, "m8n8k4", 512);
  run<1, 6>(out, sms, "m16n8k4", 1024);
  run<2, 6>(out, sms, "m16n8k8", 2048);
  run<3, 6>(out, sms, "m16n8k16", 4096);
  for (int cps : {1, 2, 3, 4}) {
    const int ctas = sms * cps, iters = 2000;
    const double ms = time_ms([&] { k_dmma_tile<4><<<ctas, 256>>>(out, iters, 0.999999, 1e-9); });
    const double flop = 512.0 * 4 * 2 * 12 * iters * (double)ctas * 8;
    std::printf("tile 4x3 m8n8k4, %d CTA/SM: %7.2f TFLOP/s\n", cps, flop / ms * 1e-9);
  }
  for (int tpb : {32, 64, 128, 192}) {  // warps per SM: 1, 2, 4, 6
    const int ctas = sms, iters = 2000;
    const double ms = time_ms([&] { k_dmma_tile<4><<<ctas, tpb>>>(out, iters, 0.999999, 1e-9); });
    const double flop = 512.0 * 4 * 2 * 12 * iters * (double)ctas * (tpb / 32);
    std::printf("tile 4x3 m8n8k4, %d warps/SM: %7.2f TFLOP/s\n", tpb / 32, flop / ms * 1e-9);
  }
  for (int cps : {1, 2, 3, 4}) {
    const int ctas = sms * cps, iters = 2000;
    const double ms = time_ms([&] { k_dmma_tile<2><<<ctas, 256>>>(out, iters, 0.999999, 1e-9); });
    const double flop = 512.0 * 4 * 2 * 6 * iters * (double)ctas * 8;
    std::printf("tile 2x3 m8n8k4, %d CTA/SM: %7.2f TFLOP/s\n", cps, flop / ms * 1e-9);
  }
  run_mix<0>(out, sms);
  run_mix<1>(out, sms);
  run_mix<3>(out, sms);
  run_mix<6>(out, sms);
  run_mix<12>(out, sms);
  run_mix<24>(out, sms);
  std::printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
