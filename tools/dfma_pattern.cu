// Micro-benchmark: DFMA throughput of the k_bilinear inner pattern with every operand already in
// registers (no loads in the loop): acc[r][q] += t[s] * w[q][s - r + 7], 8 rows x 3 slots x 8
// lags, in different instruction orders.  Tells how much of the FP64 pipe peak the register
// file can feed for this operand pattern.   build: nvcc -arch=sm_100a -O3 -o build/dfma_pattern
#include <cuda_runtime.h>

#include <cstdio>

// WMODE 0: windows loop-invariant and warp-uniform (ptxas keeps them in uniform registers)
//       1: windows loop-invariant in vector registers (3 vector operands per DFMA)
//       2: windows re-loaded from shared memory every iteration (LDS.128, uniform address)
template <int ORDER, int WMODE>
__global__ void __launch_bounds__(352, WMODE == 1 ? 1 : 2) k_pat(const double *__restrict__ in, double *out, int iters, int zero) {
  __shared__ __align__(16) double s_w[3][352];
  double t[8], w[3][16], acc[8][3];
  for (int i = threadIdx.x; i < 3 * 352; i += 352) s_w[i / 352][i % 352] = in[4096 + i];
  __syncthreads();
#pragma unroll
  for (int s = 0; s < 8; s++) t[s] = in[threadIdx.x + 352 * s];
#pragma unroll
  for (int q = 0; q < 3; q++)
#pragma unroll
    for (int j = 0; j < 16; j++) w[q][j] = in[4096 + q * 16 + j + (WMODE == 1 ? threadIdx.x * zero : 0)];
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int q = 0; q < 3; q++) acc[r][q] = 0.0;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    double wn[3][16];
    if (WMODE == 4) {  // loop-carried: the windows used now were loaded during the previous chunk
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const double2 *wp = reinterpret_cast<const double2 *>(&s_w[q][((it + 1) & 31) * 8]);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const double2 v = wp[i];
          wn[q][2 * i] = v.x, wn[q][2 * i + 1] = v.y;
        }
      }
    }
    if (WMODE == 3) {  // all three slots' windows at the top of the chunk: 96 registers' worth
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const double2 *wp = reinterpret_cast<const double2 *>(&s_w[q][(it & 31) * 8]);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const double2 v = wp[i];
          w[q][2 * i] = v.x, w[q][2 * i + 1] = v.y;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 3; q++) {
      if (WMODE == 2) {
        const double2 *wp = reinterpret_cast<const double2 *>(&s_w[q][(it & 31) * 8]);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const double2 v = wp[i];
          w[q][2 * i] = v.x, w[q][2 * i + 1] = v.y;
        }
      }
      if (ORDER == 0) {  // lag-major: consecutive DFMAs share t[s]
#pragma unroll
        for (int s = 0; s < 8; s++)
#pragma unroll
          for (int r = 0; r < 8; r++) acc[r][q] = fma(t[s], w[q][s - r + 7], acc[r][q]);
      } else if (ORDER == 1) {  // window-major: consecutive DFMAs share w[j]
#pragma unroll
        for (int j = 0; j < 15; j++)
#pragma unroll
          for (int s = 0; s < 8; s++) {
            const int r = s + 7 - j;
            if (r >= 0 && r < 8) acc[r][q] = fma(t[s], w[q][j], acc[r][q]);
          }
      } else if (ORDER == 2) {  // row-major: consecutive DFMAs share the accumulator (dependent chain!)
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
          for (int s = 0; s < 8; s++) acc[r][q] = fma(t[s], w[q][s - r + 7], acc[r][q]);
      } else {  // snake over (s, j): every DFMA shares t or w with its predecessor
#pragma unroll
        for (int s = 0; s < 8; s++)
#pragma unroll
          for (int rr = 0; rr < 8; rr++) {
            const int r = (s & 1) ? 7 - rr : rr;
            acc[r][q] = fma(t[s], w[q][s - r + 7], acc[r][q]);
          }
      }
    }
    if (WMODE == 4) {
#pragma unroll
      for (int q = 0; q < 3; q++)
#pragma unroll
        for (int j = 0; j < 16; j++) w[q][j] = wn[q][j];
    }
    // rotate t so that the compiler cannot hoist products out of the loop
    const double t0 = t[0];
#pragma unroll
    for (int s = 0; s < 7; s++) t[s] = t[s + 1];
    t[7] = t0;
  }
  double s = 0;
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int q = 0; q < 3; q++) s += acc[r][q];
  out[blockIdx.x * 352 + threadIdx.x] = s;
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

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int ctas = prop.multiProcessorCount * 2 * 8, iters = 2000;
  double *in, *out;
  cudaMalloc(&in, 8192 * sizeof(double));
  cudaMemset(in, 0, 8192 * sizeof(double));
  cudaMalloc(&out, (size_t)ctas * 352 * sizeof(double));
  const double ms_peak = time_ms([&] { k_peak<<<prop.multiProcessorCount * 8, 256>>>(out, 2048, 0.999999, 1e-9); });
  const double peak = 2.0 * 512 * 2048 * prop.multiProcessorCount * 8.0 * 256 / ms_peak * 1e-9;
  printf("peak (2 reused operands)      %7.2f TFLOP/s\n", peak);
  const double flop = 2.0 * 192 * iters * (double)ctas * 352;
  double ms;
#define RUN(O, W, name)                                                                   \
  ms = time_ms([&] { k_pat<O, W><<<ctas, 352>>>(in, out, iters, 0); });                     \
  printf("%-44s %7.2f TFLOP/s = %5.1f %%\n", name, flop / ms * 1e-9, 100 * flop / ms * 1e-9 / peak);
  RUN(0, 0, "lag-major, windows in uniform registers");
  RUN(1, 0, "window-major, windows in uniform registers");
  RUN(0, 1, "lag-major, windows in vector registers");
  RUN(1, 1, "window-major, windows in vector registers");
  RUN(2, 1, "row-major, windows in vector registers");
  RUN(0, 2, "lag-major, windows from LDS.128 per chunk");
  RUN(1, 2, "window-major, windows from LDS.128 per chunk");
  RUN(0, 4, "lag-major, windows prefetched one chunk ahead (loop-carried)");
  RUN(1, 4, "window-major, windows prefetched one chunk ahead");
  RUN(0, 3, "lag-major, all windows loaded at chunk top");
  RUN(1, 3, "window-major, all windows loaded at chunk top");
  return 0;
}
