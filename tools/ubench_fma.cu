// Micro-benchmark: FP32 FMA issue rates on B200 (FFMA 3-register, with/without same-bank
// operands, and the packed FFMA2).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fma ubench_fma.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
  float2 a[16], x[16];
  const float2 h = make_float2(seed, seed * 0.5f), g = make_float2(seed * 0.25f, seed * 0.125f);
#pragma unroll
  for (int j = 0; j < 16; ++j) { a[j] = make_float2(j, -j); x[j] = make_float2(threadIdx.x + j, 1.0f / (1 + j)); }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (MODE == 0) {          // complex MAC, scalar FFMA (what k_cmac_batch v1 does): 4 FFMA
        a[j].x = fmaf(h.x, x[j].x, a[j].x);
        a[j].x = fmaf(-h.y, x[j].y, a[j].x);
        a[j].y = fmaf(h.x, x[j].y, a[j].y);
        a[j].y = fmaf(h.y, x[j].x, a[j].y);
      } else if (MODE == 1) {   // 2 x FFMA2 (4 lane-FMAs)
        a[j] = __ffma2_rn(make_float2(h.x, h.x), x[j], a[j]);
        a[j] = __ffma2_rn(make_float2(h.y, h.y), x[j], a[j]);
      } else {                  // 4 FFMA, operands chosen so that acc and x have opposite parity when allocated as pairs
        a[j].x = fmaf(h.x, x[j].y, a[j].x);
        a[j].y = fmaf(h.y, x[j].x, a[j].y);
        a[j].x = fmaf(g.x, x[j].y, a[j].x);
        a[j].y = fmaf(g.y, x[j].x, a[j].y);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += a[j].x + a[j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int ctas_per_sm) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  const int iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * ctas_per_sm, 256>>>(out, 10, 1.0f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<148 * ctas_per_sm, 256>>>(out, iters, 1.0f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = (double)148 * ctas_per_sm * 256 * iters * 16 * 4;
  printf("%-28s ctas/sm=%d  %.3f ms  %.2f TFLOP/s  (%.1f lane-FMA/clk/SM @1.965GHz)\n", name, ctas_per_sm, ms,
         2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}

int main() {
  for (int c = 1; c <= 2; ++c) {
    run<0>("FFMA complex-MAC", c);
    run<1>("FFMA2 (packed f32x2)", c);
    run<2>("FFMA opposite-parity", c);
  }
  return 0;
}
