// Microbenchmark: ex2.approx.f16x2 (two half-precision exponentials per instruction) vs ex2.approx.ftz.f32, elements per
// clock and SM; mode 1 adds the FFMA2-equivalent scaling and the f32 -> f16x2 pack the softmax would need in front of it.
// Build + run on the GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_h2 tools/ubench/mufu_h2.cu && /tmp/mufu_h2
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  unsigned x[8];
  float f[16];
  for (int i = 0; i < 8; ++i) x[i] = 0x38003800u + threadIdx.x + i;
  for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-3f + i * 0.01f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      unsigned y = x[i];
      if (MODE >= 1) {
        float a = fmaf(f[2 * i], 0.999f, -1.0f), b = fmaf(f[2 * i + 1], 0.999f, -1.0f);
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(b), "f"(a));
        f[2 * i] = a; f[2 * i + 1] = b;
        y ^= x[i] & 0x00010001u;
      }
      asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(y));
      x[i] = y;
    }
  }
  long long t1 = clock64();
  unsigned s = 0; for (int i = 0; i < 8; ++i) s ^= x[i];
  float fs = 0; for (int i = 0; i < 16; ++i) fs += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s + fs;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(int warps) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  int iters = 2000;
  k<MODE><<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
  k<MODE><<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("f16x2 mode %d warps/SM %2d: %.1f exponentials / clk / SM\n", MODE, warps, (double)warps * 32 * 16 * iters / avg);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8, 16, 32}) run<0>(w);
  for (int w : {4, 8, 16, 32}) run<1>(w);
  return 0;
}
