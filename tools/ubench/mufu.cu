// Microbenchmark: MUFU.EX2 throughput per SM (ops / clk), alone and mixed with FFMA2 / F2FP, vs warps per SM.
// Build + run on the GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/mufu tools/ubench/mufu.cu && tools/ubench/mufu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[i]));
      if (MODE >= 1) y = fmaf(y, 0.999f, -1.0f);
      x[i] = y;
    }
    if (MODE >= 2) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) { unsigned u; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(x[i]), "f"(x[i + 1])); acc ^= u; }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(int warps) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  int iters = 2000;
  k<MODE><<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
  k<MODE><<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("mode %d warps/SM %2d: %.1f ex2 / clk / SM\n", MODE, warps, (double)warps * 32 * 8 * iters / avg);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8, 16, 32}) run<0>(w);
  for (int w : {4, 8, 16, 32}) run<1>(w);
  for (int w : {4, 8, 16, 32}) run<2>(w);
  return 0;
}
