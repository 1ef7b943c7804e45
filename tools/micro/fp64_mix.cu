// Does the FP64 pipe keep its 2 cycles / warp instruction when integer / LDS instructions are interleaved?
// NI integer ops (LOP3/IADD chain) and NL shared loads per 4 DFMAs, 8 warps per SMSP resident.
#include <cstdio>
#include <cuda_runtime.h>
template <int NI, int NL>
__global__ void k(double *out, int iters, double a, double b, long long *cyc, int seed) {
  __shared__ double sm[1024];
  sm[threadIdx.x & 1023] = threadIdx.x;
  __syncthreads();
  double x0 = a + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  unsigned u[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) u[j] = seed + threadIdx.x * (j + 1);
  double acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
#pragma unroll
    for (int j = 0; j < NI; ++j) u[j & 7] = (u[j & 7] ^ (u[(j + 1) & 7] >> 3)) + seed;
#pragma unroll
    for (int j = 0; j < NL; ++j) acc += sm[(u[j & 7] + i) & 1023];
  }
  long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += u[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + s + acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NI, int NL>
void run() {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  k<NI, NL><<<148, 1024>>>(out, iters, 1.0000001, 1e-9, cyc, 3);
  cudaDeviceSynchronize();
  k<NI, NL><<<148, 1024>>>(out, iters, 1.0000001, 1e-9, cyc, 3);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  // 8 warps per SMSP, 4 DFMA per iteration each -> ideal 64 cycles per iteration
  printf("int ops %2d (x2 instr) lds %d per 4 DFMA: %.1f cycles / iteration (FP64-bound ideal 64) -> FP64 pipe %.0f%%\n", NI, NL,
         double(h) / iters, 100.0 * 64.0 * iters / double(h));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0, 0>(); run<1, 0>(); run<2, 0>(); run<4, 0>(); run<6, 0>(); run<0, 1>(); run<0, 2>(); run<2, 1>(); run<4, 2>();
  return 0;
}
