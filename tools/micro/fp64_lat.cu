// FP64 pipe microbenchmark: dependent-chain latency and throughput as a function of resident warps and ILP.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, double a, double b, long long *cyc) {
  double x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) x[j] = a + threadIdx.x + j;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int warpsPerSM) {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 148 * 2048 * 8); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  int threads = warpsPerSM * 32;  // one block per SM
  int blocks = 148;
  if (threads > 1024) { blocks = 148 * (threads / 1024); threads = 1024; }
  k<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, cyc);
  cudaDeviceSynchronize();
  k<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double perInstrPerWarp = double(h) / (double(iters) * ILP);
  double warpsPerSmsp = warpsPerSM / 4.0;
  printf("ILP %d warps/SM %3d: %.2f cycles per DFMA per warp -> SMSP pipe busy %.0f%% (2 cyc/instr)\n", ILP, warpsPerSM,
         perInstrPerWarp, 100.0 * 2.0 * warpsPerSmsp / perInstrPerWarp);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8, 16, 32, 64}) run<1>(w);
  for (int w : {4, 8, 16, 32, 64}) run<2>(w);
  for (int w : {4, 8, 16, 32}) run<4>(w);
  for (int w : {4, 8, 16}) run<8>(w);
  return 0;
}
