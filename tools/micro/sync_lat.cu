// host round-trip costs that the rebuild path pays: tiny kernel + 8-byte D2H + stream sync, and back-to-back launches
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long *p) { if (threadIdx.x == 0) *p += 1; }
int main() {
  long long *d, h = 0, *pin;
  cudaMalloc(&d, 8); cudaMemset(d, 0, 8); cudaMallocHost(&pin, 8);
  cudaStream_t s; cudaStreamCreate(&s);
  auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  for (int rep = 0; rep < 3; ++rep) {
    double t0 = now();
    for (int i = 0; i < 1000; ++i) { k<<<1, 32, 0, s>>>(d); cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); }
    double t1 = now();
    for (int i = 0; i < 1000; ++i) { k<<<1, 32, 0, s>>>(d); cudaMemcpyAsync(pin, d, 8, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); }
    double t2 = now();
    for (int i = 0; i < 1000; ++i) k<<<1, 32, 0, s>>>(d);
    cudaStreamSynchronize(s);
    double t3 = now();
    printf("kernel + 8 B D2H (pageable) + sync: %.1f us; (pinned): %.1f us; back-to-back launch: %.2f us per kernel\n", (t1 - t0) / 1000, (t2 - t1) / 1000, (t3 - t2) / 1000);
  }
  return 0;
}
