// Measurement helpers behind the C ABI: kernel-launch count, device timing of the phases of apb_run_steps and a live
// FP64 peak measurement (the roofline denominator for the force kernels; MEASURED_PEAKS.json holds no FP64 figure).
#include "internal.cuh"

extern "C" int apb_get_launch_count(apb_handle h, int64_t *out) {
  if (!h || !out) return APB_ERR_INVALID_ARGUMENT;
  *out = h->launchCount;
  return APB_OK;
}

// phase timing: events are recorded by apb_run_steps when enabled
extern "C" int apb_enable_loop_timing(apb_handle h, int32_t enable) {
  APB_ENTRY(h);
  h->loopTiming = enable != 0;
  return APB_OK;
}

int apbLoopTimingRecord(apb_handle h, int phase, bool begin) {
  if (!h->loopTiming) return APB_OK;
  cudaEvent_t ev;
  if (h->eventPool.empty()) {
    APB_CUDA(cudaEventCreate(&ev));
  } else {
    ev = static_cast<cudaEvent_t>(h->eventPool.back());
    h->eventPool.pop_back();
  }
  APB_CUDA(cudaEventRecord(ev, h->stream));
  h->loopEvents.push_back({ev, phase, begin});
  return APB_OK;
}

// out_ms[0] force kernels (+ statistics reduction), out_ms[1] rebuild (migrate + halo generation + structure build),
// out_ms[2] halo refresh, out_ms[3] integration; out_counts: number of intervals per phase. Resets the record.
extern "C" int apb_get_loop_timing(apb_handle h, double *out_ms, int64_t *out_counts) {
  APB_ENTRY(h);
  if (!out_ms || !out_counts) return h->fail(APB_ERR_INVALID_ARGUMENT, "apb_get_loop_timing: null argument");
  for (int k = 0; k < 4; ++k) {
    out_ms[k] = 0.;
    out_counts[k] = 0;
  }
  APB_CUDA(cudaStreamSynchronize(h->stream));
  cudaEvent_t open[4] = {nullptr, nullptr, nullptr, nullptr};
  for (auto &e : h->loopEvents) {
    cudaEvent_t ev = static_cast<cudaEvent_t>(e.ev);
    if (e.begin) {
      open[e.phase] = ev;
    } else if (open[e.phase]) {
      float ms = 0.f;
      APB_CUDA(cudaEventElapsedTime(&ms, open[e.phase], ev));
      out_ms[e.phase] += ms;
      out_counts[e.phase] += 1;
      open[e.phase] = nullptr;
    }
  }
  for (auto &e : h->loopEvents) h->eventPool.push_back(e.ev);
  h->loopEvents.clear();
  return APB_OK;
}

// ---- FP64 peak: independent DFMA chains, 2 flops per FMA -----------------------------------------------------------
__global__ void __launch_bounds__(256) kDfmaPeak(double *out, int iters, double a, double b) {
  double r0 = threadIdx.x * 1e-3, r1 = r0 + 1., r2 = r0 + 2., r3 = r0 + 3., r4 = r0 + 4., r5 = r0 + 5., r6 = r0 + 6.,
         r7 = r0 + 7.;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      r0 = fma(r0, a, b);
      r1 = fma(r1, a, b);
      r2 = fma(r2, a, b);
      r3 = fma(r3, a, b);
      r4 = fma(r4, a, b);
      r5 = fma(r5, a, b);
      r6 = fma(r6, a, b);
      r7 = fma(r7, a, b);
    }
  }
  out[static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x] = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7;
}

// Sustained DFMA rate of the device in TFLOP/s (best of `repeats` launches of >= ~20 ms each).
extern "C" int apb_measure_fp64_peak(int32_t device, int32_t repeats, double *out_tflops, double *out_ms) {
  if (!out_tflops) return APB_ERR_INVALID_ARGUMENT;
  if (cudaSetDevice(device) != cudaSuccess) return APB_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return APB_ERR_CUDA;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 40000;
  double *buf = nullptr;
  if (cudaMalloc(&buf, sizeof(double) * blocks * threads) != cudaSuccess) return APB_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0., bestMs = 0.;
  for (int r = 0; r < (repeats > 0 ? repeats : 3) + 1; ++r) {
    cudaEventRecord(e0);
    kDfmaPeak<<<blocks, threads>>>(buf, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) {
      cudaFree(buf);
      return APB_ERR_CUDA;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 64.0 * iters * static_cast<double>(blocks) * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) {  // first launch is warm-up
      best = tf;
      bestMs = ms;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  *out_tflops = best;
  if (out_ms) *out_ms = bestMs;
  return APB_OK;
}
