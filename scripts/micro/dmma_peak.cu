// Microbenchmark: issue rate of mma.sync.m8n8k4.f64 (DMMA) on sm_100a, register operands only.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/dmma_peak scripts/micro/dmma_peak.cu && build/dmma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int ACC>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double seed) {
  double c[ACC][2];
  double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

template <int ACC>
void run(int warps_per_sm_blocks, int threads, const char* label) {
  double* out;
  cudaMalloc(&out, 8);
  const int iters = 20000;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * warps_per_sm_blocks;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  dmma_kernel<ACC><<<blocks, threads>>>(out, 100, 1.0);
  cudaEventRecord(e0);
  dmma_kernel<ACC><<<blocks, threads>>>(out, iters, 1.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double dmma = (double)blocks * (threads / 32) * iters * ACC;
  printf("%-28s ACC=%2d blocks/SM=%d threads=%d: %.2f TFLOP/s (%.3f ms, %.2f cycles/DMMA/SM at 1.965 GHz)\n", label,
         ACC, warps_per_sm_blocks, threads, dmma * 512 / (ms * 1e-3) / 1e12, ms,
         ms * 1e-3 * 1.965e9 / (dmma / sms));
  cudaFree(out);
}

int main() {
  run<12>(2, 256, "16 warps/SM, 12 chains");
  run<18>(2, 256, "16 warps/SM, 18 chains");
  run<12>(1, 256, "8 warps/SM, 12 chains");
  run<4>(2, 256, "16 warps/SM, 4 chains");
  run<12>(4, 256, "32 warps/SM, 12 chains");
  run<12>(1, 128, "4 warps/SM, 12 chains");
  run<1>(2, 256, "16 warps/SM, 1 chain");
  return 0;
}
