// Effective mass of a CV from its gradient (cvpack/utils.py:172-184) and the FP32 FMA
// microbenchmark that gives the pair kernels' roofline denominator.
#include "common.cuh"

namespace cvp {
namespace {

// One block per frame: sum_i |g_i|^2/m_i over atoms with non-zero |g_i|^2 (double, fixed order).
template <typename T>
__global__ void __launch_bounds__(256)
    effective_mass_kernel(const T* __restrict__ grad, const double* __restrict__ masses,
                          int64_t num_atoms, T* __restrict__ emass) {
  __shared__ double scratch[2 * 8];
  const int64_t b = blockIdx.x;
  const T* g = grad + b * num_atoms * 3;
  double v[2] = {0.0, 0.0};
  for (int64_t i = threadIdx.x; i < num_atoms; i += blockDim.x) {
    const double gx = g[3 * i], gy = g[3 * i + 1], gz = g[3 * i + 2];
    const double sq = gx * gx + gy * gy + gz * gz;
    if (sq != 0.0) {
      v[0] += sq / masses[i];
      v[1] += 1.0;
    }
  }
  block_sum<2>(v, scratch);
  if (threadIdx.x == 0) emass[b] = v[1] == 0.0 ? T(INFINITY) : T(1.0 / v[0]);
}

__global__ void __launch_bounds__(256) fp32_fma_kernel(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f,
        a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float m = 0.999f, c = 1e-3f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      a0 = fmaf(a0, m, c);
      a1 = fmaf(a1, m, c);
      a2 = fmaf(a2, m, c);
      a3 = fmaf(a3, m, c);
      a4 = fmaf(a4, m, c);
      a5 = fmaf(a5, m, c);
      a6 = fmaf(a6, m, c);
      a7 = fmaf(a7, m, c);
    }
  }
  float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456f) out[0] = s;  // keep the loop alive
}

__global__ void fp64_fma_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3., a4 = a0 + 4., a5 = a0 + 5.,
         a6 = a0 + 6., a7 = a0 + 7.;
  const double m = 0.999, c = 1e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      a0 = fma(a0, m, c);
      a1 = fma(a1, m, c);
      a2 = fma(a2, m, c);
      a3 = fma(a3, m, c);
      a4 = fma(a4, m, c);
      a5 = fma(a5, m, c);
      a6 = fma(a6, m, c);
      a7 = fma(a7, m, c);
    }
  }
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456) out[0] = s;
}

}  // namespace

int effective_mass(int dtype, const void* grad, const double* masses, int64_t B, int64_t N,
                   void* emass, cudaStream_t stream) {
  if (dtype == CVP_F32)
    effective_mass_kernel<float><<<(unsigned)B, 256, 0, stream>>>(
        static_cast<const float*>(grad), masses, N, static_cast<float*>(emass));
  else
    effective_mass_kernel<double><<<(unsigned)B, 256, 0, stream>>>(
        static_cast<const double*>(grad), masses, N, static_cast<double*>(emass));
  CVP_LAUNCH_CHECK();
  return CVP_OK;
}

int measure_fp32_peak(double* tflops) {
  int device = 0;
  CVP_CUDA_CHECK(cudaGetDevice(&device));
  cudaDeviceProp prop;
  CVP_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  float* out = nullptr;
  CVP_CUDA_CHECK(cudaMalloc(&out, sizeof(float)));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  cudaEvent_t start, stop;
  CVP_CUDA_CHECK(cudaEventCreate(&start));
  CVP_CUDA_CHECK(cudaEventCreate(&stop));
  fp32_fma_kernel<<<blocks, threads>>>(out, iters);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CVP_CUDA_CHECK(cudaEventRecord(start));
    fp32_fma_kernel<<<blocks, threads>>>(out, iters);
    CVP_CUDA_CHECK(cudaEventRecord(stop));
    CVP_CUDA_CHECK(cudaEventSynchronize(stop));
    float ms = 0.f;
    CVP_CUDA_CHECK(cudaEventElapsedTime(&ms, start, stop));
    const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
    best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  g_launch_count.fetch_add(6);
  cudaEventDestroy(start);
  cudaEventDestroy(stop);
  cudaFree(out);
  *tflops = best;
  return CVP_OK;
}

int measure_fp64_peak(double* tflops) {
  int device = 0;
  CVP_CUDA_CHECK(cudaGetDevice(&device));
  cudaDeviceProp prop;
  CVP_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  double* out = nullptr;
  CVP_CUDA_CHECK(cudaMalloc(&out, sizeof(double)));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1024;
  cudaEvent_t start, stop;
  CVP_CUDA_CHECK(cudaEventCreate(&start));
  CVP_CUDA_CHECK(cudaEventCreate(&stop));
  fp64_fma_kernel<<<blocks, threads>>>(out, iters);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CVP_CUDA_CHECK(cudaEventRecord(start));
    fp64_fma_kernel<<<blocks, threads>>>(out, iters);
    CVP_CUDA_CHECK(cudaEventRecord(stop));
    CVP_CUDA_CHECK(cudaEventSynchronize(stop));
    float ms = 0.f;
    CVP_CUDA_CHECK(cudaEventElapsedTime(&ms, start, stop));
    const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
    best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  g_launch_count.fetch_add(6);
  cudaEventDestroy(start);
  cudaEventDestroy(stop);
  cudaFree(out);
  *tflops = best;
  return CVP_OK;
}

}  // namespace cvp
