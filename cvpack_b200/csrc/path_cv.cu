// PathInCVSpace: progress / deviation along a path of milestones in the space of other collective
// variables (cvpack/path_in_cv_space.py:124-169 builds, per milestone, the expression
//   x_i = sum_j (delta_ij / scale_j)^2,  delta_ij = m_ij - cv_j,  periodic cv: min(|d|, P - |d|)
// and cvpack/base_path_cv.py:36-62 combines them: w_i = exp(lambda (xmin - x_i)), lambda =
// 1/(2 sigma^2), s = sum_i i w_i / ((n-1) sum w), z = xmin - ln(sum w)/lambda).
//
// The inner CVs are evaluated by their own kernels; here
//   path_cv_combine_kernel   one thread per frame: value and d value / d cv_j  (chain rule exactly
//                            as Lepton applies it: min() by the branch taken, second argument on
//                            ties; abs'(0) = +1)
//   axpy_frames_kernel       gradient [B][N][3] (+)= coef[b] * inner gradient [B][N][3]
#include "common.cuh"

namespace cvp {
namespace {

// m(delta) and dm/d cv (delta = milestone - cv)
__device__ __forceinline__ void path_delta(double delta, double period, double& m, double& dm) {
  if (period > 0.0) {
    const double a = fabs(delta);
    const double sign = delta >= 0.0 ? 1.0 : -1.0;  // d|d|/dd, +1 at 0
    if (a < period - a) {
      m = a;
      dm = -sign;  // d delta / d cv = -1
    } else {
      m = period - a;
      dm = sign;
    }
  } else {
    m = delta;
    dm = -1.0;
  }
}

template <typename T>
__global__ void path_cv_combine_kernel(const T* __restrict__ values /* [nv][B] */, int nv, int n,
                                       const double* __restrict__ milestones /* [n][nv] */,
                                       const double* __restrict__ periods,
                                       const double* __restrict__ inv_scale2, double lambda,
                                       int progress, int64_t num_frames, T* __restrict__ out_value,
                                       T* __restrict__ out_coef /* [nv][B] or null */) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= num_frames) return;
  auto sqdist = [&](int i) {
    double x = 0.0;
    for (int j = 0; j < nv; ++j) {
      double m, dm;
      path_delta(milestones[i * nv + j] - (double)values[(int64_t)j * num_frames + b], periods[j], m,
                 dm);
      x += m * m * inv_scale2[j];
    }
    return x;
  };
  // xmin0 = min(x0, x1), xmin_k = min(xmin_{k-1}, x_{k+1})
  double xmin = sqdist(0);
  for (int i = 1; i < n; ++i) xmin = fmin(xmin, sqdist(i));
  double wsum = 0.0, iwsum = 0.0;
  for (int i = 0; i < n; ++i) {
    const double w = exp(lambda * (xmin - sqdist(i)));
    wsum += w;
    iwsum += i * w;
  }
  const double s = iwsum / ((n - 1) * wsum);
  out_value[b] = T(progress ? s : xmin - log(wsum) / lambda);
  if (out_coef == nullptr) return;
  for (int j = 0; j < nv; ++j) {
    double c = 0.0;
    for (int i = 0; i < n; ++i) {
      const double w = exp(lambda * (xmin - sqdist(i)));
      // d value / d x_i (the derivative through xmin cancels: base_path_cv.py's expression is
      // invariant to the shift)
      const double dx = progress ? -lambda * w * ((double)i / (n - 1) - s) / wsum : w / wsum;
      double m, dm;
      path_delta(milestones[i * nv + j] - (double)values[(int64_t)j * num_frames + b], periods[j], m,
                 dm);
      c += dx * 2.0 * m * dm * inv_scale2[j];
    }
    out_coef[(int64_t)j * num_frames + b] = T(c);
  }
}

template <typename T>
__global__ void axpy_frames_kernel(const T* __restrict__ coef, const T* __restrict__ x,
                                   T* __restrict__ y, int64_t per_frame, int accumulate,
                                   int64_t total) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const T v = coef[t / per_frame] * x[t];
    y[t] = accumulate ? y[t] + v : v;
  }
}

}  // namespace
}  // namespace cvp

using namespace cvp;

extern "C" int cvp_path_cv_combine(const cvp_path_cv_desc* desc, int dtype, const void* values,
                                   int64_t num_frames, void* out_value, void* out_coef,
                                   void* stream_ptr) {
  CVP_REQUIRE(desc != nullptr && values != nullptr && out_value != nullptr, "null argument");
  CVP_REQUIRE(dtype == CVP_F32 || dtype == CVP_F64, "dtype must be CVP_F32 or CVP_F64");
  CVP_REQUIRE(desc->num_variables >= 1 && desc->num_milestones >= 2 && desc->milestones &&
                  desc->periods && desc->scales && desc->sigma > 0,
              "a path in CV space needs >= 1 variable, >= 2 milestones and sigma > 0");
  CVP_REQUIRE(desc->metric == CVP_COMBINE_PATH_PROGRESS || desc->metric == CVP_COMBINE_PATH_DEVIATION,
              "metric must be CVP_COMBINE_PATH_PROGRESS or CVP_COMBINE_PATH_DEVIATION");
  if (num_frames <= 0) return CVP_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream_ptr);
  const int nv = desc->num_variables, n = desc->num_milestones;
  std::vector<double> host((size_t)n * nv + 2 * (size_t)nv);
  std::copy(desc->milestones, desc->milestones + (size_t)n * nv, host.begin());
  for (int j = 0; j < nv; ++j) {
    CVP_REQUIRE(desc->scales[j] > 0 && desc->periods[j] >= 0, "scales must be positive, periods >= 0");
    host[(size_t)n * nv + j] = desc->periods[j];
    host[(size_t)n * nv + nv + j] = 1.0 / (desc->scales[j] * desc->scales[j]);
  }
  double* dev = nullptr;
  CVP_CUDA_CHECK(cudaMallocAsync(&dev, host.size() * sizeof(double), st));
  CVP_CUDA_CHECK(cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(double),
                                 cudaMemcpyHostToDevice, st));
  CVP_CUDA_CHECK(cudaStreamSynchronize(st));  // (`host` is pageable and goes out of scope)
  const double lambda = 1.0 / (2.0 * desc->sigma * desc->sigma);
  const int progress = desc->metric == CVP_COMBINE_PATH_PROGRESS;
  const unsigned blocks = (unsigned)div_up(num_frames, 128);
  if (dtype == CVP_F32)
    path_cv_combine_kernel<float><<<blocks, 128, 0, st>>>(
        static_cast<const float*>(values), nv, n, dev, dev + (size_t)n * nv,
        dev + (size_t)n * nv + nv, lambda, progress, num_frames, static_cast<float*>(out_value),
        static_cast<float*>(out_coef));
  else
    path_cv_combine_kernel<double><<<blocks, 128, 0, st>>>(
        static_cast<const double*>(values), nv, n, dev, dev + (size_t)n * nv,
        dev + (size_t)n * nv + nv, lambda, progress, num_frames, static_cast<double*>(out_value),
        static_cast<double*>(out_coef));
  CVP_LAUNCH_CHECK();
  CVP_CUDA_CHECK(cudaFreeAsync(dev, st));
  return CVP_OK;
}

extern "C" int cvp_axpy_frames(int dtype, const void* coef, const void* x, void* y,
                               int64_t num_frames, int64_t num_atoms, int accumulate,
                               void* stream_ptr) {
  CVP_REQUIRE(dtype == CVP_F32 || dtype == CVP_F64, "dtype must be CVP_F32 or CVP_F64");
  CVP_REQUIRE(coef != nullptr && x != nullptr && y != nullptr, "null argument");
  if (num_frames <= 0 || num_atoms <= 0) return CVP_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream_ptr);
  const int64_t per_frame = num_atoms * 3, total = num_frames * per_frame;
  const unsigned blocks = (unsigned)std::min<int64_t>(div_up(total, 256), 148 * 32);
  if (dtype == CVP_F32)
    axpy_frames_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(coef),
                                                     static_cast<const float*>(x),
                                                     static_cast<float*>(y), per_frame, accumulate,
                                                     total);
  else
    axpy_frames_kernel<double><<<blocks, 256, 0, st>>>(static_cast<const double*>(coef),
                                                      static_cast<const double*>(x),
                                                      static_cast<double*>(y), per_frame,
                                                      accumulate, total);
  CVP_LAUNCH_CHECK();
  return CVP_OK;
}
