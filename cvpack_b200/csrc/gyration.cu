// Radius of gyration (and its square) of an atom group, value + gradient, for a batch of frames.
//
// Replaces the CustomCentroidBondForce constructions of cvpack/radius_of_gyration.py:85-92 and
// radius_of_gyration_sq.py:87-89 (base_radius_of_gyration.py:22-37):
//   rg^2 = (1/n) sum_i |d_i|^2,  d_i = r_i - c  (minimum image if periodic),  c = sum_i w_i r_i
//   d rg^2/d r_k = (2/n) [ d_k - w_k sum_i d_i ],   d rg = d rg^2 / (2 rg)
// Streaming, HBM-bound: coordinates are read once per pass, sums are carried in double (the
// single-pass identity sum|r-c|^2 = sum|r|^2 - 2 c.sum r + n|c|^2 cancels badly in float).
#include "common.cuh"
#include "stream_fused.cuh"

namespace cvp {
namespace {

constexpr int GYR_THREADS = 256;
constexpr int GYR_ATOMS_PER_BLOCK = 4096;

// pass 1: per-(frame, slice) partial sums {sum w r (3), sum r (3), sum |r|^2}
template <typename T>
__global__ void __launch_bounds__(GYR_THREADS)
    gyration_sums_kernel(const T* __restrict__ pos, int64_t num_atoms,
                         const int32_t* __restrict__ group, const double* __restrict__ weights,
                         int n, int first_atom, int nslice, double* __restrict__ partials) {
  __shared__ double scratch[7 * (GYR_THREADS / 32)];
  const int64_t b = blockIdx.x / nslice;
  const int slice = blockIdx.x - (int)b * nslice;
  const int begin = slice * GYR_ATOMS_PER_BLOCK;
  const int end = min(n, begin + GYR_ATOMS_PER_BLOCK);
  const T* frame = pos + b * num_atoms * 3;
  double v[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int i = begin + threadIdx.x; i < end; i += GYR_THREADS) {
    const int atom = group ? group[i] : first_atom + i;
    const double x = frame[3 * (int64_t)atom], y = frame[3 * (int64_t)atom + 1],
                 z = frame[3 * (int64_t)atom + 2];
    const double w = weights[i];
    v[0] += w * x;
    v[1] += w * y;
    v[2] += w * z;
    v[3] += x;
    v[4] += y;
    v[5] += z;
    v[6] += x * x + y * y + z * z;
  }
  block_sum<7>(v, scratch);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) partials[((size_t)b * nslice + slice) * 7 + k] = v[k];
  }
}

// Per-frame record shared by the later passes.
struct GyrationFrame {
  double c[3];     // centre
  double s[3];     // sum_i d_i
  double rg2;      // squared radius of gyration
  double factor;   // gradient prefactor: 2/n (squared) or 1/(n rg)
};

// pass 2 (one warp per frame): centre; non-periodic value.
template <typename T>
__global__ void gyration_centre_kernel(const double* __restrict__ partials, int nslice, int n,
                                       int periodic, int squared, GyrationFrame* __restrict__ rec,
                                       T* __restrict__ value, int accumulate, int64_t num_frames) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (b >= num_frames) return;
  double sum[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    double v = 0.0;
    for (int s = lane; s < nslice; s += 32) v += partials[((size_t)b * nslice + s) * 7 + k];
    sum[k] = warp_sum(v);
  }
  if (lane != 0) return;
  GyrationFrame r;
  r.c[0] = sum[0];
  r.c[1] = sum[1];
  r.c[2] = sum[2];
  if (!periodic) {
    double cc = r.c[0] * r.c[0] + r.c[1] * r.c[1] + r.c[2] * r.c[2];
    double cs = r.c[0] * sum[3] + r.c[1] * sum[4] + r.c[2] * sum[5];
    r.rg2 = fmax(0.0, (sum[6] - 2.0 * cs + n * cc) / n);
    r.s[0] = sum[3] - n * r.c[0];
    r.s[1] = sum[4] - n * r.c[1];
    r.s[2] = sum[5] - n * r.c[2];
    double rg = sqrt(r.rg2);
    r.factor = squared ? 2.0 / n : (rg > 0.0 ? 1.0 / (n * rg) : 0.0);
    double result = squared ? r.rg2 : rg;
    value[b] = accumulate ? T((double)value[b] + result) : T(result);
  }
  rec[b] = r;
}

// pass 3 (periodic only): partial sums {sum d (3), sum |d|^2} with d = minimum image of r - c
template <typename T, int PBC>
__global__ void __launch_bounds__(GYR_THREADS)
    gyration_periodic_sums_kernel(const T* __restrict__ pos, int64_t num_atoms,
                                  const int32_t* __restrict__ group, int n, int first_atom,
                                  int nslice, const T* __restrict__ box, int box_mode,
                                  const GyrationFrame* __restrict__ rec,
                                  double* __restrict__ partials) {
  __shared__ double scratch[4 * (GYR_THREADS / 32)];
  const int64_t b = blockIdx.x / nslice;
  const int slice = blockIdx.x - (int)b * nslice;
  const int begin = slice * GYR_ATOMS_PER_BLOCK;
  const int end = min(n, begin + GYR_ATOMS_PER_BLOCK);
  const T* frame = pos + b * num_atoms * 3;
  const Box<double> bx = [&] {
    Box<T> t = load_box(box, box_mode, b);
    Box<double> d;
    d.ax = t.ax; d.bx = t.bx; d.by = t.by; d.cx = t.cx; d.cy = t.cy; d.cz = t.cz;
    d.inv_ax = 1.0 / d.ax; d.inv_by = 1.0 / d.by; d.inv_cz = 1.0 / d.cz;
    return d;
  }();
  const double cx = rec[b].c[0], cy = rec[b].c[1], cz = rec[b].c[2];
  double v[4] = {0, 0, 0, 0};
  for (int i = begin + threadIdx.x; i < end; i += GYR_THREADS) {
    const int atom = group ? group[i] : first_atom + i;
    double dx = frame[3 * (int64_t)atom] - cx, dy = frame[3 * (int64_t)atom + 1] - cy,
           dz = frame[3 * (int64_t)atom + 2] - cz;
    min_image<PBC>(dx, dy, dz, bx);
    v[0] += dx;
    v[1] += dy;
    v[2] += dz;
    v[3] += dx * dx + dy * dy + dz * dz;
  }
  block_sum<4>(v, scratch);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) partials[((size_t)b * nslice + slice) * 4 + k] = v[k];
  }
}

template <typename T>
__global__ void gyration_periodic_finish_kernel(const double* __restrict__ partials, int nslice,
                                                int n, int squared, GyrationFrame* __restrict__ rec,
                                                T* __restrict__ value, int accumulate,
                                                int64_t num_frames) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (b >= num_frames) return;
  double sum[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double v = 0.0;
    for (int s = lane; s < nslice; s += 32) v += partials[((size_t)b * nslice + s) * 4 + k];
    sum[k] = warp_sum(v);
  }
  if (lane != 0) return;
  GyrationFrame r = rec[b];
  r.s[0] = sum[0];
  r.s[1] = sum[1];
  r.s[2] = sum[2];
  r.rg2 = sum[3] / n;
  double rg = sqrt(r.rg2);
  r.factor = squared ? 2.0 / n : (rg > 0.0 ? 1.0 / (n * rg) : 0.0);
  double result = squared ? r.rg2 : rg;
  value[b] = accumulate ? T((double)value[b] + result) : T(result);
  rec[b] = r;
}

// gradient pass: g_k = factor * (d_k - w_k S)
template <typename T, int PBC>
__global__ void __launch_bounds__(GYR_THREADS)
    gyration_gradient_kernel(const T* __restrict__ pos, int64_t num_atoms,
                             const int32_t* __restrict__ group, const double* __restrict__ weights,
                             int n, int first_atom, int nslice, const T* __restrict__ box,
                             int box_mode, const GyrationFrame* __restrict__ rec,
                             T* __restrict__ grad, int accumulate) {
  const int64_t b = blockIdx.x / nslice;
  const int slice = blockIdx.x - (int)b * nslice;
  const int begin = slice * GYR_ATOMS_PER_BLOCK;
  const int end = min(n, begin + GYR_ATOMS_PER_BLOCK);
  const T* frame = pos + b * num_atoms * 3;
  T* gframe = grad + b * num_atoms * 3;
  Box<double> bx;
  if (PBC != 0) {
    Box<T> t = load_box(box, box_mode, b);
    bx.ax = t.ax; bx.bx = t.bx; bx.by = t.by; bx.cx = t.cx; bx.cy = t.cy; bx.cz = t.cz;
    bx.inv_ax = 1.0 / bx.ax; bx.inv_by = 1.0 / bx.by; bx.inv_cz = 1.0 / bx.cz;
  }
  const GyrationFrame r = rec[b];
  for (int i = begin + threadIdx.x; i < end; i += GYR_THREADS) {
    const int64_t atom = group ? group[i] : first_atom + i;
    double dx = frame[3 * atom] - r.c[0], dy = frame[3 * atom + 1] - r.c[1],
           dz = frame[3 * atom + 2] - r.c[2];
    min_image<PBC>(dx, dy, dz, bx);
    const double w = weights[i];
    const double gx = r.factor * (dx - w * r.s[0]);
    const double gy = r.factor * (dy - w * r.s[1]);
    const double gz = r.factor * (dz - w * r.s[2]);
    if (accumulate) {
      gframe[3 * atom] += T(gx);
      gframe[3 * atom + 1] += T(gy);
      gframe[3 * atom + 2] += T(gz);
    } else {
      gframe[3 * atom] = T(gx);
      gframe[3 * atom + 1] = T(gy);
      gframe[3 * atom + 2] = T(gz);
    }
  }
}

// ---- streaming fast path: contiguous, 16-byte aligned group, no periodic box, FP32 -------------------
// 4 atoms (3 x float4) per thread iteration; weights read only for a mass-weighted centre.
constexpr int GYR_FAST_SLICE = 4096;

__device__ __forceinline__ void gyr_unpack4(const float4& a0, const float4& a1, const float4& a2,
                                            float (&x)[4], float (&y)[4], float (&z)[4]) {
  x[0] = a0.x; y[0] = a0.y; z[0] = a0.z;
  x[1] = a0.w; y[1] = a1.x; z[1] = a1.y;
  x[2] = a1.z; y[2] = a1.w; z[2] = a2.x;
  x[3] = a2.y; y[3] = a2.z; z[3] = a2.w;
}

__global__ void __launch_bounds__(GYR_THREADS)
    gyration_sums_fast_kernel(const float* __restrict__ pos, int64_t num_atoms, int first_atom, int n,
                              const float* __restrict__ weights, int nslice,
                              double* __restrict__ partials) {
  __shared__ double scratch[7 * (GYR_THREADS / 32)];
  const int64_t b = blockIdx.x / nslice;
  const int s = blockIdx.x - (int)b * nslice;
  const float4* px = reinterpret_cast<const float4*>(pos + (b * num_atoms + first_atom) * 3);
  const float4* pw = reinterpret_cast<const float4*>(weights);
  const int g_begin = s * (GYR_FAST_SLICE / 4);
  const int g_end = min(n / 4, g_begin + GYR_FAST_SLICE / 4);
  const double uniform = 1.0 / n;
  double v[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int g = g_begin + threadIdx.x; g < g_end; g += GYR_THREADS) {
    const float4 a0 = __ldg(px + 3 * g), a1 = __ldg(px + 3 * g + 1), a2 = __ldg(px + 3 * g + 2);
    float x[4], y[4], z[4];
    gyr_unpack4(a0, a1, a2, x, y, z);
    float w[4];
    if (weights != nullptr) {
      const float4 wv = __ldg(pw + g);
      w[0] = wv.x; w[1] = wv.y; w[2] = wv.z; w[3] = wv.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double xd = x[k], yd = y[k], zd = z[k];
      if (weights != nullptr) {
        v[0] += (double)w[k] * xd;
        v[1] += (double)w[k] * yd;
        v[2] += (double)w[k] * zd;
      }
      v[3] += xd;
      v[4] += yd;
      v[5] += zd;
      v[6] += xd * xd + yd * yd + zd * zd;
    }
  }
  if (weights == nullptr) {
    v[0] = v[3] * uniform;
    v[1] = v[4] * uniform;
    v[2] = v[5] * uniform;
  }
  block_sum<7>(v, scratch);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) partials[((size_t)b * nslice + s) * 7 + k] = v[k];
  }
}

__global__ void __launch_bounds__(GYR_THREADS)
    gyration_gradient_fast_kernel(const float* __restrict__ pos, int64_t num_atoms, int first_atom,
                                  int n, const float* __restrict__ weights, int nslice,
                                  const GyrationFrame* __restrict__ rec, float* __restrict__ grad,
                                  int accumulate) {
  const int64_t b = blockIdx.x / nslice;
  const int s = blockIdx.x - (int)b * nslice;
  const float4* px = reinterpret_cast<const float4*>(pos + (b * num_atoms + first_atom) * 3);
  const float4* pw = reinterpret_cast<const float4*>(weights);
  float4* pg = reinterpret_cast<float4*>(grad + (b * num_atoms + first_atom) * 3);
  const int g_begin = s * (GYR_FAST_SLICE / 4);
  const int g_end = min(n / 4, g_begin + GYR_FAST_SLICE / 4);
  const GyrationFrame r = rec[b];
  const float f = (float)r.factor;
  const float c[3] = {(float)r.c[0], (float)r.c[1], (float)r.c[2]};
  const float sw[3] = {(float)r.s[0], (float)r.s[1], (float)r.s[2]};
  const float uniform = 1.0f / n;
  for (int g = g_begin + threadIdx.x; g < g_end; g += GYR_THREADS) {
    const float4 a0 = __ldg(px + 3 * g), a1 = __ldg(px + 3 * g + 1), a2 = __ldg(px + 3 * g + 2);
    float x[4], y[4], z[4], w[4] = {uniform, uniform, uniform, uniform};
    gyr_unpack4(a0, a1, a2, x, y, z);
    if (weights != nullptr) {
      const float4 wv = __ldg(pw + g);
      w[0] = wv.x; w[1] = wv.y; w[2] = wv.z; w[3] = wv.w;
    }
    float gx[4], gy[4], gz[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      gx[k] = f * ((x[k] - c[0]) - w[k] * sw[0]);
      gy[k] = f * ((y[k] - c[1]) - w[k] * sw[1]);
      gz[k] = f * ((z[k] - c[2]) - w[k] * sw[2]);
    }
    float4 o0 = make_float4(gx[0], gy[0], gz[0], gx[1]);
    float4 o1 = make_float4(gy[1], gz[1], gx[2], gy[2]);
    float4 o2 = make_float4(gz[2], gx[3], gy[3], gz[3]);
    if (accumulate) {
      const float4 p0 = pg[3 * g], p1 = pg[3 * g + 1], p2 = pg[3 * g + 2];
      o0.x += p0.x; o0.y += p0.y; o0.z += p0.z; o0.w += p0.w;
      o1.x += p1.x; o1.y += p1.y; o1.z += p1.z; o1.w += p1.w;
      o2.x += p2.x; o2.y += p2.y; o2.z += p2.z; o2.w += p2.w;
    }
    __stcs(pg + 3 * g, o0);
    __stcs(pg + 3 * g + 1, o1);
    __stcs(pg + 3 * g + 2, o2);
  }
}

// One-read fused evaluation of a contiguous FP32 group without a periodic box (stream_fused.cuh).
// WEIGHTED: the centre is a weighted mean (weights = the per-atom auxiliary float); otherwise the
// geometric centre, for which sum_i d_i vanishes and the gradient is f (r_k - c).
template <bool WEIGHTED>
struct GyrationFusedPolicy {
  static constexpr int NS = WEIGHTED ? 7 : 4;
  static constexpr int AUX = WEIGHTED ? 1 : 0;
  static constexpr int AUXD = 0;
  static constexpr int FRAME_FLOATS = 7;
  struct Params {
    int n;
    int squared;
  };
  struct Frame {
    float c[3];
    float sw[3];
    float f;
    float pad[9];
  };
  __device__ static __forceinline__ void accumulate(const Params&, const float* aux, const double2*,
                                                    const float (&x)[4], const float (&y)[4],
                                                    const float (&z)[4], double (&v)[NS]) {
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    if (WEIGHTED) {
      const float4 wv = *reinterpret_cast<const float4*>(aux);  // shared memory
      w[0] = wv.x; w[1] = wv.y; w[2] = wv.z; w[3] = wv.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double xd = x[k], yd = y[k], zd = z[k];
      v[0] += xd;
      v[1] += yd;
      v[2] += zd;
      v[3] = fma(xd, xd, v[3]);
      v[3] = fma(yd, yd, v[3]);
      v[3] = fma(zd, zd, v[3]);
      if (WEIGHTED) {
        const double wd = w[k];
        v[NS - 3] += wd * xd;
        v[NS - 2] += wd * yd;
        v[NS - 1] += wd * zd;
      }
    }
  }
  __device__ static __noinline__ void complete(const Params& p, const double (&sums)[NS],
                                               Frame* frame, float* value, int accumulate) {
    const double n = p.n;
    double c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) c[k] = WEIGHTED ? sums[NS - 3 + k] : sums[k] / n;
    const double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    const double cs = c[0] * sums[0] + c[1] * sums[1] + c[2] * sums[2];
    const double rg2 = fmax(0.0, (sums[3] - 2.0 * cs + n * cc) / n);
    const double rg = sqrt(rg2);
    Frame fr;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      fr.c[k] = (float)c[k];
      fr.sw[k] = WEIGHTED ? (float)(sums[k] - n * c[k]) : 0.f;
    }
    fr.f = (float)(p.squared ? 2.0 / n : (rg > 0.0 ? 1.0 / (n * rg) : 0.0));
#pragma unroll
    for (int k = 0; k < 9; ++k) fr.pad[k] = 0.f;
    *frame = fr;
    const double result = p.squared ? rg2 : rg;
    *value = accumulate ? (float)((double)*value + result) : (float)result;
  }
  __device__ static __forceinline__ void gradient(const Params&, const Frame& fr, const float* aux,
                                                  const float (&x)[4], const float (&y)[4],
                                                  const float (&z)[4], float (&gx)[4],
                                                  float (&gy)[4], float (&gz)[4]) {
    float w[4] = {0.f, 0.f, 0.f, 0.f};
    if (WEIGHTED) {
      const float4 wv = *reinterpret_cast<const float4*>(aux);
      w[0] = wv.x; w[1] = wv.y; w[2] = wv.z; w[3] = wv.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (WEIGHTED) {
        gx[k] = fr.f * ((x[k] - fr.c[0]) - w[k] * fr.sw[0]);
        gy[k] = fr.f * ((y[k] - fr.c[1]) - w[k] * fr.sw[1]);
        gz[k] = fr.f * ((z[k] - fr.c[2]) - w[k] * fr.sw[2]);
      } else {
        gx[k] = fr.f * (x[k] - fr.c[0]);
        gy[k] = fr.f * (y[k] - fr.c[1]);
        gz[k] = fr.f * (z[k] - fr.c[2]);
      }
    }
  }
};

struct GyrationCV : cvp_cv {
  bool uniform_weights = false;
  std::vector<float> weights_f32;
  DeviceBuffer d_weights_f32;
  std::vector<int32_t> group;
  std::vector<double> weights;
  int periodic = 0, squared = 0;
  bool contiguous = false;
  DeviceBuffer d_group, d_weights;
  ~GyrationCV() override {
    d_group.release();
    d_weights.release();
    d_weights_f32.release();
  }
  int upload() override {
    int status = cvp::upload(d_group, group);
    if (status != CVP_OK) return status;
    if ((status = cvp::upload(d_weights_f32, weights_f32)) != CVP_OK) return status;
    return cvp::upload(d_weights, weights);
  }
  int nslice() const { return (int)div_up((int64_t)group.size(), GYR_ATOMS_PER_BLOCK); }
  int use_fused = 1;    // option "fused": one-read persistent kernel for the streaming fast path
  // measured on the 100k-atom batch: rounds of 1024-atom items (27 MB), gradient one round later
  int fused_lag = 1;    // option "fused_lag"
  int fused_slice = 1024;  // option "fused_slice"
  int fused_slots = 0;  // option "fused_slots"
  int fused_hints = 3;  // option "fused_hints"
  size_t workspace_per_frame(int, bool) const override {
    return std::max((size_t)nslice() * 7 * sizeof(double) + sizeof(GyrationFrame) + 64,
                    FusedLaunch<GyrationFusedPolicy<true>>::workspace_per_frame((int)group.size()));
  }
  size_t workspace_fixed(int) const override {
    return 4 * 256 + FusedLaunch<GyrationFusedPolicy<true>>::workspace_fixed();
  }
  int set_option(const char* name, int value) override {
    const std::string key(name);
    if (key == "fused") {
      std::swap(use_fused, value);
      return value;
    }
    if (key == "fused_lag") {
      std::swap(fused_lag, value);
      return value;
    }
    if (key == "fused_slice") {
      std::swap(fused_slice, value);
      return value;
    }
    if (key == "fused_slots") {
      std::swap(fused_slots, value);
      return value;
    }
    if (key == "fused_hints") {
      std::swap(fused_hints, value);
      return value;
    }
    return cvp_cv::set_option(name, value);
  }
  template <typename Policy>
  int run_fused(const FusedGeom& geom, const EvalArgs& a, void* ws, const float* aux) {
    typename Policy::Params params;
    params.n = (int)group.size();
    params.squared = squared;
    const bool accumulate = (a.flags & CVP_FLAG_ACCUMULATE) != 0;
    const bool dense = (int64_t)group.size() == a.num_atoms;
    if (a.grad != nullptr && !accumulate && !dense)
      CVP_CUDA_CHECK(cudaMemsetAsync(a.grad, 0, (size_t)a.num_frames * a.num_atoms * 3 * sizeof(float),
                                     a.stream));
    timer.begin("gyration_fused", a.stream);
    int status = FusedLaunch<Policy>::launch(static_cast<const float*>(a.positions), aux, nullptr,
                                             geom, sm_count(), params, ws, static_cast<float*>(a.value),
                                             static_cast<float*>(a.grad), accumulate, a.stream);
    timer.end(a.stream);
    return status;
  }
  int64_t max_frames_per_run() const override {
    return std::max<int64_t>(1, (int64_t(1) << 31) / nslice() - 1);
  }

  template <typename T, int PBC>
  int run_impl(const EvalArgs& a, void* ws) {
    const int64_t B = a.num_frames, N = a.num_atoms;
    const int n = (int)group.size();
    const int ns = nslice();
    const bool accumulate = (a.flags & CVP_FLAG_ACCUMULATE) != 0;
    const T* pos = static_cast<const T*>(a.positions);
    const T* box = static_cast<const T*>(a.box);
    T* value = static_cast<T*>(a.value);
    T* grad = static_cast<T*>(a.grad);
    cudaStream_t st = a.stream;
    Carver carve(ws);
    double* partials = carve.take<double>((size_t)B * ns * 7);
    GyrationFrame* rec = carve.take<GyrationFrame>((size_t)B);
    const int32_t* idx = contiguous ? nullptr : static_cast<const int32_t*>(d_group.ptr);
    const double* w = static_cast<const double*>(d_weights.ptr);
    const int first = group.front();
    const unsigned grid = (unsigned)(B * ns);
    const int warp_blocks = (int)div_up(B * 32, 128);

    // streaming fast path: FP32, no box, contiguous group starting on a 16-byte boundary
    const bool use_fast = PBC == 0 && std::is_same<T, float>::value && contiguous && n % 4 == 0 &&
                          first % 4 == 0 && N % 4 == 0 && aligned16(a.positions, a.grad);
    const int fast_ns = (int)div_up(n, GYR_FAST_SLICE);
    const float* wf = uniform_weights ? nullptr : static_cast<const float*>(d_weights_f32.ptr);
    if (use_fast && use_fused && B >= 64) {
      FusedGeom geom;
      if (uniform_weights) {
        if (FusedLaunch<GyrationFusedPolicy<false>>::plan(n, N, first, B, grad != nullptr, fused_lag, fused_slice, fused_slots, fused_hints, sm_count(), &geom)) {
          return run_fused<GyrationFusedPolicy<false>>(geom, a, ws, nullptr);
        }
      } else if (FusedLaunch<GyrationFusedPolicy<true>>::plan(n, N, first, B, grad != nullptr,
                                                              fused_lag, fused_slice, fused_slots, fused_hints, sm_count(),
                                                              &geom)) {
        return run_fused<GyrationFusedPolicy<true>>(geom, a, ws, wf);
      }
    }
    timer.begin("gyration_sums", st);
    if (use_fast)
      gyration_sums_fast_kernel<<<(unsigned)(B * fast_ns), GYR_THREADS, 0, st>>>(
          reinterpret_cast<const float*>(pos), N, first, n, wf, fast_ns, partials);
    else
      gyration_sums_kernel<T><<<grid, GYR_THREADS, 0, st>>>(pos, N, idx, w, n, first, ns, partials);
    timer.end(st);
    CVP_LAUNCH_CHECK();
    const int ns_used = use_fast ? fast_ns : ns;
    gyration_centre_kernel<T><<<warp_blocks, 128, 0, st>>>(partials, ns_used, n, PBC != 0, squared, rec,
                                                            value, accumulate ? 1 : 0, B);
    CVP_LAUNCH_CHECK();
    if (PBC != 0) {
      gyration_periodic_sums_kernel<T, PBC><<<grid, GYR_THREADS, 0, st>>>(
          pos, N, idx, n, first, ns, box, a.box_mode, rec, partials);
      CVP_LAUNCH_CHECK();
      gyration_periodic_finish_kernel<T><<<warp_blocks, 128, 0, st>>>(
          partials, ns, n, squared, rec, value, accumulate ? 1 : 0, B);
      CVP_LAUNCH_CHECK();
    }
    if (grad == nullptr) return CVP_OK;
    const bool dense = contiguous && n == N;
    if (!accumulate && !dense)
      CVP_CUDA_CHECK(cudaMemsetAsync(grad, 0, (size_t)B * N * 3 * sizeof(T), st));
    timer.begin("gyration_gradient", st);
    if (use_fast)
      gyration_gradient_fast_kernel<<<(unsigned)(B * fast_ns), GYR_THREADS, 0, st>>>(
          reinterpret_cast<const float*>(pos), N, first, n, wf, fast_ns, rec,
          reinterpret_cast<float*>(grad), accumulate ? 1 : 0);
    else
      gyration_gradient_kernel<T, PBC><<<grid, GYR_THREADS, 0, st>>>(
          pos, N, idx, w, n, first, ns, box, a.box_mode, rec, grad, accumulate ? 1 : 0);
    timer.end(st);
    CVP_LAUNCH_CHECK();
    return CVP_OK;
  }

  template <typename T>
  int run_type(const EvalArgs& a, void* ws) {
    if (!periodic) return run_impl<T, 0>(a, ws);
    CVP_REQUIRE(a.box != nullptr && a.box_mode != CVP_BOX_NONE,
                "this collective variable uses periodic boundary conditions but no box was given");
    return (a.flags & CVP_FLAG_TRICLINIC) ? run_impl<T, 2>(a, ws) : run_impl<T, 1>(a, ws);
  }
  int run(const EvalArgs& a, void* ws) override {
    return a.dtype == CVP_F32 ? run_type<float>(a, ws) : run_type<double>(a, ws);
  }
};

}  // namespace

int gyration_create(const cvp_gyration_desc* desc, cvp_cv** out) {
  CVP_REQUIRE(desc != nullptr && out != nullptr, "null argument");
  CVP_REQUIRE(desc->num_atoms > 0 && desc->n > 0 && desc->group && desc->weights, "empty group");
  auto cv = new GyrationCV();
  cv->num_atoms = desc->num_atoms;
  cv->group.assign(desc->group, desc->group + desc->n);
  cv->weights.assign(desc->weights, desc->weights + desc->n);
  cv->periodic = desc->periodic;
  cv->squared = desc->squared;
  std::vector<char> seen(desc->num_atoms, 0);
  bool contiguous = true;
  for (int i = 0; i < desc->n; ++i) {
    int32_t a = cv->group[i];
    if (a < 0 || a >= desc->num_atoms || seen[a]) {
      delete cv;
      set_error(a < 0 || a >= desc->num_atoms ? "atom index out of range in group"
                                              : "repeated atom in group");
      return CVP_ERR_INVALID;
    }
    seen[a] = 1;
    if (a != cv->group[0] + i) contiguous = false;
  }
  cv->contiguous = contiguous;
  cv->uniform_weights = true;
  for (int i = 0; i < desc->n; ++i)
    if (cv->weights[i] != cv->weights[0]) cv->uniform_weights = false;
  cv->weights_f32.assign(cv->weights.begin(), cv->weights.end());
  cv->active_atoms = cv->group;
  std::sort(cv->active_atoms.begin(), cv->active_atoms.end());
  *out = cv;
  return CVP_OK;
}

}  // namespace cvp
