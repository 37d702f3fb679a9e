// Shared device/host helpers of the cvpack_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <type_traits>
#include <utility>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cvpack_b200.h"

namespace cvp {

// ---- errors ------------------------------------------------------------------------------------
void set_error(const std::string& message);
extern std::atomic<int64_t> g_launch_count;

#define CVP_CUDA_CHECK(expr)                                                                      \
  do {                                                                                            \
    cudaError_t cvp_err__ = (expr);                                                               \
    if (cvp_err__ != cudaSuccess) {                                                               \
      ::cvp::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(cvp_err__));         \
      return CVP_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

#define CVP_REQUIRE(cond, message)                                                                \
  do {                                                                                            \
    if (!(cond)) {                                                                                \
      ::cvp::set_error(message);                                                                  \
      return CVP_ERR_INVALID;                                                                     \
    }                                                                                             \
  } while (0)

#define CVP_LAUNCH_CHECK()                                                                        \
  do {                                                                                            \
    ::cvp::g_launch_count.fetch_add(1, std::memory_order_relaxed);                                \
    CVP_CUDA_CHECK(cudaGetLastError());                                                           \
  } while (0)

// ---- device buffers ----------------------------------------------------------------------------
// Growable device allocation owned by a spec.
struct DeviceBuffer {
  void* ptr = nullptr;
  size_t bytes = 0;
  int reserve(size_t want) {
    if (want <= bytes) return CVP_OK;
    if (ptr) CVP_CUDA_CHECK(cudaFree(ptr));
    ptr = nullptr;
    bytes = 0;
    CVP_CUDA_CHECK(cudaMalloc(&ptr, want));
    bytes = want;
    return CVP_OK;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
  }
};

template <typename U>
int upload(DeviceBuffer& buffer, const std::vector<U>& host) {
  size_t bytes = host.size() * sizeof(U);
  int status = buffer.reserve(bytes ? bytes : sizeof(U));
  if (status != CVP_OK) return status;
  if (bytes) CVP_CUDA_CHECK(cudaMemcpy(buffer.ptr, host.data(), bytes, cudaMemcpyHostToDevice));
  return CVP_OK;
}

// The vectorised (float4) and TMA bulk paths need 16-byte aligned base addresses; a C-ABI caller or a
// tensor view with a storage offset may hand in less (null counts as aligned).
inline bool aligned16(const void* a, const void* b = nullptr) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u) == 0;
}

// Linear carve-up of a workspace allocation (256-byte aligned pieces).
struct Carver {
  char* base;
  size_t offset = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename U>
  U* take(size_t count) {
    U* result = reinterpret_cast<U*>(base + offset);
    offset += (count * sizeof(U) + 255) & ~size_t(255);
    return result;
  }
};
inline size_t carve_bytes(size_t count, size_t elem) { return (count * elem + 255) & ~size_t(255); }

// ---- the opaque spec ---------------------------------------------------------------------------
struct EvalArgs {
  int dtype;
  const void* positions;  // device [B][N][3]
  const void* box;        // device [9] or [B][9] or null
  int box_mode;
  int64_t num_frames;
  int64_t num_atoms;
  void* value;  // device [B]
  void* grad;   // device [B][N][3] or null
  int flags;
  cudaStream_t stream;
};

}  // namespace cvp

namespace cvp {
// Optional per-kernel device timing (CUDA events on the launching stream), switched on with
// cvp_set_option(cv, "time_kernels", 1) and read back with cvp_get_stat(cv, "<kernel>_ms", ...).
struct KernelTimer {
  struct Span {
    std::string name;
    cudaEvent_t start, stop;
  };
  bool enabled = false;
  std::vector<Span> spans;
  void begin(const char* name, cudaStream_t stream) {
    if (!enabled) return;
    Span span{name, nullptr, nullptr};
    cudaEventCreate(&span.start);
    cudaEventCreate(&span.stop);
    cudaEventRecord(span.start, stream);
    spans.push_back(span);
  }
  void end(cudaStream_t stream) {
    if (!enabled || spans.empty()) return;
    cudaEventRecord(spans.back().stop, stream);
  }
  // total milliseconds and number of spans recorded under `name`; consumes them
  void collect(const std::string& name, double* ms, double* count) {
    *ms = 0.0;
    *count = 0.0;
    std::vector<Span> keep;
    for (auto& span : spans) {
      if (span.name != name) {
        keep.push_back(span);
        continue;
      }
      cudaEventSynchronize(span.stop);
      float elapsed = 0.f;
      cudaEventElapsedTime(&elapsed, span.start, span.stop);
      *ms += elapsed;
      *count += 1.0;
      cudaEventDestroy(span.start);
      cudaEventDestroy(span.stop);
    }
    spans.swap(keep);
  }
  void clear() {
    for (auto& span : spans) {
      cudaEventDestroy(span.start);
      cudaEventDestroy(span.stop);
    }
    spans.clear();
  }
};
}  // namespace cvp

namespace cvp {
// Device staging buffers, streams and events of the host-buffer path (cvp_evaluate_host); they live
// as long as the spec so that repeated calls neither allocate nor free device memory.
struct HostPipeline {
  static constexpr int SLOTS = 4;  // two in use unless the spec can run two chunks at once
  DeviceBuffer pos[SLOTS], grad[SLOTS], value[SLOTS], box[SLOTS], emass[SLOTS], shared_box, masses;
  cudaStream_t stream[SLOTS] = {};
  cudaEvent_t kernels_done[SLOTS] = {};
  int ensure() {
    for (int k = 0; k < SLOTS; ++k) {
      if (!stream[k]) CVP_CUDA_CHECK(cudaStreamCreateWithFlags(&stream[k], cudaStreamNonBlocking));
      if (!kernels_done[k])
        CVP_CUDA_CHECK(cudaEventCreateWithFlags(&kernels_done[k], cudaEventDisableTiming));
    }
    return CVP_OK;
  }
  void release() {
    for (int k = 0; k < SLOTS; ++k) {
      if (stream[k]) {
        cudaStreamSynchronize(stream[k]);
        cudaStreamDestroy(stream[k]);
        stream[k] = nullptr;
      }
      if (kernels_done[k]) {
        cudaEventDestroy(kernels_done[k]);
        kernels_done[k] = nullptr;
      }
      pos[k].release();
      grad[k].release();
      value[k].release();
      box[k].release();
      emass[k].release();
    }
    shared_box.release();
    masses.release();
  }
};
}  // namespace cvp

struct cvp_cv {
  int num_atoms = 0;
  cvp::KernelTimer timer;
  cvp::HostPipeline host;
  int device = -1;  // device the spec's arrays live on (-1: not uploaded yet)
  int64_t workspace_limit = int64_t(4) << 30;  // per spec and evaluate call (cvp_set_workspace_limit)
  std::vector<int32_t> active_atoms;  // sorted, unique
  cvp::DeviceBuffer workspace;
  // A second workspace for the host path: specs whose run() touches nothing but its workspace and
  // constant tables (concurrent_runs()) get one per pipeline slot, so that the kernels of
  // consecutive chunks may overlap (the tail of one launch fills with the CTAs of the next).
  cvp::DeviceBuffer workspace_alt;
  int workspace_select = 0;  // which of the two cvp_evaluate uses (set by the host path only)
  virtual bool concurrent_runs() const { return false; }
  cvp::DeviceBuffer grad_scratch;  // cvp_evaluate_with_mass: gradients that never leave the library
  virtual ~cvp_cv() {
    timer.clear();
    host.release();
    workspace.release();
    workspace_alt.release();
    grad_scratch.release();
  }
  // Upload index/parameter arrays to the current device.
  virtual int upload() = 0;
  // Workspace bytes needed per frame and independent of the frame count.
  virtual size_t workspace_per_frame(int dtype, bool need_grad) const = 0;
  virtual size_t workspace_fixed(int dtype) const { return 0; }
  // Cutoff of a sum that uses the minimum-image convention (0: none): OpenMM refuses to run such a
  // force in a box whose smallest edge is less than twice the cutoff.
  virtual double periodic_cutoff() const { return 0.0; }
  // Largest number of frames a single run() may be given.
  virtual int64_t max_frames_per_run() const { return 1 << 20; }
  // Evaluate frames [0, args.num_frames) of the given (already offset) pointers.
  virtual int run(const cvp::EvalArgs& args, void* workspace) = 0;
  virtual int set_option(const char* name, int value) {
    if (std::string(name) == "time_kernels") {
      int previous = timer.enabled ? 1 : 0;
      timer.enabled = value != 0;
      if (!timer.enabled) timer.clear();
      return previous;
    }
    return -1;
  }
  // Named statistics: "<kernel>_ms" / "<kernel>_launches" from the kernel timer.
  virtual int get_stat(const char* name, double* out) {
    std::string key(name);
    double ms = 0.0, count = 0.0;
    auto ends_with = [&](const std::string& suffix) {
      return key.size() > suffix.size() &&
             key.compare(key.size() - suffix.size(), suffix.size(), suffix) == 0;
    };
    if (ends_with("_ms")) {
      timer.collect(key.substr(0, key.size() - 3), &ms, &count);
      out[0] = ms;
      out[1] = count;
      return CVP_OK;
    }
    return CVP_ERR_INVALID;
  }
};

namespace cvp {

// ---- device math -------------------------------------------------------------------------------
template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  using type = float4;
};
template <>
struct Vec4<double> {
  using type = double4;
};

template <typename T>
__device__ __forceinline__ T rsqrt_t(T x);
template <>
__device__ __forceinline__ float rsqrt_t<float>(float x) {
  return rsqrtf(x);
}
template <>
__device__ __forceinline__ double rsqrt_t<double>(double x) {
  return 1.0 / sqrt(x);
}
template <typename T>
__device__ __forceinline__ T sqrt_t(T x);
template <>
__device__ __forceinline__ float sqrt_t<float>(float x) {
  return sqrtf(x);
}
template <>
__device__ __forceinline__ double sqrt_t<double>(double x) {
  return sqrt(x);
}
template <typename T>
__device__ __forceinline__ T rint_t(T x);
template <>
__device__ __forceinline__ float rint_t<float>(float x) {
  return rintf(x);
}
template <>
__device__ __forceinline__ double rint_t<double>(double x) {
  return rint(x);
}

// x^n for a non-negative integer n (exact same multiplication tree on host oracle is not needed:
// results are compared within floating-point tolerance).
template <typename T>
__device__ __forceinline__ T powi(T x, int n) {
  T result = T(1);
  T base = x;
  while (n > 0) {
    if (n & 1) result *= base;
    base *= base;
    n >>= 1;
  }
  return result;
}

// Periodic box of one frame, OpenMM reduced form: a=(ax,0,0) b=(bx,by,0) c=(cx,cy,cz).
template <typename T>
struct Box {
  T ax, bx, by, cx, cy, cz;
  T inv_ax, inv_by, inv_cz;
};

template <typename T>
__device__ __forceinline__ Box<T> load_box(const T* box, int box_mode, int64_t frame) {
  Box<T> result;
  const T* p = box + (box_mode == CVP_BOX_PER_FRAME ? frame * 9 : 0);
  result.ax = p[0];
  result.bx = p[3];
  result.by = p[4];
  result.cx = p[6];
  result.cy = p[7];
  result.cz = p[8];
  result.inv_ax = T(1) / result.ax;
  result.inv_by = T(1) / result.by;
  result.inv_cz = T(1) / result.cz;
  return result;
}

// Minimum-image displacement.  MODE 0: none, 1: rectangular box, 2: triclinic (reduce along c,
// then b, then a -- the order OpenMM's ReferenceForce::getDeltaRPeriodicTriclinic uses).
template <int MODE, typename T>
__device__ __forceinline__ void min_image(T& dx, T& dy, T& dz, const Box<T>& box) {
  if (MODE == 1) {
    dx -= box.ax * rint_t(dx * box.inv_ax);
    dy -= box.by * rint_t(dy * box.inv_by);
    dz -= box.cz * rint_t(dz * box.inv_cz);
  } else if (MODE == 2) {
    T s = rint_t(dz * box.inv_cz);
    dx -= s * box.cx;
    dy -= s * box.cy;
    dz -= s * box.cz;
    s = rint_t(dy * box.inv_by);
    dx -= s * box.bx;
    dy -= s * box.by;
    s = rint_t(dx * box.inv_ax);
    dx -= s * box.ax;
  }
}

// ---- reductions --------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int offset = 16; offset > 0; offset >>= 1) v += __shfl_xor_sync(0xffffffffu, v, offset);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int offset = 16; offset > 0; offset >>= 1) v += __shfl_xor_sync(0xffffffffu, v, offset);
  return v;
}

// Deterministic block sum of K doubles per thread; result valid in thread 0.  `scratch` must hold
// K * (blockDim.x/32) doubles.  All threads must call.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* scratch) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int num_warps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();  // scratch may still be in use from a previous call
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) scratch[k * num_warps + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double x = lane < num_warps ? scratch[k * num_warps + lane] : 0.0;
      v[k] = warp_sum(x);
    }
  }
}

// ---- step functions S(x) -----------------------------------------------------------------------
struct StepParams {
  int kind;
  int n;  // exponent n (p0)
  int m;  // exponent m (p1), CVP_STEP_PLUMED only
};

// S(x) and dS/dx for x >= 0.
template <typename T>
__device__ __forceinline__ void step_function(const StepParams& p, T x, T& s, T& ds) {
  switch (p.kind) {
    case CVP_STEP_RATIONAL: {
      T xn1 = powi(x, p.n - 1);
      T inv = T(1) / (T(1) + xn1 * x);
      s = inv;
      ds = -T(p.n) * xn1 * inv * inv;
      break;
    }
    case CVP_STEP_HEAVISIDE: {
      s = x <= T(1) ? T(1) : T(0);
      ds = T(0);
      break;
    }
    case CVP_STEP_RATIONAL_PAIR: {
      T xn1 = powi(x, p.n - 1);
      T xn = xn1 * x;
      T num = T(1) + xn;
      T den = num + xn * xn;
      T inv = T(1) / den;
      s = num * inv;
      // d/dx [num/den] = (num' den - num den')/den^2, num' = n x^(n-1), den' = num' (1 + 2 x^n)
      T dnum = T(p.n) * xn1;
      ds = (dnum * den - num * dnum * (T(1) + T(2) * xn)) * inv * inv;
      break;
    }
    default: {  // CVP_STEP_PLUMED: (1 - x^n)/(1 - x^m)
      // Evaluated as the ratio of the geometric sums P_k(x) = 1 + x + ... + x^(k-1)
      // ((1 - x^k) = (1 - x) P_k(x)): every term is positive for x >= 0, so there is neither the
      // removable singularity at x = 1 nor the cancellation of the literal form around it (which
      // costs FP32 three digits at |x - 1| ~ 1e-4).  Horner with the derivative carried along.
      const int hi = p.n > p.m ? p.n : p.m;
      T poly = T(1), dpoly = T(0), pn = T(1), dpn = T(0), pm = T(1), dpm = T(0);
      for (int k = 2; k <= hi; ++k) {
        dpoly = fma(dpoly, x, poly);
        poly = fma(poly, x, T(1));
        if (k == p.n) { pn = poly; dpn = dpoly; }
        if (k == p.m) { pm = poly; dpm = dpoly; }
      }
      const T inv = T(1) / pm;
      s = pn * inv;
      ds = (dpn * pm - pn * dpm) * inv * inv;
      break;
    }
  }
}

// ---- TMA (bulk async copy) + mbarrier, raw PTX -------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared bulk copy (UBLKCP); bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_addr(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
      : "memory");
}

inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Number of SMs of the current device (cached per device).
inline int sm_count() {
  static int cached[64] = {0};
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess || device < 0 || device >= 64) return 148;
  if (cached[device] == 0) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0)
      sms = 148;
    cached[device] = sms;
  }
  return cached[device];
}

}  // namespace cvp
