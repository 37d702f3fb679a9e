// extern "C" entry points of libcvpack_b200.so (see include/cvpack_b200.h).
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace cvp {

std::atomic<int64_t> g_launch_count{0};
static thread_local std::string t_error;
void set_error(const std::string& message) { t_error = message; }

int pair_sum_create(const cvp_pair_sum_desc*, cvp_cv**);
int centroid_pair_sum_create(const cvp_centroid_pair_sum_desc*, cvp_cv**);
int rmsd_create(const cvp_rmsd_desc*, cvp_cv**);
int gyration_create(const cvp_gyration_desc*, cvp_cv**);
int bonded_create(const cvp_bonded_desc*, cvp_cv**);
int effective_mass(int dtype, const void* grad, const double* masses, int64_t B, int64_t N,
                   void* emass, cudaStream_t stream);
int measure_fp32_peak(double* tflops);
int measure_fp64_peak(double* tflops);

namespace {

int prepare(cvp_cv* cv) {
  int device = -1;
  CVP_CUDA_CHECK(cudaGetDevice(&device));
  if (cv->device == device) return CVP_OK;
  if (cv->device != -1) {
    set_error("this spec was uploaded to another device; build one spec per device");
    return CVP_ERR_UNSUPPORTED;
  }
  cudaDeviceProp prop;
  CVP_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("libcvpack_b200 is built for sm_100a only; device is sm_" +
              std::to_string(prop.major) + std::to_string(prop.minor));
    return CVP_ERR_NO_DEVICE;
  }
  int status = cv->upload();
  if (status != CVP_OK) return status;
  cv->device = device;
  return CVP_OK;
}

size_t elem_size(int dtype) { return dtype == CVP_F32 ? 4 : 8; }

// Entry points that take a device ordinal switch to it for the duration of the call only: the
// caller's current device (torch's, in the Python host) is restored on every return path.
struct DeviceGuard {
  int previous = -1;
  cudaError_t status;
  explicit DeviceGuard(int device) {
    status = cudaGetDevice(&previous);
    if (status == cudaSuccess && previous != device) status = cudaSetDevice(device);
  }
  ~DeviceGuard() {
    int now = -1;
    if (previous >= 0 && cudaGetDevice(&now) == cudaSuccess && now != previous) cudaSetDevice(previous);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

int check_args(cvp_cv* cv, int dtype, const void* positions, int box_mode, int64_t B, int64_t N,
               void* value) {
  CVP_REQUIRE(cv != nullptr, "null spec");
  CVP_REQUIRE(dtype == CVP_F32 || dtype == CVP_F64, "dtype must be CVP_F32 or CVP_F64");
  CVP_REQUIRE(B >= 0, "negative frame count");
  CVP_REQUIRE(B == 0 || (positions != nullptr && value != nullptr),
              "null positions or value buffer");
  CVP_REQUIRE(N == cv->num_atoms, "the batch has " + std::to_string(N) +
                                      " atoms per frame but the spec was built for " +
                                      std::to_string(cv->num_atoms));
  CVP_REQUIRE(box_mode >= CVP_BOX_NONE && box_mode <= CVP_BOX_PER_FRAME, "bad box mode");
  return CVP_OK;
}

// frames per run() call for a given workspace budget
int64_t frames_per_chunk(cvp_cv* cv, int dtype, bool need_grad, int64_t B) {
  size_t per_frame = std::max<size_t>(1, cv->workspace_per_frame(dtype, need_grad));
  size_t fixed = cv->workspace_fixed(dtype);
  int64_t budget = cv->workspace_limit > (int64_t)fixed ? cv->workspace_limit - (int64_t)fixed : 0;
  int64_t frames = std::max<int64_t>(1, budget / (int64_t)per_frame);
  frames = std::min(frames, cv->max_frames_per_run());
  frames = std::min<int64_t>(frames, std::max<int64_t>(B, 1));
  // a short last chunk leaves most SMs idle in kernels that give a CTA a whole frame or a tile of 32
  // frames (PathInRMSDSpace: 928 + 96 frames cost 8 % more than 2 x 512): split evenly then.  A
  // remainder of at least half a chunk is left alone (the pair sweep measured 3 % slower with
  // 3 x 1366 frames than with 1483 + 1483 + 1130: its launches end in a long tail either way).
  const int64_t total = std::max<int64_t>(B, 1);
  const int64_t rest = total % frames;
  if (rest == 0 || 2 * rest >= frames) return frames;
  const int64_t chunks = (total + frames - 1) / frames;
  return (total + chunks - 1) / chunks;
}

}  // namespace
}  // namespace cvp

using namespace cvp;

extern "C" {

int cvp_pair_sum_create(const cvp_pair_sum_desc* desc, cvp_cv** out) {
  return pair_sum_create(desc, out);
}
int cvp_centroid_pair_sum_create(const cvp_centroid_pair_sum_desc* desc, cvp_cv** out) {
  return centroid_pair_sum_create(desc, out);
}
int cvp_rmsd_create(const cvp_rmsd_desc* desc, cvp_cv** out) { return rmsd_create(desc, out); }
int cvp_gyration_create(const cvp_gyration_desc* desc, cvp_cv** out) {
  return gyration_create(desc, out);
}
int cvp_bonded_create(const cvp_bonded_desc* desc, cvp_cv** out) { return bonded_create(desc, out); }

int cvp_destroy(cvp_cv* cv) {
  delete cv;
  return CVP_OK;
}

int cvp_periodic_cutoff(const cvp_cv* cv, double* cutoff) {
  CVP_REQUIRE(cv != nullptr && cutoff != nullptr, "null argument");
  *cutoff = cv->periodic_cutoff();
  return CVP_OK;
}

int cvp_active_atoms(const cvp_cv* cv, int32_t* count, int32_t* atoms) {
  CVP_REQUIRE(cv != nullptr && count != nullptr, "null argument");
  *count = (int32_t)cv->active_atoms.size();
  if (atoms != nullptr)
    std::copy(cv->active_atoms.begin(), cv->active_atoms.end(), atoms);
  return CVP_OK;
}

int cvp_evaluate(cvp_cv* cv, int dtype, const void* positions, const void* box, int box_mode,
                 int64_t num_frames, int64_t num_atoms, void* value, void* grad, int flags,
                 void* stream) {
  int status = check_args(cv, dtype, positions, box_mode, num_frames, num_atoms, value);
  if (status != CVP_OK) return status;
  if (num_frames == 0) return CVP_OK;
  if ((status = prepare(cv)) != CVP_OK) return status;
  const bool need_grad = grad != nullptr;
  const int64_t chunk = frames_per_chunk(cv, dtype, need_grad, num_frames);
  const size_t bytes = cv->workspace_fixed(dtype) + (size_t)chunk * cv->workspace_per_frame(dtype, need_grad);
  DeviceBuffer& workspace = cv->workspace_select ? cv->workspace_alt : cv->workspace;
  if (bytes > workspace.bytes) {
    // growing frees the old allocation: make sure nothing queued earlier still uses it
    if (workspace.ptr) CVP_CUDA_CHECK(cudaDeviceSynchronize());
    if ((status = workspace.reserve(bytes)) != CVP_OK) return status;
  }
  const size_t es = elem_size(dtype);
  for (int64_t first = 0; first < num_frames; first += chunk) {
    EvalArgs a;
    a.dtype = dtype;
    a.num_frames = std::min(chunk, num_frames - first);
    a.num_atoms = num_atoms;
    a.positions = static_cast<const char*>(positions) + (size_t)first * num_atoms * 3 * es;
    a.box = box;
    if (box != nullptr && box_mode == CVP_BOX_PER_FRAME)
      a.box = static_cast<const char*>(box) + (size_t)first * 9 * es;
    a.box_mode = box == nullptr ? CVP_BOX_NONE : box_mode;
    a.value = static_cast<char*>(value) + (size_t)first * es;
    a.grad = need_grad ? static_cast<char*>(grad) + (size_t)first * num_atoms * 3 * es : nullptr;
    a.flags = flags;
    a.stream = static_cast<cudaStream_t>(stream);
    if ((status = cv->run(a, workspace.ptr)) != CVP_OK) return status;
  }
  return CVP_OK;
}

}  // extern "C"

namespace {
// Chunks of the host path: at most ~256 MiB of coordinates each, at least 1 frame, and >= 4 chunks
// when possible.  The first copy in and the last copy out overlap nothing: a short first and a short
// last chunk (a quarter of a chunk, but a whole number of 148-frame waves so that the kernels take
// the same path as for the long chunks) shorten both; the frames in between go in equal chunks.
std::vector<int64_t> host_chunks(int64_t num_frames, size_t frame_bytes) {
  std::vector<int64_t> sizes;
  if (num_frames <= 0) return sizes;
  int64_t chunk = std::max<int64_t>(1, (int64_t)((size_t(256) << 20) / std::max<size_t>(frame_bytes, 1)));
  chunk = std::min(chunk, std::max<int64_t>(1, (num_frames + 3) / 4));
  const int64_t edge = (chunk / 4 + 147) / 148 * 148;
  if (chunk >= 592 && num_frames >= 2 * edge + chunk) {
    const int64_t rest = num_frames - 2 * edge;
    const int64_t middle = (rest + chunk - 1) / chunk;
    sizes.push_back(edge);
    for (int64_t m = 0; m < middle; ++m) sizes.push_back(rest / middle + (m < rest % middle ? 1 : 0));
    sizes.push_back(edge);
  } else {
    for (int64_t first = 0; first < num_frames; first += chunk) sizes.push_back(std::min(chunk, num_frames - first));
  }
  return sizes;
}

// Host-buffer path: frames flow through H2D copy -> kernels -> D2H copy in chunks; two slots so the
// copies of one chunk overlap the kernels of the other (copy engines in both directions + SMs).
// With `masses` (host, double [N]) the effective mass of every frame is computed on the device
// from the chunk's gradient and only [B] numbers travel back; `grad` may then be NULL.
int evaluate_host_impl(cvp_cv* cv, int dtype, const void* positions, const void* box, int box_mode,
                       int64_t num_frames, int64_t num_atoms, void* value, void* grad,
                       const double* masses, void* emass, int flags, int device) {
  int status = check_args(cv, dtype, positions, box_mode, num_frames, num_atoms, value);
  if (status != CVP_OK) return status;
  CVP_REQUIRE((flags & CVP_FLAG_ACCUMULATE) == 0, "accumulate is not supported on the host path");
  CVP_REQUIRE((masses == nullptr) == (emass == nullptr), "masses and emass go together");
  if (num_frames == 0) return CVP_OK;
  DeviceGuard guard(device);
  CVP_CUDA_CHECK(guard.status);
  const size_t es = elem_size(dtype);
  const bool copy_grad = grad != nullptr;
  const bool need_grad = copy_grad || emass != nullptr;
  const size_t frame_bytes = (size_t)num_atoms * 3 * es;
  const std::vector<int64_t> sizes = host_chunks(num_frames, frame_bytes);
  const int64_t chunk = *std::max_element(sizes.begin(), sizes.end());
  // Pipeline slots (staging buffers + stream): two, so that the copies of one chunk overlap the
  // kernels of the other.  A spec that can run twice at once (concurrent_runs(): two workspaces)
  // gets four: the kernels of chunks i and i + 1 overlap -- the tail of a launch, when SMs run out
  // of CTAs, fills with the next chunk -- and chunk i + 2, whose kernels wait for the workspace of
  // chunk i, has its coordinates on the device long before.
  const bool concurrent = cv->concurrent_runs();
  const int SLOTS = concurrent ? 4 : 2;
  static_assert(HostPipeline::SLOTS >= 4, "slots");
  HostPipeline& hp = cv->host;
  const bool trace = std::getenv("CVP_HOST_TIMING") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(now() - t0).count();
  };
  const auto t_start = now();
  const bool have_box = box != nullptr && box_mode != CVP_BOX_NONE;
  if (have_box && cv->periodic_cutoff() > 0) {
    // the boxes are host data here: refuse what OpenMM refuses (half the smallest edge < cutoff)
    const int64_t boxes = box_mode == CVP_BOX_PER_FRAME ? num_frames : 1;
    for (int64_t k = 0; k < boxes; ++k) {
      double edge[3];
      for (int a = 0; a < 3; ++a)
        edge[a] = dtype == CVP_F32 ? (double)static_cast<const float*>(box)[9 * k + 4 * a]
                                   : static_cast<const double*>(box)[9 * k + 4 * a];
      CVP_REQUIRE(cv->periodic_cutoff() <= 0.5 * std::min(edge[0], std::min(edge[1], edge[2])),
                  "The cutoff distance cannot be greater than half the periodic box size.");
    }
  }
  if ((status = hp.ensure()) != CVP_OK) return status;
  if (have_box && box_mode == CVP_BOX_SHARED) {
    if ((status = hp.shared_box.reserve(9 * es)) != CVP_OK) return status;
    CVP_CUDA_CHECK(cudaMemcpy(hp.shared_box.ptr, box, 9 * es, cudaMemcpyHostToDevice));
  }
  for (int k = 0; k < SLOTS; ++k) {
    if ((status = hp.pos[k].reserve((size_t)chunk * frame_bytes)) != CVP_OK) return status;
    if (need_grad && (status = hp.grad[k].reserve((size_t)chunk * frame_bytes)) != CVP_OK)
      return status;
    if ((status = hp.value[k].reserve((size_t)chunk * es)) != CVP_OK) return status;
    if (have_box && box_mode == CVP_BOX_PER_FRAME &&
        (status = hp.box[k].reserve((size_t)chunk * 9 * es)) != CVP_OK)
      return status;
    if (emass != nullptr && (status = hp.emass[k].reserve((size_t)chunk * es)) != CVP_OK)
      return status;
  }
  if (emass != nullptr) {
    if ((status = hp.masses.reserve((size_t)num_atoms * sizeof(double))) != CVP_OK) return status;
    CVP_CUDA_CHECK(cudaMemcpy(hp.masses.ptr, masses, (size_t)num_atoms * sizeof(double),
                              cudaMemcpyHostToDevice));
  }
  // the spec's workspace is shared by both slots: kernels of consecutive chunks are serialised by
  // an event chain, copies are not -- unless the spec can run twice at once (concurrent_runs()):
  // then every slot has its own workspace and only the stream order of a slot remains
  {
    // workspaces sized for the longest chunk before anything is queued (growing one later would
    // have to wait for the device)
    if ((status = prepare(cv)) != CVP_OK) return status;
    const int64_t longest = *std::max_element(sizes.begin(), sizes.end());
    const size_t bytes = cv->workspace_fixed(dtype) +
                         (size_t)frames_per_chunk(cv, dtype, need_grad, longest) *
                             cv->workspace_per_frame(dtype, need_grad);
    for (DeviceBuffer* buffer : {&cv->workspace, &cv->workspace_alt}) {
      if (buffer == &cv->workspace_alt && !concurrent) continue;
      if (bytes > buffer->bytes) {
        if (buffer->ptr) CVP_CUDA_CHECK(cudaDeviceSynchronize());
        if ((status = buffer->reserve(bytes)) != CVP_OK) return status;
      }
    }
  }
  const double setup_ms = ms_since(t_start);
  const auto t_loop = now();
  auto drain = [&]() {
    for (int k = 0; k < SLOTS; ++k) cudaStreamSynchronize(hp.stream[k]);
  };
#define CVP_HOST_CHECK(expr)                                                             \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      set_error(std::string(#expr) + " failed: " + cudaGetErrorString(e__));             \
      drain();                                                                           \
      return CVP_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)
  int64_t index = 0;
  status = CVP_OK;
  // CVP_HOST_TIMING: device timestamps per chunk (copy in done, kernels done, copy out done)
  std::vector<cudaEvent_t> marks;
  auto mark = [&](cudaStream_t s) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    marks.push_back(e);
  };
  for (int64_t first = 0; first < num_frames && status == CVP_OK; first += sizes[index], ++index) {
    const int k = (int)(index % SLOTS);
    cudaStream_t stream = hp.stream[k];
    const int64_t count = sizes[index];
    mark(stream);
    CVP_HOST_CHECK(cudaMemcpyAsync(hp.pos[k].ptr,
                                   static_cast<const char*>(positions) + (size_t)first * frame_bytes,
                                   (size_t)count * frame_bytes, cudaMemcpyHostToDevice, stream));
    const void* dbox = hp.shared_box.ptr;
    if (have_box && box_mode == CVP_BOX_PER_FRAME) {
      CVP_HOST_CHECK(cudaMemcpyAsync(hp.box[k].ptr,
                                     static_cast<const char*>(box) + (size_t)first * 9 * es,
                                     (size_t)count * 9 * es, cudaMemcpyHostToDevice, stream));
      dbox = hp.box[k].ptr;
    }
    mark(stream);
    // the previous user of the workspace: chunk i - 1, or chunk i - 2 with two workspaces
    const int64_t back = concurrent ? 2 : 1;
    if (index >= back)
      CVP_HOST_CHECK(cudaStreamWaitEvent(stream, hp.kernels_done[(index - back) % SLOTS], 0));
    cv->workspace_select = concurrent ? (int)(index % 2) : 0;
    status = cvp_evaluate(cv, dtype, hp.pos[k].ptr, have_box ? dbox : nullptr,
                          have_box ? box_mode : CVP_BOX_NONE, count, num_atoms, hp.value[k].ptr,
                          need_grad ? hp.grad[k].ptr : nullptr, flags, stream);
    cv->workspace_select = 0;
    if (status != CVP_OK) break;
    CVP_HOST_CHECK(cudaEventRecord(hp.kernels_done[k], stream));
    mark(stream);
    CVP_HOST_CHECK(cudaMemcpyAsync(static_cast<char*>(value) + (size_t)first * es, hp.value[k].ptr,
                                   (size_t)count * es, cudaMemcpyDeviceToHost, stream));
    if (emass != nullptr) {
      status = effective_mass(dtype, hp.grad[k].ptr, static_cast<const double*>(hp.masses.ptr),
                              count, num_atoms, hp.emass[k].ptr, stream);
      if (status != CVP_OK) break;
      CVP_HOST_CHECK(cudaMemcpyAsync(static_cast<char*>(emass) + (size_t)first * es,
                                     hp.emass[k].ptr, (size_t)count * es, cudaMemcpyDeviceToHost,
                                     stream));
    }
    if (copy_grad)
      CVP_HOST_CHECK(cudaMemcpyAsync(static_cast<char*>(grad) + (size_t)first * frame_bytes,
                                     hp.grad[k].ptr, (size_t)count * frame_bytes,
                                     cudaMemcpyDeviceToHost, stream));
    mark(stream);
  }
  drain();
  if (trace && !marks.empty()) {
    for (size_t m = 0; m + 3 < marks.size(); m += 4) {
      float t[4];
      for (int q = 0; q < 4; ++q) cudaEventElapsedTime(&t[q], marks[0], marks[m + q]);
      fprintf(stderr, "[cvp_evaluate_host] chunk %zu (%lld frames): copy in %.1f-%.1f, kernels -%.1f, copy out -%.1f ms\n",
              m / 4, (long long)sizes[m / 4], t[0], t[1], t[2], t[3]);
    }
    for (cudaEvent_t e : marks) cudaEventDestroy(e);
  }
  cudaError_t last = cudaGetLastError();
  if (trace)
    fprintf(stderr,
            "[cvp_evaluate_host] device %d: setup %.1f ms, pipeline %.1f ms, %lld frames in chunks "
            "of %lld\n",
            device, setup_ms, ms_since(t_loop), (long long)num_frames, (long long)chunk);
  if (status == CVP_OK && last != cudaSuccess) {
    set_error(std::string("host evaluation failed: ") + cudaGetErrorString(last));
    return CVP_ERR_CUDA;
  }
  return status;
#undef CVP_HOST_CHECK
}
}  // namespace

extern "C" {

int cvp_host_chunks(int64_t num_frames, int64_t frame_bytes, int64_t* sizes, int32_t capacity,
                    int32_t* count) {
  CVP_REQUIRE(num_frames >= 0 && frame_bytes > 0 && count != nullptr, "invalid argument");
  const std::vector<int64_t> plan = host_chunks(num_frames, (size_t)frame_bytes);
  *count = (int32_t)plan.size();
  if (sizes != nullptr)
    for (int32_t k = 0; k < std::min<int32_t>(capacity, *count); ++k) sizes[k] = plan[k];
  return CVP_OK;
}

int cvp_evaluate_host(cvp_cv* cv, int dtype, const void* positions, const void* box, int box_mode,
                      int64_t num_frames, int64_t num_atoms, void* value, void* grad, int flags,
                      int device) {
  return evaluate_host_impl(cv, dtype, positions, box, box_mode, num_frames, num_atoms, value, grad,
                            nullptr, nullptr, flags, device);
}

int cvp_evaluate_with_mass_host(cvp_cv* cv, int dtype, const void* positions, const void* box,
                                int box_mode, int64_t num_frames, int64_t num_atoms,
                                const double* masses, void* value, void* emass, void* grad,
                                int flags, int device) {
  CVP_REQUIRE(masses != nullptr && emass != nullptr, "null masses or emass buffer");
  return evaluate_host_impl(cv, dtype, positions, box, box_mode, num_frames, num_atoms, value, grad,
                            masses, emass, flags, device);
}

// Value and effective mass on DEVICE buffers without materialising [B][N][3] for the caller: the
// gradient of a slab of frames lives in a spec-owned scratch buffer (<= 1 GiB) between the CV
// kernels and the effective-mass reduction.
int cvp_evaluate_with_mass(cvp_cv* cv, int dtype, const void* positions, const void* box,
                           int box_mode, int64_t num_frames, int64_t num_atoms,
                           const double* masses, void* value, void* emass, int flags,
                           void* stream) {
  int status = check_args(cv, dtype, positions, box_mode, num_frames, num_atoms, value);
  if (status != CVP_OK) return status;
  CVP_REQUIRE(masses != nullptr && emass != nullptr, "null masses or emass buffer");
  CVP_REQUIRE((flags & CVP_FLAG_ACCUMULATE) == 0, "accumulate is not supported with effective mass");
  if (num_frames == 0) return CVP_OK;
  const size_t es = elem_size(dtype);
  const size_t frame_bytes = (size_t)num_atoms * 3 * es;
  const int64_t slab =
      std::max<int64_t>(1, std::min<int64_t>(num_frames, (int64_t)((size_t(1) << 30) / frame_bytes)));
  if ((size_t)slab * frame_bytes > cv->grad_scratch.bytes) {
    if (cv->grad_scratch.ptr) CVP_CUDA_CHECK(cudaDeviceSynchronize());
    if ((status = cv->grad_scratch.reserve((size_t)slab * frame_bytes)) != CVP_OK) return status;
  }
  for (int64_t first = 0; first < num_frames; first += slab) {
    const int64_t count = std::min(slab, num_frames - first);
    const void* fbox = box;
    if (box != nullptr && box_mode == CVP_BOX_PER_FRAME)
      fbox = static_cast<const char*>(box) + (size_t)first * 9 * es;
    status = cvp_evaluate(cv, dtype, static_cast<const char*>(positions) + (size_t)first * frame_bytes,
                          fbox, box_mode, count, num_atoms, static_cast<char*>(value) + (size_t)first * es,
                          cv->grad_scratch.ptr, flags, stream);
    if (status != CVP_OK) return status;
    status = effective_mass(dtype, cv->grad_scratch.ptr, masses, count, num_atoms,
                            static_cast<char*>(emass) + (size_t)first * es,
                            static_cast<cudaStream_t>(stream));
    if (status != CVP_OK) return status;
  }
  return CVP_OK;
}

int cvp_effective_mass(int dtype, const void* grad, const double* masses, int64_t num_frames,
                       int64_t num_atoms, void* emass, void* stream) {
  CVP_REQUIRE(dtype == CVP_F32 || dtype == CVP_F64, "dtype must be CVP_F32 or CVP_F64");
  CVP_REQUIRE(grad != nullptr && masses != nullptr && emass != nullptr, "null argument");
  if (num_frames == 0) return CVP_OK;
  return effective_mass(dtype, grad, masses, num_frames, num_atoms, emass,
                        static_cast<cudaStream_t>(stream));
}

const char* cvp_last_error(void) { return t_error.c_str(); }
int cvp_version(void) { return 100; }

int cvp_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    cudaGetLastError();
    set_error("no CUDA device " + std::to_string(device));
    return CVP_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  CVP_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return CVP_OK;
}

int64_t cvp_launch_count(void) { return g_launch_count.load(); }


int cvp_set_workspace_limit(cvp_cv* cv, int64_t bytes) {
  CVP_REQUIRE(cv != nullptr && bytes > 0, "bad workspace limit");
  cv->workspace_limit = bytes;
  return CVP_OK;
}

int cvp_set_option(cvp_cv* cv, const char* name, int value) {
  CVP_REQUIRE(cv != nullptr && name != nullptr, "null argument");
  return cv->set_option(name, value);
}

int cvp_get_stat(cvp_cv* cv, const char* name, double* out) {
  CVP_REQUIRE(cv != nullptr && name != nullptr && out != nullptr, "null argument");
  int status = cv->get_stat(name, out);
  if (status != CVP_OK) set_error(std::string("unknown statistic: ") + name);
  return status;
}

int cvp_measure_fp64_peak(int device, double* tflops) {
  CVP_REQUIRE(tflops != nullptr, "null argument");
  DeviceGuard guard(device);
  CVP_CUDA_CHECK(guard.status);
  return measure_fp64_peak(tflops);
}

int cvp_measure_fp32_peak(int device, double* tflops) {
  CVP_REQUIRE(tflops != nullptr, "null argument");
  DeviceGuard guard(device);
  CVP_CUDA_CHECK(guard.status);
  return measure_fp32_peak(tflops);
}

}  // extern "C"
