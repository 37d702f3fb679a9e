// Gathering per-walker gradients across the GPUs of one box, B200-native.
//
// The batch is sharded by frame / walker (cvpack evaluates one context at a time, utils.py:140, so
// frames never interact): the only exchange is the all-gather of results.  north_star: "NCCL over
// NVLink used only to gather CV values and gradients".  For the values ([B] numbers) NCCL through
// the host language is fine.  For gradients ([B][N][3], 2.26 GB per CV at configs[4]) this file is
// the library's own collective over NVLink/NVSwitch peer memory:
//
//   * every rank owns a *gather buffer* [B_total][N][3] allocated here (cudaMalloc) and exported as
//     a CUDA IPC handle; the ranks exchange the 64-byte handles through the host language (any
//     out-of-band channel) and map each other's buffers (cvp_peer_buffer_open);
//   * cvp_peer_push writes the calling rank's frames straight into *every* rank's buffer with one
//     kernel: each 16-byte vector is read once from local HBM and stored `world` times, the remote
//     stores travel over NVLink as peer writes.  With `active` (the atoms that can carry a gradient,
//     cvp_active_atoms) the kernel moves only those atoms and scatters them to their place in the
//     dense buffers: the compaction, the transfer and the re-expansion of the NCCL formulation are
//     one pass, and the untouched atoms stay zero from the allocation;
//   * pushes are stream-ordered, so a chunk of walkers can be pushed on a side stream while the next
//     chunk is being evaluated -- the transfer overlaps the math chunk by chunk;
//   * cvp_peer_signal / cvp_peer_wait are a stream-ordered barrier over flags in the buffers
//     (release store at system scope after the data, acquire loads on the consumer side): no host
//     synchronisation sits between the last kernel of a step and the first consumer of the gathered
//     gradients.  The wait spins for at most ~4 s and then raises the buffer's error flag
//     (cvp_peer_status) instead of hanging the GPU.
#include <cstring>

#include "common.cuh"

struct cvp_peer_buffer {
  void* local = nullptr;
  size_t bytes = 0;       // payload bytes (the flag block follows)
  int rank = -1, world = 0;
  std::vector<void*> peers;  // [world] mapped base pointers (peers[rank] == local)
  uint32_t epoch = 0;
  int device = -1;
};

namespace cvp {
namespace {

constexpr size_t PEER_FLAG_BYTES = 1024;  // [world] epochs + error word, after the payload
constexpr int PEER_MAX_WORLD = 64;

struct PeerPointers {
  char* base[PEER_MAX_WORLD];
};

__device__ __forceinline__ uint32_t* flag_block(char* base, size_t payload) {
  return reinterpret_cast<uint32_t*>(base + payload);
}

// dense: [count][per_frame] elements of VEC bytes each, copied to every rank at `dst_offset` bytes
template <typename V>
__global__ void peer_push_dense_kernel(const V* __restrict__ src, PeerPointers peers, int world,
                                       size_t dst_offset, int64_t total) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const V v = src[t];
    for (int p = 0; p < world; ++p) reinterpret_cast<V*>(peers.base[p] + dst_offset)[t] = v;
  }
}

// compact: only the `num_active` listed atoms of every frame travel; they land at their place in
// the dense [B_total][N][3] buffers
template <typename T>
__global__ void peer_push_active_kernel(const T* __restrict__ src, PeerPointers peers, int world,
                                        int64_t first_frame, int64_t num_atoms,
                                        const int32_t* __restrict__ active, int num_active,
                                        int64_t total) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t frame = t / num_active;
    const int atom = active[(int)(t - frame * num_active)];
    const T* s = src + (frame * num_atoms + atom) * 3;
    const T x = s[0], y = s[1], z = s[2];
    const int64_t at = ((first_frame + frame) * num_atoms + atom) * 3;
    for (int p = 0; p < world; ++p) {
      T* d = reinterpret_cast<T*>(peers.base[p]) + at;
      d[0] = x;
      d[1] = y;
      d[2] = z;
    }
  }
}

// rows: bytes [row_offset, row_offset + row_bytes) of every frame (a contiguous run of atoms), as
// vectors of sizeof(V) bytes; frame b of the source lands in frame first_frame + b of every rank
template <typename V>
__global__ void peer_push_rows_kernel(const char* __restrict__ src, PeerPointers peers, int world,
                                      int64_t first_frame, size_t frame_bytes, size_t row_offset,
                                      int row_vectors, int64_t total) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t frame = t / row_vectors;
    const size_t within = row_offset + (size_t)(t - frame * row_vectors) * sizeof(V);
    const V v = *reinterpret_cast<const V*>(src + (size_t)frame * frame_bytes + within);
    const size_t at = (size_t)(first_frame + frame) * frame_bytes + within;
    for (int p = 0; p < world; ++p) *reinterpret_cast<V*>(peers.base[p] + at) = v;
  }
}

__global__ void peer_signal_kernel(PeerPointers peers, int world, int rank, size_t payload,
                                   uint32_t epoch) {
  // everything this stream wrote before is ordered before the flags, for observers on any GPU
  __threadfence_system();
  if (threadIdx.x < world) {
    uint32_t* flag = flag_block(peers.base[threadIdx.x], payload) + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
  }
}

__global__ void peer_wait_kernel(char* local, int world, size_t payload, uint32_t epoch,
                                 long long timeout_cycles) {
  uint32_t* flags = flag_block(local, payload);
  if (threadIdx.x < world) {
    const long long start = clock64();
    uint32_t seen;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + threadIdx.x) : "memory");
      // (epochs only grow; a later epoch also means this one has been signalled)
      if ((int32_t)(seen - epoch) >= 0) break;
      if (clock64() - start > timeout_cycles) {
        atomicExch(flags + PEER_MAX_WORLD, 1u);  // error word
        break;
      }
      __nanosleep(200);
    } while (true);
  }
}

int fill_pointers(const cvp_peer_buffer* buf, PeerPointers* out) {
  CVP_REQUIRE(buf->world > 0 && (int)buf->peers.size() == buf->world, "peer buffer is not open");
  for (int p = 0; p < buf->world; ++p) out->base[p] = static_cast<char*>(buf->peers[p]);
  return CVP_OK;
}

}  // namespace
}  // namespace cvp

using namespace cvp;

extern "C" {

int cvp_peer_buffer_create(int64_t bytes, cvp_peer_buffer** out) {
  CVP_REQUIRE(out != nullptr && bytes > 0, "bad peer buffer size");
  auto* buf = new cvp_peer_buffer();
  buf->bytes = ((size_t)bytes + 255) / 256 * 256;
  if (cudaGetDevice(&buf->device) != cudaSuccess ||
      cudaMalloc(&buf->local, buf->bytes + PEER_FLAG_BYTES) != cudaSuccess ||
      cudaMemset(buf->local, 0, buf->bytes + PEER_FLAG_BYTES) != cudaSuccess) {
    set_error(std::string("peer buffer allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
    if (buf->local) cudaFree(buf->local);
    delete buf;
    return CVP_ERR_CUDA;
  }
  CVP_CUDA_CHECK(cudaDeviceSynchronize());
  *out = buf;
  return CVP_OK;
}

int cvp_peer_buffer_handle(cvp_peer_buffer* buf, void* handle64) {
  CVP_REQUIRE(buf != nullptr && handle64 != nullptr, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are 64 bytes");
  cudaIpcMemHandle_t handle;
  CVP_CUDA_CHECK(cudaIpcGetMemHandle(&handle, buf->local));
  std::memcpy(handle64, &handle, sizeof(handle));
  return CVP_OK;
}

int cvp_peer_buffer_open(cvp_peer_buffer* buf, int world, int rank, const void* handles) {
  CVP_REQUIRE(buf != nullptr && handles != nullptr, "null argument");
  CVP_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, "bad rank / world size");
  CVP_REQUIRE(buf->peers.empty(), "peer buffer is already open");
  buf->peers.assign(world, nullptr);
  for (int p = 0; p < world; ++p) {
    if (p == rank) {
      buf->peers[p] = buf->local;
      continue;
    }
    cudaIpcMemHandle_t handle;
    std::memcpy(&handle, static_cast<const char*>(handles) + 64 * (size_t)p, sizeof(handle));
    cudaError_t status = cudaIpcOpenMemHandle(&buf->peers[p], handle, cudaIpcMemLazyEnablePeerAccess);
    if (status != cudaSuccess) {
      set_error("cudaIpcOpenMemHandle failed for rank " + std::to_string(p) + ": " +
                cudaGetErrorString(status));
      for (int q = 0; q < p; ++q)
        if (q != rank && buf->peers[q]) cudaIpcCloseMemHandle(buf->peers[q]);
      buf->peers.clear();
      return CVP_ERR_CUDA;
    }
  }
  buf->rank = rank;
  buf->world = world;
  return CVP_OK;
}

int cvp_peer_buffer_local(cvp_peer_buffer* buf, void** ptr, int64_t* bytes) {
  CVP_REQUIRE(buf != nullptr && ptr != nullptr, "null argument");
  *ptr = buf->local;
  if (bytes) *bytes = (int64_t)buf->bytes;
  return CVP_OK;
}

int cvp_peer_push(cvp_peer_buffer* buf, int dtype, const void* src, int64_t first_frame,
                  int64_t num_frames, int64_t num_atoms, const int32_t* active, int32_t num_active,
                  void* stream) {
  CVP_REQUIRE(buf != nullptr && src != nullptr, "null argument");
  CVP_REQUIRE(dtype == CVP_F32 || dtype == CVP_F64, "dtype must be CVP_F32 or CVP_F64");
  CVP_REQUIRE(first_frame >= 0 && num_frames >= 0 && num_atoms > 0, "bad frame range");
  const size_t es = dtype == CVP_F32 ? 4 : 8;
  const size_t frame_bytes = (size_t)num_atoms * 3 * es;
  CVP_REQUIRE((size_t)(first_frame + num_frames) * frame_bytes <= buf->bytes,
              "frames do not fit the gather buffer");
  if (num_frames == 0) return CVP_OK;
  PeerPointers peers;
  int status = fill_pointers(buf, &peers);
  if (status != CVP_OK) return status;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (active != nullptr) {
    CVP_REQUIRE(num_active > 0, "empty list of active atoms");
    const int64_t total = num_frames * (int64_t)num_active;
    const unsigned blocks = (unsigned)std::min<int64_t>(div_up(total, 256), 148 * 8);
    if (dtype == CVP_F32)
      peer_push_active_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(src), peers,
                                                            buf->world, first_frame, num_atoms, active,
                                                            num_active, total);
    else
      peer_push_active_kernel<double><<<blocks, 256, 0, st>>>(static_cast<const double*>(src), peers,
                                                             buf->world, first_frame, num_atoms, active,
                                                             num_active, total);
  } else {
    const size_t offset = (size_t)first_frame * frame_bytes;
    const size_t nbytes = (size_t)num_frames * frame_bytes;
    const bool wide = aligned16(src) && offset % 16 == 0 && nbytes % 16 == 0;
    if (wide) {
      const int64_t total = (int64_t)(nbytes / 16);
      const unsigned blocks = (unsigned)std::min<int64_t>(div_up(total, 256), 148 * 8);
      peer_push_dense_kernel<uint4><<<blocks, 256, 0, st>>>(static_cast<const uint4*>(src), peers,
                                                          buf->world, offset, total);
    } else {
      const int64_t total = (int64_t)(nbytes / 4);
      const unsigned blocks = (unsigned)std::min<int64_t>(div_up(total, 256), 148 * 8);
      peer_push_dense_kernel<uint32_t><<<blocks, 256, 0, st>>>(static_cast<const uint32_t*>(src), peers,
                                                             buf->world, offset, total);
    }
  }
  CVP_LAUNCH_CHECK();
  return CVP_OK;
}

int cvp_peer_push_rows(cvp_peer_buffer* buf, int dtype, const void* src, int64_t first_frame,
                       int64_t num_frames, int64_t num_atoms, int64_t first_atom, int64_t atom_count,
                       void* stream) {
  CVP_REQUIRE(buf != nullptr && src != nullptr, "null argument");
  CVP_REQUIRE(dtype == CVP_F32 || dtype == CVP_F64, "dtype must be CVP_F32 or CVP_F64");
  CVP_REQUIRE(first_frame >= 0 && num_frames >= 0 && num_atoms > 0 && first_atom >= 0 && atom_count >= 0 &&
                  first_atom + atom_count <= num_atoms,
              "bad frame or atom range");
  const size_t es = dtype == CVP_F32 ? 4 : 8;
  const size_t frame_bytes = (size_t)num_atoms * 3 * es;
  CVP_REQUIRE((size_t)(first_frame + num_frames) * frame_bytes <= buf->bytes,
              "frames do not fit the gather buffer");
  if (num_frames == 0 || atom_count == 0) return CVP_OK;
  PeerPointers peers;
  int status = fill_pointers(buf, &peers);
  if (status != CVP_OK) return status;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t row_offset = (size_t)first_atom * 3 * es, row_bytes = (size_t)atom_count * 3 * es;
  const bool wide = aligned16(src) && frame_bytes % 16 == 0 && row_offset % 16 == 0 && row_bytes % 16 == 0;
  const size_t vec = wide ? 16 : 4;
  const int64_t total = num_frames * (int64_t)(row_bytes / vec);
  const unsigned blocks = (unsigned)std::min<int64_t>(div_up(total, 256), 148 * 8);
  if (wide)
    peer_push_rows_kernel<uint4><<<blocks, 256, 0, st>>>(static_cast<const char*>(src), peers, buf->world,
                                                        first_frame, frame_bytes, row_offset,
                                                        (int)(row_bytes / 16), total);
  else
    peer_push_rows_kernel<uint32_t><<<blocks, 256, 0, st>>>(static_cast<const char*>(src), peers,
                                                           buf->world, first_frame, frame_bytes,
                                                           row_offset, (int)(row_bytes / 4), total);
  CVP_LAUNCH_CHECK();
  return CVP_OK;
}

int cvp_peer_signal(cvp_peer_buffer* buf, void* stream) {
  CVP_REQUIRE(buf != nullptr, "null argument");
  PeerPointers peers;
  int status = fill_pointers(buf, &peers);
  if (status != CVP_OK) return status;
  buf->epoch += 1;
  peer_signal_kernel<<<1, PEER_MAX_WORLD, 0, static_cast<cudaStream_t>(stream)>>>(
      peers, buf->world, buf->rank, buf->bytes, buf->epoch);
  CVP_LAUNCH_CHECK();
  return CVP_OK;
}

int cvp_peer_wait(cvp_peer_buffer* buf, void* stream) {
  CVP_REQUIRE(buf != nullptr && buf->world > 0, "peer buffer is not open");
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, buf->device);
  const long long timeout = 4LL * 1000LL * (long long)std::max(khz, 1000000);  // ~4 s of SM cycles
  peer_wait_kernel<<<1, PEER_MAX_WORLD, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<char*>(buf->local), buf->world, buf->bytes, buf->epoch, timeout);
  CVP_LAUNCH_CHECK();
  return CVP_OK;
}

int cvp_peer_status(cvp_peer_buffer* buf, int* timed_out) {
  CVP_REQUIRE(buf != nullptr && timed_out != nullptr, "null argument");
  uint32_t word = 0;
  CVP_CUDA_CHECK(cudaMemcpy(&word, static_cast<char*>(buf->local) + buf->bytes + 4 * PEER_MAX_WORLD, 4,
                            cudaMemcpyDeviceToHost));
  *timed_out = (int)word;
  return CVP_OK;
}

int cvp_peer_buffer_destroy(cvp_peer_buffer* buf) {
  if (buf == nullptr) return CVP_OK;
  for (int p = 0; p < (int)buf->peers.size(); ++p)
    if (p != buf->rank && buf->peers[p]) cudaIpcCloseMemHandle(buf->peers[p]);
  if (buf->local) cudaFree(buf->local);
  delete buf;
  return CVP_OK;
}

}  // extern "C"
