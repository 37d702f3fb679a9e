// One-read streaming evaluation of "reduce over the atoms of a frame, solve per frame, write a
// per-atom gradient" collective variables (RMSD, radius of gyration) on contiguous FP32 groups.
//
// Two separate launches (sums, then gradient) read every coordinate twice from HBM: 36 B/atom for
// 24 B/atom of algorithmic traffic.  This kernel is persistent and schedules the gradient of a
// frame a fixed number of *rounds* after its sums, so that the second read is served by the 126 MB
// L2 instead of HBM.
//
// Work decomposition (static, no queue):
//   * a frame's group is cut into SS sub-slices of `sub` atoms; warp (s, r) owns sub-slice s of the
//     frames f = t*RR + r, t = 0, 1, ... (RR frames are in flight per round, SS*RR warps work);
//   * iteration t of a warp: SUM(frame t*RR + r), then GRAD(frame (t-L)*RR + r): the lag is L rounds
//     = L*RR frames = L * (about 12 MB of coordinates), whatever the size of the group;
//   * every warp is autonomous: its own TMA ring (cp.async.bulk global->shared, one mbarrier per
//     stage, refilled by lane 0 as soon as the lanes have copied a chunk to registers), its own
//     resident sub-slice of the per-atom auxiliary data (centred reference / weights) in shared
//     memory, its own output staging for TMA bulk stores (or bulk reduce-add when accumulating).
//     There is no block-wide barrier anywhere, and the loads in flight do not depend on registers
//     or occupancy.
// Dependencies inside the launch:
//   * a SUM item ends with a release-add on the frame's counter.  One extra *solver warp* per CTA
//     owns the frames f = cta, cta + #CTAs, ...: it waits for the counter to reach SS, adds the
//     sub-slice partials in slice order (deterministic), runs the per-frame solve in one lane and
//     publishes the frame constants with a release flag.  (Letting the last streaming warp solve
//     puts the solve on the critical path of that warp's frame class: it is late for the next
//     frame, hence last again, and ends up solving every frame of the class serially.)
//   * GRAD(f) needs that flag.  SUM items never wait and the schedule is static, so the flag a warp
//     waits for only depends on iterations < t of other warps: no deadlock as long as all CTAs are
//     co-resident, which the cooperative launch guarantees.  The flag and the constants of the
//     GRAD item are fetched at the start of the SUM item that precedes it, so in the steady state
//     nobody waits.
#pragma once

#include <cooperative_groups.h>

#include "common.cuh"

namespace cvp {

constexpr int FUSED_STAGES = 4;         // chunks in flight per warp
constexpr int FUSED_CHUNK_ATOMS = 128;  // 4 atoms (3 x float4) per lane
constexpr int FUSED_CHUNK_BYTES = FUSED_CHUNK_ATOMS * 12;
constexpr int FUSED_MAX_WARPS = 16;
constexpr int FUSED_SMEM_LIMIT = 227 * 1024;
constexpr int FUSED_SCRATCH_ROW = 33;  // doubles per row of the warp reduction scratch (padded)

__device__ __forceinline__ void fused_unpack4(const float4& a0, const float4& a1, const float4& a2,
                                              float (&x)[4], float (&y)[4], float (&z)[4]) {
  x[0] = a0.x; y[0] = a0.y; z[0] = a0.z;
  x[1] = a0.w; y[1] = a1.x; z[1] = a1.y;
  x[2] = a1.z; y[2] = a1.w; z[2] = a2.x;
  x[3] = a2.y; y[3] = a2.z; z[3] = a2.w;
}

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}
__device__ __forceinline__ void tma_load_1d_hint(void* smem_dst, const void* gmem_src,
                                                 uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], "
      "%2, [%3], %4;" ::"r"(smem_addr(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes,
                                             uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(
                   gmem_dst),
               "r"(smem_addr(smem_src)), "r"(bytes), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_f32_1d(void* gmem_dst, const void* smem_src,
                                                      uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(
                   gmem_dst),
               "r"(smem_addr(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_commit_group() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_wait_group_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Geometry of a launch (host-chosen, see FusedLaunch::plan).
struct FusedGeom {
  int64_t num_atoms;   // atoms per frame of the batch
  int64_t num_frames;  // B
  int first_atom, n;   // the group: atoms first_atom .. first_atom + n - 1 (both multiples of 4)
  int sub;             // atoms per sub-slice (multiple of FUSED_CHUNK_ATOMS)
  int SS, RR;          // sub-slices per frame, frames per round
  int L;               // lag between the sums and the gradient of a frame, in rounds
  int T;               // rounds = ceil(B / RR)
  int warps_per_cta;
  int per_warp_bytes;  // shared memory per warp
  int ring_rounds;     // rounds of partial sums kept (ring) -- T when there is no gradient
};

// Policy:
//   static constexpr int NS                    number of double sums per frame (<= 16)
//   static constexpr int AUX                   floats of auxiliary data per atom (0, 1 or 3)
//   struct Params                              per-launch constants
//   struct Frame                               per-frame constants of the gradient; 64 bytes
//   accumulate(params, aux, x, y, z, v)        add 4 atoms to the sums; aux = this lane's 4*AUX
//                                              floats in shared memory
//   complete(params, sums, frame*, value*, accumulate)   one lane, NOT inlined: solve from the
//                                              frame's sums, fill Frame, write the value
//   gradient(params, frame, aux, x, y, z, gx, gy, gz)
template <typename Policy>
__global__ void __launch_bounds__((FUSED_MAX_WARPS + 1) * 32, 1)
    stream_fused_kernel(const float* __restrict__ pos, const float* __restrict__ aux_global,
                        FusedGeom geom, typename Policy::Params params, int* __restrict__ done,
                        int* __restrict__ ready, double* __restrict__ partials,
                        typename Policy::Frame* __restrict__ frames, float* __restrict__ value,
                        float* __restrict__ grad, int accumulate) {
  constexpr int NS = Policy::NS;
  constexpr int AUX = Policy::AUX;
  constexpr int NST = FUSED_STAGES;
  static_assert(sizeof(typename Policy::Frame) == 64, "Frame must be 64 bytes");
  static_assert(NS <= 16, "at most 16 sums");
  extern __shared__ __align__(128) unsigned char fused_smem[];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * geom.warps_per_cta + warp;
  const int SS = geom.SS, RR = geom.RR, L = geom.L, T = geom.T;
  const int64_t B = geom.num_frames;
  const bool has_grad = grad != nullptr;

  if (warp == geom.warps_per_cta) {
    // ---- solver warp: frames blockIdx.x, blockIdx.x + gridDim.x, ... ------------------------------
    for (int64_t frame = blockIdx.x; frame < B; frame += gridDim.x) {
      if (lane == 0) {
        while (ld_acquire_gpu(done + frame) < SS) __nanosleep(100);
      }
      __syncwarp();
      __threadfence();
      const int64_t slot =
          has_grad ? (int64_t)((frame / RR) % geom.ring_rounds) * RR + frame % RR : frame;
      double sums[NS];
#pragma unroll
      for (int k = 0; k < NS; ++k) sums[k] = 0.0;
      const double* in = partials + slot * SS * NS;
      for (int q = lane; q < SS; q += 32) {
#pragma unroll
        for (int k = 0; k < NS; ++k) sums[k] += __ldcg(in + (size_t)q * NS + k);
      }
#pragma unroll
      for (int k = 0; k < NS; ++k) sums[k] = warp_sum(sums[k]);
      if (lane == 0) {
        Policy::complete(params, sums, frames + frame + 1, value + frame, accumulate);
        __threadfence();
        st_release_gpu(ready + frame, 1);
      }
      __syncwarp();
    }
    return;
  }
  if (gw >= SS * RR) return;  // (no block-wide barrier anywhere)
  const int s = gw % SS, r = gw / SS;
  if (r >= B) return;

  // this warp's sub-slice
  const int a0 = s * geom.sub;
  const int sub_n = min(geom.sub, geom.n - a0);  // atoms (multiple of 4)
  const int ng = sub_n >> 2;                     // groups of 4 atoms
  const int nch = (ng + 31) >> 5;                // chunks per item

  // shared memory of this warp
  unsigned char* base = fused_smem + (size_t)warp * geom.per_warp_bytes;
  float* s_aux = reinterpret_cast<float*>(base);
  unsigned char* s_ring = base + (size_t)geom.sub * AUX * 4;
  constexpr int OUT_BYTES =
      (2 * FUSED_CHUNK_BYTES > NS * FUSED_SCRATCH_ROW * 8 ? 2 * FUSED_CHUNK_BYTES
                                                          : NS * FUSED_SCRATCH_ROW * 8 + 15) /
      16 * 16;
  unsigned char* s_out = s_ring + NST * FUSED_CHUNK_BYTES;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_out + OUT_BYTES);  // NST ring barriers + 1 (aux)

  const uint64_t pol_keep = l2_policy_evict_last();
  const uint64_t pol_stream = l2_policy_evict_first();

  if (lane == 0) {
    for (int k = 0; k <= NST; ++k) mbar_init(s_bar + k, 1);
    mbar_fence_init();
    if (AUX > 0) {
      const uint32_t bytes = (uint32_t)sub_n * AUX * 4;
      mbar_expect_tx(s_bar + NST, bytes);
      tma_load_1d_hint(s_aux, aux_global + (size_t)a0 * AUX, bytes, s_bar + NST, pol_keep);
    }
  }
  __syncwarp();

  // ---- producer cursor (uniform across the warp; lane 0 issues) ----------------------------------
  auto job_valid = [&](int t, int half) -> bool {
    if (half == 0) return t < T && (int64_t)t * RR + r < B;
    return has_grad && t >= L && (int64_t)(t - L) * RR + r < B;
  };
  const int t_end = T + (has_grad ? L : 0);
  int p_t = 0, p_half = 0, p_c = 0;  // (0, 0) is valid because r < B
  bool p_done = false;
  uint32_t prod_q = 0, cons_q = 0;
  const float* slice_base = pos + ((int64_t)geom.first_atom + a0) * 3;
  const int64_t frame_stride = geom.num_atoms * 3;

  auto produce_one = [&]() {
    if (p_done) return;
    const int64_t frame = (int64_t)(p_half == 0 ? p_t : p_t - L) * RR + r;
    const int first_group = p_c << 5;
    const uint32_t bytes = (uint32_t)min(32, ng - first_group) * 48u;
    const int stage = prod_q % NST;
    if (lane == 0) {
      mbar_expect_tx(s_bar + stage, bytes);
      tma_load_1d_hint(s_ring + stage * FUSED_CHUNK_BYTES,
                       slice_base + frame * frame_stride + (int64_t)first_group * 12, bytes,
                       s_bar + stage, (p_half == 0 && has_grad) ? pol_keep : pol_stream);
    }
    ++prod_q;
    if (++p_c == nch) {
      p_c = 0;
      for (;;) {
        if (p_half == 0) {
          p_half = 1;
        } else {
          p_half = 0;
          ++p_t;
        }
        if (p_t >= t_end) {
          p_done = true;
          break;
        }
        if (job_valid(p_t, p_half)) break;
      }
    }
  };
  for (int k = 0; k < NST; ++k) produce_one();
  if (AUX > 0) mbar_wait(s_bar + NST, 0);

  // frame constants of the next GRAD item, fetched early
  int pre_flag = 0;
  float4 pre[4];
  auto prefetch_frame = [&](int64_t frame) {
    pre_flag = ld_relaxed_gpu(ready + frame);
    // slot 0 is a dummy: the address depends on the flag, so the loads cannot pass it
    const float4* src = reinterpret_cast<const float4*>(frames + (pre_flag != 0 ? frame + 1 : 0));
#pragma unroll
    for (int k = 0; k < 4; ++k) pre[k] = __ldcg(src + k);
  };

  int out_buf = 0;
  for (int t = 0; t < t_end; ++t) {
    const bool grad_valid = job_valid(t, 1);
    const int64_t grad_frame = (int64_t)(t - L) * RR + r;
    if (grad_valid) prefetch_frame(grad_frame);

    if (job_valid(t, 0)) {
      // ---- SUM -----------------------------------------------------------------------------------
      const int64_t frame = (int64_t)t * RR + r;
      double v[NS];
#pragma unroll
      for (int k = 0; k < NS; ++k) v[k] = 0.0;
      for (int c = 0; c < nch; ++c) {
        const int stage = cons_q % NST;
        mbar_wait(s_bar + stage, (cons_q / NST) & 1);
        const int g = (c << 5) + lane;
        const bool active = g < ng;
        const float4* px =
            reinterpret_cast<const float4*>(s_ring + stage * FUSED_CHUNK_BYTES) + 3 * lane;
        float4 x0, x1, x2;
        if (active) {
          x0 = px[0];
          x1 = px[1];
          x2 = px[2];
        }
        __syncwarp();
        ++cons_q;
        produce_one();
        if (active) {
          float x[4], y[4], z[4];
          fused_unpack4(x0, x1, x2, x, y, z);
          Policy::accumulate(params, s_aux + (size_t)g * 4 * AUX, x, y, z, v);
        }
      }
      // warp reduction through shared memory (the output staging is idle: drain its readers)
      if (lane == 0) tma_wait_group_read<0>();
      __syncwarp();
      double* scratch = reinterpret_cast<double*>(s_out);
#pragma unroll
      for (int k = 0; k < NS; ++k) scratch[k * FUSED_SCRATCH_ROW + lane] = v[k];
      __syncwarp();
      {
        const int k2 = lane >> 1, h = lane & 1;
        double acc = 0.0;
        if (k2 < NS) {
          const double* row = scratch + k2 * FUSED_SCRATCH_ROW + h * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i) acc += row[i];
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        const int64_t slot = has_grad ? (int64_t)(t % geom.ring_rounds) * RR + r : frame;
        if (k2 < NS && h == 0) __stcg(partials + (slot * SS + s) * NS + k2, acc);
      }
      __syncwarp();
      if (lane == 0) red_release_add(done + frame, 1);
    }

    if (grad_valid) {
      // ---- GRAD ----------------------------------------------------------------------------------
      if (!__all_sync(0xffffffffu, pre_flag != 0)) {
        if (lane == 0) {
          while (ld_acquire_gpu(ready + grad_frame) == 0) __nanosleep(200);
        }
        __syncwarp();
        __threadfence();
        const float4* src = reinterpret_cast<const float4*>(frames + grad_frame + 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) pre[k] = __ldcg(src + k);
      }
      typename Policy::Frame fr;
      {
        float4* dst = reinterpret_cast<float4*>(&fr);
#pragma unroll
        for (int k = 0; k < 4; ++k) dst[k] = pre[k];
      }
      float* gbase = grad + grad_frame * frame_stride + ((int64_t)geom.first_atom + a0) * 3;
      for (int c = 0; c < nch; ++c) {
        const int stage = cons_q % NST;
        mbar_wait(s_bar + stage, (cons_q / NST) & 1);
        const int g = (c << 5) + lane;
        const bool active = g < ng;
        const float4* px =
            reinterpret_cast<const float4*>(s_ring + stage * FUSED_CHUNK_BYTES) + 3 * lane;
        float4 x0, x1, x2;
        if (active) {
          x0 = px[0];
          x1 = px[1];
          x2 = px[2];
        }
        // the staging buffer written two chunks ago must have been read by its bulk store
        if (lane == 0) tma_wait_group_read<1>();
        __syncwarp();
        ++cons_q;
        produce_one();
        float4* po = reinterpret_cast<float4*>(s_out + out_buf * FUSED_CHUNK_BYTES) + 3 * lane;
        if (active) {
          float x[4], y[4], z[4], gx[4], gy[4], gz[4];
          fused_unpack4(x0, x1, x2, x, y, z);
          Policy::gradient(params, fr, s_aux + (size_t)g * 4 * AUX, x, y, z, gx, gy, gz);
          po[0] = make_float4(gx[0], gy[0], gz[0], gx[1]);
          po[1] = make_float4(gy[1], gz[1], gx[2], gy[2]);
          po[2] = make_float4(gz[2], gx[3], gy[3], gz[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int first_group = c << 5;
          const uint32_t bytes = (uint32_t)min(32, ng - first_group) * 48u;
          float* dst = gbase + (int64_t)first_group * 12;
          if (accumulate)
            tma_reduce_add_f32_1d(dst, s_out + out_buf * FUSED_CHUNK_BYTES, bytes);
          else
            tma_store_1d(dst, s_out + out_buf * FUSED_CHUNK_BYTES, bytes, pol_stream);
          tma_commit_group();
        }
        out_buf ^= 1;
      }
    }
  }
  if (lane == 0) tma_wait_group_all();
}

// Host helper: geometry, workspace layout and launch.
template <typename Policy>
struct FusedLaunch {
  static constexpr int out_bytes() {
    return (2 * FUSED_CHUNK_BYTES > Policy::NS * FUSED_SCRATCH_ROW * 8
                ? 2 * FUSED_CHUNK_BYTES
                : Policy::NS * FUSED_SCRATCH_ROW * 8 + 15) /
           16 * 16;
  }
  static int per_warp_bytes(int sub) {
    const int bytes = sub * Policy::AUX * 4 + FUSED_STAGES * FUSED_CHUNK_BYTES + out_bytes() +
                      (FUSED_STAGES + 1) * 8;
    return (bytes + 127) / 128 * 128;
  }
  // workspace: done[B], ready[B], frames[B + 1], partials[B * SS_max * NS]; SS_max for sub = 128
  static size_t workspace_per_frame(int n) {
    const size_t ss_max = (size_t)div_up(n, FUSED_CHUNK_ATOMS);
    return ss_max * Policy::NS * sizeof(double) + sizeof(typename Policy::Frame) + 2 * sizeof(int) +
           16;
  }
  static size_t workspace_fixed() { return 8 * 256 + sizeof(typename Policy::Frame); }

  // Chooses the sub-slice size: fills the resident warps (SS * RR of SMs * warps_per_cta) with as
  // little rounding loss as possible; larger sub-slices (fewer reductions) win ties.  Returns false
  // when the problem does not fit the scheme (group too small or too large, too few frames).
  static bool plan(int n, int64_t num_atoms, int first_atom, int64_t num_frames, bool has_grad,
                   int lag_rounds, int sub_option, int sms, FusedGeom* geom) {
    if (n < 4 * FUSED_CHUNK_ATOMS || n % 4 != 0 || first_atom % 4 != 0 || num_atoms % 4 != 0)
      return false;
    double best = 0.0;
    FusedGeom g{};
    for (int sub = FUSED_CHUNK_ATOMS; sub <= 8 * FUSED_CHUNK_ATOMS; sub += FUSED_CHUNK_ATOMS) {
      if (sub_option > 0 && sub != sub_option) continue;
      const int pwb = per_warp_bytes(sub);
      const int wpc = std::min(FUSED_MAX_WARPS, (FUSED_SMEM_LIMIT - 1024) / pwb);  // + the solver warp
      if (wpc < 1) continue;
      const int64_t W = (int64_t)sms * wpc;
      const int64_t SS = div_up(n, sub);
      if (SS > W) continue;
      const int64_t RR = std::min<int64_t>(W / SS, num_frames);
      // fraction of the machine's warp slots doing useful work, discounted for the per-item
      // reduction (about 60 instructions against 25 per atom and lane)
      const double fill = (double)(SS * RR) / (double)W * (double)n / (double)(SS * sub);
      const double overhead = 1.0 / (1.0 + 60.0 / (25.0 * sub / 32.0));
      const double score = fill * overhead * (double)wpc / FUSED_MAX_WARPS;
      if (score > best) {
        best = score;
        g.sub = sub;
        g.SS = (int)SS;
        g.RR = (int)RR;
        g.warps_per_cta = wpc;
        g.per_warp_bytes = pwb;
      }
    }
    if (best == 0.0) return false;
    g.num_atoms = num_atoms;
    g.num_frames = num_frames;
    g.first_atom = first_atom;
    g.n = n;
    g.T = (int)div_up(num_frames, g.RR);
    g.L = std::max(1, std::min(lag_rounds, g.T));
    g.ring_rounds = has_grad ? std::min(g.T, g.L + 2) : g.T;
    *geom = g;
    return true;
  }

  static int launch(const float* pos, const float* aux, const FusedGeom& geom, int sms,
                    const typename Policy::Params& params, void* workspace, float* value,
                    float* grad, bool accumulate, cudaStream_t stream) {
    using Frame = typename Policy::Frame;
    const size_t B = (size_t)geom.num_frames;
    Carver carve(workspace);
    int* flags = carve.take<int>(2 * B);
    Frame* frames = carve.take<Frame>(B + 1);
    double* partials = carve.take<double>((size_t)geom.ring_rounds * geom.RR * geom.SS * Policy::NS);
    CVP_CUDA_CHECK(cudaMemsetAsync(flags, 0, 2 * B * sizeof(int), stream));
    int* done = flags;
    int* ready = flags + B;
    const size_t smem = (size_t)geom.warps_per_cta * geom.per_warp_bytes;
    static bool configured = false;
    if (!configured) {
      CVP_CUDA_CHECK(cudaFuncSetAttribute(stream_fused_kernel<Policy>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          FUSED_SMEM_LIMIT));
      configured = true;
    }
    const int blocks = (int)std::min<int64_t>(sms, div_up((int64_t)geom.SS * geom.RR, geom.warps_per_cta));
    FusedGeom g = geom;
    typename Policy::Params p = params;
    int acc = accumulate ? 1 : 0;
    void* args[] = {(void*)&pos, (void*)&aux, (void*)&g,      (void*)&p,     (void*)&done, (void*)&ready,
                    (void*)&partials, (void*)&frames, (void*)&value, (void*)&grad, (void*)&acc};
    // cooperative: all CTAs are guaranteed to be co-resident (GRAD items wait for other CTAs)
    CVP_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)stream_fused_kernel<Policy>, dim3(blocks),
                                               dim3((geom.warps_per_cta + 1) * 32), args, smem, stream));
    return CVP_OK;
  }
};

}  // namespace cvp
