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
//     stage, refilled by lane 0 as soon as the lanes have copied a chunk to registers) and its own
//     output staging for TMA bulk stores (or bulk reduce-add when accumulating).  The sub-slice of
//     the per-atom auxiliary data (centred reference / weights) stays resident in shared memory,
//     one copy per *slot* shared by the warps of the CTA that work on the same sub-slice (for
//     different frames).  There is no block-wide barrier after the set-up, and the loads in flight
//     do not depend on registers or occupancy.
// Dependencies inside the launch:
//   * a SUM item ends by writing its partial sums as tagged records: every double travels as
//     {lo, tag, hi, tag} (one 16-byte store; each 8-byte half carries the round's tag, the way
//     NCCL's LL protocol does), so the streaming warps need neither a fence nor a counter.  One
//     extra *solver warp* per CTA owns the frames f = cta, cta + #CTAs, ...: it reads the records
//     of the frame's SS sub-slices, re-reading those whose tags have not arrived, adds them in
//     slice order (deterministic), runs the per-frame solve in one lane and publishes the frame
//     constants the same way, as tagged records {c0, tag, c1, tag}: no fence, no flag.  (Letting the last streaming warp solve
//     puts the solve on the critical path of that warp's frame class: it is late for the next
//     frame, hence last again, and ends up solving every frame of the class serially.)
//   * GRAD(f) needs those constants.  SUM items never wait and the schedule is static, so the
//     records a warp waits for only depend on iterations < t of other warps: no deadlock as long as
//     all CTAs are co-resident, which the cooperative launch guarantees.  The constants of the GRAD
//     item are fetched at the start of the SUM item that precedes it (one L2 round trip, checked
//     when needed), so in the steady state nobody waits.
#pragma once

#include <cooperative_groups.h>

#include "common.cuh"

namespace cvp {

#ifndef FUSED_STAGES_VALUE
#define FUSED_STAGES_VALUE 2
#endif
#ifndef FUSED_OUT_BUFS_VALUE
#define FUSED_OUT_BUFS_VALUE 1
#endif
constexpr int FUSED_STAGES = FUSED_STAGES_VALUE;  // ring stages per warp (refilled right after the LDS)
constexpr int FUSED_OUT_BUFS = FUSED_OUT_BUFS_VALUE;  // output staging buffers per warp (1 or 2)
constexpr int FUSED_CHUNK_ATOMS = 256;  // 2 x 4 atoms (2 x 3 float4) per lane
constexpr int FUSED_CHUNK_BYTES = FUSED_CHUNK_ATOMS * 12;
constexpr int FUSED_MAX_WARPS = 15;  // streaming warps; with the solver warp 16 warps = 128 registers
constexpr int FUSED_SMEM_LIMIT = 227 * 1024;
constexpr int FUSED_SCRATCH_ROW = 33;  // doubles per row of the warp reduction scratch (padded)
constexpr int FUSED_GATHER_BYTES = 24 * 1024;  // solver's staging of one segment of partial records
constexpr int FUSED_GATHER_UNROLL = 4;         // probes in flight per lane: <= 128 slices per segment

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
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(policy));
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
__device__ __forceinline__ uint4 ld_relaxed_gpu_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
// L2-only (cache-global) load for polling tagged records: never served by a stale L1 line, and
// unlike the .relaxed.gpu form it pipelines like an ordinary load
__device__ __forceinline__ uint4 ld_cg_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu_v4(uint4* p, const uint4& v) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Layout of the double-precision auxiliary data (Policy::AUXD == 3): blocks of 128 atoms = 32
// groups of 4 atoms = 6 double2 per group; piece j of group l of a block sits at double2 index
// j * 32 + l, so that a warp reads one piece with a conflict-free LDS.128.  The array is padded
// with zeros to a whole number of blocks.
template <typename V>
inline std::vector<double> fused_permute_doubles(const V* values, size_t atoms) {
  const size_t blocks = (atoms + 127) / 128;
  std::vector<double> out(blocks * 128 * 3, 0.0);
  for (size_t a = 0; a < atoms; ++a)
    for (int d = 0; d < 3; ++d) {
      const size_t block = a / 128, group = (a % 128) / 4, k = (a % 4) * 3 + d;  // k in 0..11
      out[block * 384 + ((k / 2) * 32 + group) * 2 + (k % 2)] = (double)values[a * 3 + d];
    }
  return out;
}

#ifdef FUSED_TIMING
// (diagnostics build) per-frame timestamps in nanoseconds: [frame][6] =
// first record written, last record written, records gathered, constants published,
// first consumer needs the constants, last consumer has them
static __device__ unsigned long long* g_fused_timing = nullptr;  // (one copy per translation unit)
__device__ __forceinline__ unsigned long long fused_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif

// Geometry of a launch (host-chosen, see FusedLaunch::plan).
struct FusedGeom {
  int64_t num_atoms;   // atoms per frame of the batch
  int64_t num_frames;  // B
  int first_atom, n;   // the group: atoms first_atom .. first_atom + n - 1 (both multiples of 4)
  int sub;             // atoms per sub-slice (multiple of FUSED_CHUNK_ATOMS)
  int SS, RR;          // sub-slices per frame, frames per round
  int L;               // lag between the sums and the gradient of a frame, in rounds
  int T;               // rounds = ceil(B / RR)
  int warps_per_cta;   // streaming warps (one more warp per CTA solves)
  int slots;           // auxiliary-data slots per CTA; warps_per_cta / slots warps share one
  int K;               // frame classes per slot column: slot g -> sub-slice g % SS, class g / SS < K
  int ring_rounds;     // rounds of partial sums kept (ring) -- T when there is no gradient
  int l2_hints;        // bit 0: evict_last for the first read, bit 1: evict_first for second read
                       // and stores
  int gather_slices;   // sub-slices per segment of the solver's gather (<= 32 * FUSED_GATHER_UNROLL)
};

// Policy:
//   static constexpr int NS                    number of double sums per frame (<= 16)
//   static constexpr int AUX                   floats of auxiliary data per atom (0, 1 or 3)
//   static constexpr int AUXD                  doubles of auxiliary data per atom for the sums (0 or
//                                              3): the same values, pre-converted, in the
//                                              conflict-free layout of fused_permute_doubles()
//   struct Params                              per-launch constants
//   struct Frame                               per-frame constants of the gradient; 64 bytes, of
//   static constexpr int FRAME_FLOATS          which the first FRAME_FLOATS floats are used
//   accumulate(params, aux, auxd, x, y, z, v)  add 4 atoms to the sums; aux = this lane's 4*AUX
//                                              floats in shared memory, auxd = its first double2
//                                              (the others follow at strides of 32 double2)
//   complete(params, sums, frame*, value*, accumulate)   one lane, NOT inlined: solve from the
//                                              frame's sums, fill Frame, write the value
//   gradient(params, frame, aux, x, y, z, gx, gy, gz)
template <typename Policy>
struct FusedLayout {
  static constexpr int NS = Policy::NS;
  static constexpr int SCRATCH_BYTES = (NS * FUSED_SCRATCH_ROW * 8 + 15) / 16 * 16;
  static constexpr int OUT_BYTES = FUSED_OUT_BUFS * FUSED_CHUNK_BYTES > SCRATCH_BYTES
                                       ? FUSED_OUT_BUFS * FUSED_CHUNK_BYTES
                                       : SCRATCH_BYTES;
  static constexpr int RING_BYTES = FUSED_STAGES * FUSED_CHUNK_BYTES;
  static constexpr int WARP_BYTES = (RING_BYTES + OUT_BYTES + FUSED_STAGES * 8 + 127) / 128 * 128;
  static int slot_bytes(int sub) {
    return (sub * (Policy::AUX * 4 + Policy::AUXD * 8) + 127) / 128 * 128;
  }
};

template <typename Policy>
__global__ void __launch_bounds__((FUSED_MAX_WARPS + 1) * 32, 1)
    stream_fused_kernel(const float* __restrict__ pos, const float* __restrict__ aux_global,
                        const double* __restrict__ auxd_global, FusedGeom geom, typename Policy::Params params,
                        uint4* __restrict__ partials, uint4* __restrict__ frames /* tagged */,
                        float* __restrict__ value,
                        float* __restrict__ grad, int accumulate) {
  using Layout = FusedLayout<Policy>;
  constexpr int NS = Policy::NS;
  constexpr int AUX = Policy::AUX;
  constexpr int AUXD = Policy::AUXD;
  constexpr int CH = FUSED_CHUNK_BYTES;
  static_assert(FUSED_STAGES >= 2 && FUSED_OUT_BUFS >= 1 && FUSED_OUT_BUFS <= 2, "ring geometry");
  static_assert(sizeof(typename Policy::Frame) == 64, "Frame must be 64 bytes");
  // frame constants travel as FREC records {c[2k], tag, c[2k+1], tag}
  constexpr int FREC = (Policy::FRAME_FLOATS + 1) / 2;
  static_assert(NS <= 16, "at most 16 sums");
  extern __shared__ __align__(128) unsigned char fused_smem[];
  __shared__ uint64_t s_aux_bar[FUSED_MAX_WARPS];
  __shared__ uint64_t s_gather_bar;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // (warp-uniform for the compiler)
  const int lane = threadIdx.x & 31;
  const int SS = geom.SS, RR = geom.RR, L = geom.L, T = geom.T;
  const int64_t B = geom.num_frames;
  const bool has_grad = grad != nullptr;
  const bool use_ring = geom.ring_rounds < T;  // partial records: ring of rounds, else one per frame
  const int wpc = geom.warps_per_cta;
  const int wps = wpc / geom.slots;  // warps per slot
  const int slot_bytes = (geom.sub * (AUX * 4 + AUXD * 8) + 127) / 128 * 128;

  // ---- set-up: barriers, auxiliary data of this CTA's slots -------------------------------------
  const uint64_t pol_keep = (geom.l2_hints & 1) ? l2_policy_evict_last() : l2_policy_evict_normal();
  const uint64_t pol_stream =
      (geom.l2_hints & 2) ? l2_policy_evict_first() : l2_policy_evict_normal();
  unsigned char* warp_base = fused_smem + (size_t)geom.slots * slot_bytes + (size_t)warp * Layout::WARP_BYTES;
  unsigned char* s_ring = warp_base;
  unsigned char* s_out = warp_base + Layout::RING_BYTES;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_out + Layout::OUT_BYTES);
  if (warp < wpc && lane == 0) {
    for (int k = 0; k < FUSED_STAGES; ++k) mbar_init(s_bar + k, 1);
    if (warp < geom.slots) mbar_init(s_aux_bar + warp, 1);
    mbar_fence_init();
  }
  if (warp == wpc && lane == 0) {
    mbar_init(&s_gather_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();  // the only block-wide barrier: the aux barriers are shared between warps

  if (warp == wpc) {
    // ---- solver warp: frames blockIdx.x, blockIdx.x + gridDim.x, ... ------------------------------
    // Solvers are dedicated to frame classes (solver c serves class c % RR, every M-th round of
    // it): within a class frames complete in order, so a solver never sits on a frame of a slow
    // class while later frames of faster classes are ready (measured: 20-50 us of head-of-line
    // blocking with a round-robin assignment).
    const int M = max(1, (int)gridDim.x / RR);  // solvers per class
    const int solver_class = blockIdx.x % RR, solver_phase = blockIdx.x / RR;
    // with fewer CTAs than classes a solver takes several classes, round by round
    const int class_stride = gridDim.x < (unsigned)RR ? (int)gridDim.x : RR;
    if (gridDim.x >= (unsigned)RR && solver_phase >= M) return;
    uint4* s_gather = reinterpret_cast<uint4*>(
        fused_smem + (size_t)geom.slots * slot_bytes + (size_t)wpc * Layout::WARP_BYTES);
    uint32_t gather_parity = 0;
    for (int round = gridDim.x >= (unsigned)RR ? solver_phase : 0; round < T;
         round += gridDim.x >= (unsigned)RR ? M : 1)
    for (int cls = solver_class; cls < RR; cls += class_stride) {
      const int64_t frame = (int64_t)round * RR + cls;
      if (frame >= B) continue;
      const uint32_t tag = (uint32_t)round + 1u;
      const int64_t slot = use_ring ? (int64_t)(round % geom.ring_rounds) * RR + frame % RR : frame;
      double sums[NS];
#pragma unroll
      for (int k = 0; k < NS; ++k) sums[k] = 0.0;
      // Gather in segments of GS sub-slices, two L2 round trips per segment once the records are
      // there: (1) probe the last record of every sub-slice of the segment (<= FUSED_GATHER_UNROLL
      // independent 16-byte loads per lane) until all carry the round's tag; (2) one bulk copy of
      // the segment's records to shared memory, every tag checked there.  (Reading the records
      // into registers, a batch of NS per lane and 32 sub-slices at a time, was SS / 32 dependent
      // round trips: 12 us of the 10 us a round lasts.)
      const int GS = geom.gather_slices;
      for (int seg0 = 0; seg0 < SS; seg0 += GS) {
        const int nseg = min(GS, SS - seg0);
        const uint4* in = partials + (slot * SS + seg0) * NS;
        for (;;) {
          uint4 probe[FUSED_GATHER_UNROLL];
#pragma unroll
          for (int i = 0; i < FUSED_GATHER_UNROLL; ++i) {
            const int q = lane + 32 * i;
            probe[i] = q < nseg ? ld_cg_v4(in + (size_t)q * NS + (NS - 1)) : make_uint4(0u, tag, 0u, tag);
          }
          bool there = true;
#pragma unroll
          for (int i = 0; i < FUSED_GATHER_UNROLL; ++i)
            there = there && probe[i].y == tag && probe[i].w == tag;
          if (__all_sync(0xffffffffu, there)) break;
          __nanosleep(100);
        }
        const uint32_t bytes = (uint32_t)nseg * NS * 16u;
        for (;;) {
          if (lane == 0) {
            mbar_expect_tx(&s_gather_bar, bytes);
            tma_load_1d_hint(s_gather, in, bytes, &s_gather_bar, pol_stream);
          }
          mbar_wait(&s_gather_bar, gather_parity);
          gather_parity ^= 1u;
          double seg_sums[NS];
#pragma unroll
          for (int k = 0; k < NS; ++k) seg_sums[k] = sums[k];
          bool valid = true;
          for (int q = lane; q < nseg; q += 32) {
            const uint4* rec = s_gather + q * NS;
#pragma unroll
            for (int k = 0; k < NS; ++k) {
              const uint4 v = rec[k];
              valid = valid && v.y == tag && v.w == tag;
              seg_sums[k] += __hiloint2double((int)v.z, (int)v.x);
            }
          }
          // (all lanes have read the staging before lane 0 may refill it)
          if (__all_sync(0xffffffffu, valid)) {
#pragma unroll
            for (int k = 0; k < NS; ++k) sums[k] = seg_sums[k];
            break;
          }
          __nanosleep(100);
        }
      }
#pragma unroll
      for (int k = 0; k < NS; ++k) sums[k] = warp_sum(sums[k]);
#ifdef FUSED_TIMING
      if (lane == 0 && g_fused_timing != nullptr) g_fused_timing[frame * 6 + 2] = fused_now();
#endif
      if (lane == 0) {
        typename Policy::Frame fr;
        Policy::complete(params, sums, &fr, value + frame, accumulate);
        const uint32_t* words = reinterpret_cast<const uint32_t*>(&fr);
#pragma unroll
        for (int k = 0; k < FREC; ++k)
          st_relaxed_gpu_v4(frames + frame * FREC + k, make_uint4(words[2 * k], 1u, words[2 * k + 1], 1u));
#ifdef FUSED_TIMING
        if (g_fused_timing != nullptr) g_fused_timing[frame * 6 + 3] = fused_now();
#endif
      }
      __syncwarp();
    }
    return;
  }

  // ---- streaming warp (s, r) ----------------------------------------------------------------------
  const int my_slot = warp / wps;
  const int gslot = blockIdx.x * geom.slots + my_slot;
  const bool working = gslot < SS * geom.K;
  const int s = gslot % SS;
  const int r = (gslot / SS) * wps + (warp - my_slot * wps);
  const int a0 = s * geom.sub;
  const int sub_n = min(geom.sub, geom.n - a0);  // atoms of this sub-slice (multiple of 4)
  // slot: [doubles, blocks of 128 atoms in the permuted layout][floats]
  double* s_auxd = reinterpret_cast<double*>(fused_smem + (size_t)my_slot * slot_bytes);
  float* s_aux = reinterpret_cast<float*>(s_auxd + (size_t)geom.sub * AUXD);
  if (AUX + AUXD > 0 && working && warp == my_slot * wps && lane == 0) {
    const uint32_t bytes = (uint32_t)sub_n * AUX * 4;
    const uint32_t bytes_d = (uint32_t)((sub_n + 127) / 128 * 128) * AUXD * 8;
    mbar_expect_tx(s_aux_bar + my_slot, bytes + bytes_d);
    if (AUX > 0)
      tma_load_1d_hint(s_aux, aux_global + (size_t)a0 * AUX, bytes, s_aux_bar + my_slot,
                       l2_policy_evict_last());
    if (AUXD > 0)
      tma_load_1d_hint(s_auxd, auxd_global + (size_t)a0 * AUXD, bytes_d, s_aux_bar + my_slot,
                       l2_policy_evict_last());
  }
  if (!working || r >= B) return;

  const int ng = sub_n >> 2;         // groups of 4 atoms
  const int nch = (ng + 63) >> 6;    // chunks per item
  const uint32_t last_bytes = (uint32_t)(ng - ((nch - 1) << 6)) * 48u;
  const int64_t frame_stride = geom.num_atoms * 3;  // floats
  const float* slice_base = pos + ((int64_t)geom.first_atom + a0) * 3;

  // producer (lane 0 keeps the state): chunks of the item sequence SUM(0) [GRAD(0-L)] SUM(1) ...
  auto job_valid = [&](int t, int half) -> bool {
    if (half == 0) return t < T && (int64_t)t * RR + r < B;
    return has_grad && t >= L && (int64_t)(t - L) * RR + r < B;
  };
  const int t_end = T + (has_grad ? L : 0);
  int p_t = 0, p_half = 0, p_left = nch, p_stage = 0;
  bool p_done = false;
  const float* p_ptr = slice_base + (int64_t)r * frame_stride;
  auto produce = [&]() {  // all lanes keep the (uniform) state, one lane issues
    if (p_done) return;
    const uint32_t bytes = p_left == 1 ? last_bytes : (uint32_t)CH;
    if (elect_one()) {
      mbar_expect_tx(s_bar + p_stage, bytes);
      tma_load_1d_hint(s_ring + p_stage * CH, p_ptr, bytes, s_bar + p_stage,
                       p_half == 0 && has_grad ? pol_keep : pol_stream);
    }
    p_stage = p_stage + 1 == FUSED_STAGES ? 0 : p_stage + 1;
    p_ptr += CH / 4;
    if (--p_left == 0) {
      for (;;) {
        p_t += p_half;
        p_half ^= 1;
        if (p_t >= t_end) {
          p_done = true;
          return;
        }
        if (job_valid(p_t, p_half)) break;
      }
      p_left = nch;
      p_ptr = slice_base + ((int64_t)(p_half == 0 ? p_t : p_t - L) * RR + r) * frame_stride;
    }
  };
  for (int k = 0; k < FUSED_STAGES; ++k) produce();
  if (AUX + AUXD > 0) mbar_wait(s_aux_bar + my_slot, 0);

  int c_stage = 0;         // ring stage of the next chunk to consume ...
  uint32_t c_parity = 0;   // ... and the parity its barrier completes with
  int out_buf = 0;
  const float4* lane_ring = reinterpret_cast<const float4*>(s_ring) + 3 * lane;
  const float* lane_aux = s_aux + lane * 4 * AUX;
  const double2* lane_auxd = reinterpret_cast<const double2*>(s_auxd) + lane;

  // frame constants of the next GRAD item, fetched early (valid once every record carries tag 1)
  uint4 pre[FREC];

  for (int t = 0; t < t_end; ++t) {
    const bool grad_valid = job_valid(t, 1);
    const int64_t grad_frame = (int64_t)(t - L) * RR + r;
    if (grad_valid) {
      // (with a lag of one round the frame's sums have only just been written: fetch when needed)
#pragma unroll
      for (int k = 0; k < FREC; ++k)
        pre[k] = L > 1 ? ld_relaxed_gpu_v4(frames + grad_frame * FREC + k) : make_uint4(0u, 0u, 0u, 0u);
    }

    if (job_valid(t, 0)) {
      // ---- SUM -----------------------------------------------------------------------------------
      const int64_t frame = (int64_t)t * RR + r;
      double v[NS];
#pragma unroll
      for (int k = 0; k < NS; ++k) v[k] = 0.0;
      int left = ng;
      const float* aux = lane_aux;
      const double2* auxd = lane_auxd;
      for (int c = 0; c < nch; ++c, left -= 64, aux += 64 * 4 * AUX, auxd += 64 * AUXD * 2) {
        const int stage = c_stage;
        mbar_wait(s_bar + stage, c_parity);
        if (++c_stage == FUSED_STAGES) {
          c_stage = 0;
          c_parity ^= 1u;
        }
        const float4* px = lane_ring + stage * (CH / 16);
        if (lane < left) {
          float x[4], y[4], z[4];
          fused_unpack4(px[0], px[1], px[2], x, y, z);
          Policy::accumulate(params, aux, auxd, x, y, z, v);
        }
        if (lane + 32 < left) {
          float x[4], y[4], z[4];
          fused_unpack4(px[96], px[97], px[98], x, y, z);
          Policy::accumulate(params, aux + 32 * 4 * AUX, auxd + 32 * AUXD * 2, x, y, z, v);
        }
        __syncwarp();
        produce();
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
        const int64_t slot = use_ring ? (int64_t)(t % geom.ring_rounds) * RR + r : frame;
        if (k2 < NS && h == 0) {
          const uint32_t tag = (uint32_t)t + 1u;
          st_relaxed_gpu_v4(partials + (slot * SS + s) * NS + k2,
                            make_uint4((uint32_t)__double2loint(acc), tag,
                                       (uint32_t)__double2hiint(acc), tag));
        }
      }
#ifdef FUSED_TIMING
      if (lane == 0 && g_fused_timing != nullptr) {
        const unsigned long long now = fused_now();
        atomicMin(g_fused_timing + frame * 6 + 0, now);
        atomicMax(g_fused_timing + frame * 6 + 1, now);
      }
#endif
      __syncwarp();  // (the scratch is the output staging of the next GRAD item)
    }

    if (grad_valid) {
      // ---- GRAD ----------------------------------------------------------------------------------
#ifdef FUSED_TIMING
      if (lane == 0 && g_fused_timing != nullptr) atomicMin(g_fused_timing + grad_frame * 6 + 4, fused_now());
#endif
      for (bool first = true;; first = false) {
        bool valid = true;
#pragma unroll
        for (int k = 0; k < FREC; ++k) valid = valid && pre[k].y == 1u && pre[k].w == 1u;
        if (__all_sync(0xffffffffu, valid)) break;
        if (!first) __nanosleep(100);  // (the first re-read follows the prefetch without a pause)
#pragma unroll
        for (int k = 0; k < FREC; ++k) pre[k] = ld_relaxed_gpu_v4(frames + grad_frame * FREC + k);
      }
#ifdef FUSED_TIMING
      if (lane == 0 && g_fused_timing != nullptr) atomicMax(g_fused_timing + grad_frame * 6 + 5, fused_now());
#endif
      typename Policy::Frame fr;
      {
        uint32_t* words = reinterpret_cast<uint32_t*>(&fr);
#pragma unroll
        for (int k = 0; k < FREC; ++k) {
          words[2 * k] = pre[k].x;
          if (2 * k + 1 < Policy::FRAME_FLOATS) words[2 * k + 1] = pre[k].z;
        }
      }
      float* gdst = grad + grad_frame * frame_stride + ((int64_t)geom.first_atom + a0) * 3;
      int left = ng;
      const float* aux = lane_aux;
      for (int c = 0; c < nch; ++c, left -= 64, aux += 64 * 4 * AUX, gdst += CH / 4) {
        const int stage = c_stage;
        mbar_wait(s_bar + stage, c_parity);
        if (++c_stage == FUSED_STAGES) {
          c_stage = 0;
          c_parity ^= 1u;
        }
        const float4* px = lane_ring + stage * (CH / 16);
        // the staging buffer about to be written must have been read by its bulk store
        if (lane == 0) tma_wait_group_read<FUSED_OUT_BUFS - 1>();
        __syncwarp();
        float4* po = reinterpret_cast<float4*>(s_out + out_buf * CH) + 3 * lane;
        if (lane < left) {
          float x[4], y[4], z[4], gx[4], gy[4], gz[4];
          fused_unpack4(px[0], px[1], px[2], x, y, z);
          Policy::gradient(params, fr, aux, x, y, z, gx, gy, gz);
          po[0] = make_float4(gx[0], gy[0], gz[0], gx[1]);
          po[1] = make_float4(gy[1], gz[1], gx[2], gy[2]);
          po[2] = make_float4(gz[2], gx[3], gy[3], gz[3]);
        }
        if (lane + 32 < left) {
          float x[4], y[4], z[4], gx[4], gy[4], gz[4];
          fused_unpack4(px[96], px[97], px[98], x, y, z);
          Policy::gradient(params, fr, aux + 32 * 4 * AUX, x, y, z, gx, gy, gz);
          po[96] = make_float4(gx[0], gy[0], gz[0], gx[1]);
          po[97] = make_float4(gy[1], gz[1], gx[2], gy[2]);
          po[98] = make_float4(gz[2], gx[3], gy[3], gz[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        produce();
        if (elect_one()) {
          const uint32_t bytes = c == nch - 1 ? last_bytes : (uint32_t)CH;
          if (accumulate)
            tma_reduce_add_f32_1d(gdst, s_out + out_buf * CH, bytes);
          else
            tma_store_1d(gdst, s_out + out_buf * CH, bytes, pol_stream);
          tma_commit_group();
        }
        if (FUSED_OUT_BUFS == 2) out_buf ^= 1;
      }
    }
  }
  if (lane == 0) tma_wait_group_all();
}

// Host helper: geometry, workspace layout and launch.
template <typename Policy>
struct FusedLaunch {
  using Layout = FusedLayout<Policy>;
  // workspace: tagged frame constants[B], partial records[B * SS_max * NS]; SS_max for the smallest
  // sub-slice
  static size_t workspace_per_frame(int n) {
    const size_t ss_max = (size_t)div_up(n, FUSED_CHUNK_ATOMS);
    return ss_max * Policy::NS * sizeof(uint4) + 2 * sizeof(typename Policy::Frame) + 16;
  }
  static size_t workspace_fixed() { return 8 * 256 + sizeof(typename Policy::Frame); }
  // segment of the solver's gather: as many sub-slices as fit the staging and the probe batch
  static int gather_slices(int SS) {
    return std::min(SS, std::min(32 * FUSED_GATHER_UNROLL, FUSED_GATHER_BYTES / (Policy::NS * 16)));
  }
  static int gather_bytes(int SS) { return (gather_slices(SS) * Policy::NS * 16 + 127) / 128 * 128; }

  // Chooses the sub-slice size and how many warps share one auxiliary-data slot: fills the
  // resident warps with as little rounding loss as possible; larger sub-slices (fewer per-item
  // reductions) win ties.  Returns false when the problem does not fit the scheme (group too small
  // or too large, not 16-byte aligned).
  static bool plan(int n, int64_t num_atoms, int first_atom, int64_t num_frames, bool has_grad,
                   int lag_rounds, int sub_option, int slots_option, int l2_hints, int sms,
                   FusedGeom* geom) {
    if (n < 2 * FUSED_CHUNK_ATOMS || n % 4 != 0 || first_atom % 4 != 0 || num_atoms % 4 != 0)
      return false;
    double best = 0.0;
    FusedGeom g{};
    for (int sub = FUSED_CHUNK_ATOMS; sub <= 8 * FUSED_CHUNK_ATOMS; sub += FUSED_CHUNK_ATOMS) {
      if (sub_option > 0 && sub != sub_option) continue;
      const int64_t SS = div_up(n, sub);
      for (int slots = 1; slots <= FUSED_MAX_WARPS; ++slots) {
        if (slots_option > 0 && slots != slots_option) continue;
        if (Policy::AUX == 0 && slots_option <= 0 && slots != FUSED_MAX_WARPS) continue;
        const int budget =
            FUSED_SMEM_LIMIT - 1024 - gather_bytes((int)SS) - slots * Layout::slot_bytes(sub);
        if (budget <= 0) continue;
        int wpc = std::min(FUSED_MAX_WARPS, budget / Layout::WARP_BYTES);
        wpc = wpc / slots * slots;
        if (wpc < slots) continue;
        const int wps = wpc / slots;
        const int64_t total_slots = (int64_t)sms * slots;
        const int64_t K = std::min<int64_t>(total_slots / SS, div_up(num_frames, wps));
        if (K < 1) continue;
        // fraction of the machine's warp slots doing useful work, discounted for the per-item
        // reduction (about 80 instructions against 25 per atom and lane) and for thin occupancy
        const double fill = (double)(SS * K) / (double)total_slots * (double)n / (double)(SS * sub);
        const double overhead = 1.0 / (1.0 + 80.0 / (25.0 * sub / 32.0));
        // a round (all resident warps, one item each) should stay well inside the L2
        const double round_mb = (double)(K * wps) * n * 12.0 / (1 << 20);
        const double l2 = round_mb > 16.0 ? std::sqrt(16.0 / round_mb) : 1.0;
        const double score = fill * overhead * l2 * std::sqrt((double)wpc / FUSED_MAX_WARPS);
        if (score > best * 1.0001) {
          best = score;
          g.sub = sub;
          g.SS = (int)SS;
          g.K = (int)K;
          g.RR = (int)(K * wps);
          g.warps_per_cta = wpc;
          g.slots = slots;
        }
      }
    }
    if (best == 0.0) {
      if (sub_option > 0 || slots_option > 0)  // the preferred shape does not fit: let the planner pick
        return plan(n, num_atoms, first_atom, num_frames, has_grad, lag_rounds, 0, 0, l2_hints, sms,
                    geom);
      return false;
    }
    g.num_atoms = num_atoms;
    g.num_frames = num_frames;
    g.first_atom = first_atom;
    g.n = n;
    g.T = (int)div_up(num_frames, g.RR);
    g.L = std::max(1, std::min(lag_rounds, g.T));
    g.ring_rounds = has_grad ? std::min(g.T, g.L + 2) : g.T;  // (== T: records indexed by frame)
    g.l2_hints = l2_hints;
    g.gather_slices = gather_slices(g.SS);
    *geom = g;
    return true;
  }

  static int launch(const float* pos, const float* aux, const double* auxd, const FusedGeom& geom,
                    int sms,
                    const typename Policy::Params& params, void* workspace, float* value,
                    float* grad, bool accumulate, cudaStream_t stream) {
    using Frame = typename Policy::Frame;
    const size_t B = (size_t)geom.num_frames;
    Carver carve(workspace);
    constexpr size_t FREC = (Policy::FRAME_FLOATS + 1) / 2;
    uint4* frames = carve.take<uint4>(B * FREC);  // tagged frame constants (tag 0 = not yet solved)
    const size_t record_frames =
        geom.ring_rounds < geom.T ? (size_t)geom.ring_rounds * geom.RR : B;  // (ring: < B frames)
    const size_t records = record_frames * geom.SS * Policy::NS;
    uint4* partials = carve.take<uint4>(records);
    CVP_CUDA_CHECK(cudaMemsetAsync(frames, 0, B * FREC * sizeof(uint4), stream));
    CVP_CUDA_CHECK(cudaMemsetAsync(partials, 0, records * sizeof(uint4), stream));  // tag 0 = empty
    const size_t smem = (size_t)geom.slots * Layout::slot_bytes(geom.sub) +
                        (size_t)geom.warps_per_cta * Layout::WARP_BYTES + gather_bytes(geom.SS);
    static bool configured = false;
    if (!configured) {
      CVP_CUDA_CHECK(cudaFuncSetAttribute(stream_fused_kernel<Policy>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          FUSED_SMEM_LIMIT - 1024));  // (static: aux barriers)
      configured = true;
    }
    const int blocks =
        (int)std::min<int64_t>(sms, div_up((int64_t)geom.SS * geom.K, geom.slots));
    FusedGeom g = geom;
    typename Policy::Params p = params;
    int acc = accumulate ? 1 : 0;
    void* args[] = {(void*)&pos, (void*)&aux, (void*)&auxd, (void*)&g, (void*)&p,
                    (void*)&partials, (void*)&frames, (void*)&value, (void*)&grad, (void*)&acc};
    // cooperative: all CTAs are guaranteed to be co-resident (GRAD items wait for other CTAs)
    CVP_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)stream_fused_kernel<Policy>, dim3(blocks),
                                               dim3((geom.warps_per_cta + 1) * 32), args, smem, stream));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return CVP_OK;
  }
};

}  // namespace cvp
