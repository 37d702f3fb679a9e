// Cell-sorted pair sweep: the fast path of the pair-sum family (cutoff present, rectangular or no box).
//
// Only ~4 % of the |g1| x |g2| pairs of a water-density frame lie inside the cutoff, so visiting
// all of them (pair_sweep_kernel) wastes the FP32 pipe twice: on out-of-range distance tests and,
// worse, on the in-range branch running with a handful of active lanes.  This path
//   1. cell_sort_kernel   bins each group of each frame on a uniform grid (serpentine cell order,
//                         deterministic counting sort in shared memory), writes the atoms in that
//                         order (wrapped into the box) and the bounding box of every run of 8
//                         consecutive atoms (a *cluster*);
//   2. cell_sweep_kernel  keeps 32 spatially close i atoms in the registers of a warp, stages the
//                         sorted j atoms and their cluster boxes in shared memory by TMA bulk
//                         copies, discards whole clusters by box-box distance (32 clusters per
//                         ballot), applies the periodic shift once per (tile, cluster) instead of a
//                         minimum-image reduction per pair, and evaluates in two phases:
//                           phase 1: 6-instruction distance test, hits recorded in a per-lane bitmask
//                           phase 2: lanes walk their own hit bits, so the expensive pair function
//                                    (rsqrt, step function, switching, gradient) runs on real hits only.
//   3. symmetric mode     (FP32 contacts with the default step function, disjoint groups, one CTA
//                         per frame): a single sweep computes both sides of every pair.  The
//                         gradient of the j atoms is accumulated in shared memory in fixed point
//                         -- integer addition is associative, so the result does not depend on the
//                         order of the atomics.  The accumulator is a 32-bit word that is allowed
//                         to wrap plus a wrap counter: every atomic add returns the old value, the
//                         adding lane sees whether its own contribution crossed +-2^31 and, only
//                         then, bumps the counter.  The sum is exact for any input, so the scale
//                         (2^26 for the headline workload: a quantum of 1.5e-8) is set by the
//                         size of one contribution, not by a worst-case bound on the sum.
// Everything is deterministic: no floating-point atomics, fixed summation order per atom.
#pragma once

#include "pair_common.cuh"

namespace cvp {
namespace {

constexpr int CELL_CLUSTER = 8;        // atoms per j cluster
constexpr int CELL_TILE = 32;          // i atoms per warp
constexpr int CELL_SORT_THREADS = 512;
constexpr int CELL_MAX_AXIS = 16;      // cells per axis (<= 4096 cells)
constexpr int CELL_MAX_CELLS = CELL_MAX_AXIS * CELL_MAX_AXIS * CELL_MAX_AXIS;
#ifndef CELL_SWEEP_THREADS_VALUE
#define CELL_SWEEP_THREADS_VALUE 512
#endif
constexpr int CELL_SWEEP_THREADS = CELL_SWEEP_THREADS_VALUE;
constexpr int CELL_SMEM_BUDGET = 64 * 1024;  // bytes of j data per CTA (2 CTAs per SM)
// Phase 2 of the sweep keeps the hit masks in a per-warp shared array (a lane that runs out of bits
// fetches its next word with one load) and addresses the staged atoms and the j accumulators through
// explicit 32-bit shared addresses.
// Phase 1 in FP32 (SQ in the kernel): the staged record is (-2 x_j, |x_j|^2) and the test reads
// |x_j|^2 - 2 x_i . x_j < rc^2 - |x_i|^2 + margin: three FMAs and a compare per pair instead of three
// subtractions, three multiply-adds and a compare.  A conservative filter (the margin covers the
// cancellation), phase 2 forms the exact x_j - x_i = fma(-0.5, -2 x_j, -x_i) and tests r^2 < rc^2
// again.  The tag (or accumulator address) of a staged atom moves to a side array.
#ifndef CELL_SMEM_BUDGET_SYM_KB
#define CELL_SMEM_BUDGET_SYM_KB 55
#endif
constexpr int CELL_SMEM_BUDGET_SYM = CELL_SMEM_BUDGET_SYM_KB * 1024;  // ... with the j gradient accumulators
constexpr int TAG_DUMMY = -1;

__host__ __device__ inline int cell_round_up(int n, int m) { return (n + m - 1) / m * m; }

// Layout of a sorted group in memory.  The sweep stages the j atoms of a frame chunk by chunk; a
// chunk of *consecutive* sorted atoms is a slab of the box, and the i atoms of a tile next to the
// slab then find very different numbers of partners in it (the ones on the near side many, the
// others none): phase 2 of the sweep, which lasts as long as its busiest lane, ran with 17 of 32
// lanes active.  So the clusters are dealt to the chunks like cards -- cluster c of the sorted order
// goes to chunk c % nchunk, place c / nchunk -- and every chunk is a uniform sample of the whole
// frame: each (tile, chunk) visit sees 1/nchunk of every i atom's neighbourhood.  The chunks are
// contiguous in memory (one bulk copy each); a tile of 32 consecutive sorted atoms is four
// clusters at four places.
struct CellLayout {
  int nchunk;  // chunks the group is dealt to (1: plain sorted order)
  int cpc;     // clusters per chunk (a multiple of 4)
  int n_pad;   // = 8 cpc nchunk atoms, padding included
};
inline CellLayout cell_layout(int n, int jchunk_max) {
  const int clusters = (n + CELL_CLUSTER - 1) / CELL_CLUSTER;
  CellLayout l;
  l.nchunk = jchunk_max > 0 ? (clusters * CELL_CLUSTER + jchunk_max - 1) / jchunk_max : 1;
  if (l.nchunk < 1) l.nchunk = 1;
  l.cpc = cell_round_up((clusters + l.nchunk - 1) / l.nchunk, CELL_TILE / CELL_CLUSTER);
  if (l.cpc < CELL_TILE / CELL_CLUSTER) l.cpc = CELL_TILE / CELL_CLUSTER;
  l.n_pad = l.cpc * l.nchunk * CELL_CLUSTER;
  return l;
}
__device__ __forceinline__ int cell_place(int cluster, int nchunk, int cpc) {
  return nchunk == 1 ? cluster : (cluster % nchunk) * cpc + cluster / nchunk;
}

// ---- 1. sort -------------------------------------------------------------------------------------
// One CTA per (frame, group).  Dynamic shared memory: uint16 cell_of[n], uint16 order[n],
// int start[CELL_MAX_CELLS + 1], int fill[CELL_MAX_CELLS].
template <typename T, int PBC>
__global__ void __launch_bounds__(CELL_SORT_THREADS)
    cell_sort_kernel(const T* __restrict__ pos, int64_t num_atoms, const int32_t* __restrict__ idx1,
                     const int32_t* __restrict__ tag1, int n1, int n1_pad,
                     const int32_t* __restrict__ idx2, const int32_t* __restrict__ tag2, int n2,
                     int n2_pad, int nchunk1, int nchunk2, const T* __restrict__ box, int box_mode,
                     typename Vec4<T>::type* __restrict__ sorted1,
                     typename Vec4<T>::type* __restrict__ sorted2,
                     typename Vec4<T>::type* __restrict__ aabb1,
                     typename Vec4<T>::type* __restrict__ aabb2) {
  using T4 = typename Vec4<T>::type;
  extern __shared__ __align__(16) unsigned char sort_smem[];
  const int64_t b = blockIdx.x >> 1;
  const bool first = (blockIdx.x & 1) == 0;
  const int n = first ? n1 : n2;
  const int n_pad = first ? n1_pad : n2_pad;
  const int nchunk = first ? nchunk1 : nchunk2;
  const int cpc = n_pad / CELL_CLUSTER / nchunk;
  const int32_t* idx = first ? idx1 : idx2;
  const int32_t* tag = first ? tag1 : tag2;
  T4* sorted = (first ? sorted1 : sorted2) + b * n_pad;
  T4* aabb = (first ? aabb1 : aabb2) + b * (n_pad / CELL_CLUSTER) * 2;
  const T* frame = pos + b * num_atoms * 3;
  const int tid = threadIdx.x;

  int* start = reinterpret_cast<int*>(sort_smem);
  int* fill = start + CELL_MAX_CELLS + 1;
  uint16_t* cell_of = reinterpret_cast<uint16_t*>(fill + CELL_MAX_CELLS);
  uint16_t* order = cell_of + cell_round_up(n, 8);
  __shared__ T bounds[6];
  __shared__ int scan_tmp[CELL_SORT_THREADS / 32];

  // box of the grid: the periodic box, or the bounding box of the group in this frame
  T lo[3], len[3];
  if (PBC == 1) {
    // the periodic cell is taken symmetric about the origin, [-L/2, L/2): wrapping into it never
    // increases the magnitude of a coordinate, hence never coarsens its FP32 spacing (wrapping
    // x = -0.15 nm to 15.85 nm in a 16 nm box would throw away five bits of every short distance)
    const Box<T> bx = load_box(box, box_mode, b);
    len[0] = bx.ax;
    len[1] = bx.by;
    len[2] = bx.cz;
    lo[0] = T(-0.5) * len[0];
    lo[1] = T(-0.5) * len[1];
    lo[2] = T(-0.5) * len[2];
  } else {
    T mn[3] = {T(3e38), T(3e38), T(3e38)}, mx[3] = {T(-3e38), T(-3e38), T(-3e38)};
    for (int i = tid; i < n; i += CELL_SORT_THREADS) {
      const T* p = frame + 3 * (int64_t)idx[i];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        mn[c] = fmin(mn[c], p[c]);
        mx[c] = fmax(mx[c], p[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int offset = 16; offset > 0; offset >>= 1) {
        mn[c] = fmin(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], offset));
        mx[c] = fmax(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], offset));
      }
    }
    // cross-warp: reuse `start` as scratch (not yet in use)
    T* scratch = reinterpret_cast<T*>(start);
    if ((tid & 31) == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        scratch[(tid >> 5) * 6 + c] = mn[c];
        scratch[(tid >> 5) * 6 + 3 + c] = mx[c];
      }
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < CELL_SORT_THREADS / 32; ++w)
        for (int c = 0; c < 3; ++c) {
          mn[c] = fmin(mn[c], scratch[w * 6 + c]);
          mx[c] = fmax(mx[c], scratch[w * 6 + 3 + c]);
        }
      for (int c = 0; c < 3; ++c) {
        bounds[c] = mn[c];
        bounds[3 + c] = fmax(mx[c] - mn[c], T(1e-6));
      }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      lo[c] = bounds[c];
      len[c] = bounds[3 + c];
    }
    __syncthreads();
  }
  // two-level grid: coarse cells of about CELL_TILE atoms (so that 32 consecutive atoms, one warp's
  // tile, form a compact blob rather than a rod), each split in 2x2x2 octants visited in Gray-code
  // order (so that 8 consecutive atoms, one cluster, cover about two adjacent octants)
  const T volume = len[0] * len[1] * len[2];
  const T edge = cbrt(volume * T(CELL_TILE) / T(n));
  int dims[3];
  T inv_cell[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    dims[c] = max(1, min(CELL_MAX_AXIS / 2, (int)(len[c] / edge + T(0.5))));
    inv_cell[c] = T(dims[c]) / len[c];
  }
  const int ncells = dims[0] * dims[1] * dims[2] * 8;

  for (int c = tid; c <= CELL_MAX_CELLS; c += CELL_SORT_THREADS) start[c] = 0;
  for (int c = tid; c < CELL_MAX_CELLS; c += CELL_SORT_THREADS) fill[c] = 0;
  __syncthreads();

  auto wrapped = [&](int i, T (&w)[3]) {
    const T* p = frame + 3 * (int64_t)idx[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T x = p[c];
      if (PBC == 1) x = fma(-len[c], rint_t(x / len[c]), x);  // single rounding: exact k*L inside
      w[c] = x;
    }
  };
  auto cell_index = [&](const T (&w)[3]) {
    int cc[3], octant = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const T f = (w[c] - lo[c]) * inv_cell[c];
      cc[c] = max(0, min(dims[c] - 1, (int)f));
      if (f - T(cc[c]) >= T(0.5)) octant |= 1 << c;
    }
    // serpentine order: consecutive coarse cells are always spatial neighbours
    const int y = (cc[2] & 1) ? dims[1] - 1 - cc[1] : cc[1];
    const int row = cc[2] * dims[1] + y;
    const int x = (row & 1) ? dims[0] - 1 - cc[0] : cc[0];
    // position of octant code (z<<2 | y<<1 | x) along the 3-bit Gray-code walk 0,1,3,2,6,7,5,4
    const int gray_position = (0x54672310 >> (4 * octant)) & 7;
    return (row * dims[0] + x) * 8 + gray_position;
  };

  // histogram
  for (int i = tid; i < n; i += CELL_SORT_THREADS) {
    T w[3];
    wrapped(i, w);
    const int cell = cell_index(w);
    cell_of[i] = (uint16_t)cell;
    atomicAdd(&start[cell + 1], 1);
  }
  __syncthreads();
  // inclusive scan of start[1..ncells] (each thread a contiguous strip, then a block scan)
  {
    const int per = (ncells + CELL_SORT_THREADS - 1) / CELL_SORT_THREADS;
    const int begin = 1 + tid * per, end = min(ncells + 1, begin + per);
    int sum = 0;
    for (int c = begin; c < end; ++c) sum += start[c];
    int incl = sum;
#pragma unroll
    for (int offset = 1; offset < 32; offset <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl, offset);
      if ((tid & 31) >= offset) incl += v;
    }
    if ((tid & 31) == 31) scan_tmp[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
      int v = tid < CELL_SORT_THREADS / 32 ? scan_tmp[tid] : 0;
      int inc = v;
#pragma unroll
      for (int offset = 1; offset < 32; offset <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, inc, offset);
        if (tid >= offset) inc += u;
      }
      if (tid < CELL_SORT_THREADS / 32) scan_tmp[tid] = inc - v;  // exclusive warp offsets
    }
    __syncthreads();
    int running = scan_tmp[tid >> 5] + incl - sum;
    for (int c = begin; c < end; ++c) {
      running += start[c];
      start[c] = running;
    }
  }
  __syncthreads();
  // scatter (arbitrary order inside a cell), then order every cell by atom slot: deterministic
  for (int i = tid; i < n; i += CELL_SORT_THREADS) {
    const int cell = cell_of[i];
    const int slot = atomicAdd(&fill[cell], 1);
    order[start[cell] + slot] = (uint16_t)i;
  }
  __syncthreads();
  for (int cell = tid; cell < ncells; cell += CELL_SORT_THREADS) {
    const int s0 = start[cell], s1 = start[cell + 1];
    for (int a = s0 + 1; a < s1; ++a) {
      const uint16_t key = order[a];
      int k = a - 1;
      while (k >= s0 && order[k] > key) {
        order[k + 1] = order[k];
        --k;
      }
      order[k + 1] = key;
    }
  }
  __syncthreads();
  // write clusters and their bounding boxes (one thread per cluster)
  for (int cl = tid; cl < n_pad / CELL_CLUSTER; cl += CELL_SORT_THREADS) {
    T mn[3] = {T(3e38), T(3e38), T(3e38)}, mx[3] = {T(-3e38), T(-3e38), T(-3e38)};
    const int place = cell_place(cl, nchunk, cpc);
#pragma unroll
    for (int k = 0; k < CELL_CLUSTER; ++k) {
      const int s = cl * CELL_CLUSTER + k;
      T4 rec;
      if (s < n) {
        const int i = order[s];
        T w[3];
        wrapped(i, w);
        rec.x = w[0];
        rec.y = w[1];
        rec.z = w[2];
        rec.w = tag_encode(tag[i], T(0));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          mn[c] = fmin(mn[c], w[c]);
          mx[c] = fmax(mx[c], w[c]);
        }
      } else {
        rec.x = rec.y = rec.z = T(1e18);  // never in range; also flagged by the tag
        rec.w = tag_encode(TAG_DUMMY, T(0));
      }
      sorted[place * CELL_CLUSTER + k] = rec;
    }
    T4 centre, half;
    if (mn[0] > mx[0]) {  // cluster of padding only
      centre.x = centre.y = centre.z = T(1e18);
      half.x = half.y = half.z = T(0);
    } else {
      centre.x = T(0.5) * (mn[0] + mx[0]);
      centre.y = T(0.5) * (mn[1] + mx[1]);
      centre.z = T(0.5) * (mn[2] + mx[2]);
      half.x = T(0.5) * (mx[0] - mn[0]);
      half.y = T(0.5) * (mx[1] - mn[1]);
      half.z = T(0.5) * (mx[2] - mn[2]);
    }
    centre.w = half.w = T(0);
    aabb[2 * place] = centre;
    aabb[2 * place + 1] = half;
  }
}

// ---- 2. sweep ------------------------------------------------------------------------------------
// survivor list of a warp: one 16-bit entry per cluster of the chunk that may hold a partner of the
// tile -- cluster index within the chunk (8 bits), image code (2 bits per axis: 0 none, 1 +L,
// 2 -L), unsafe flag
constexpr int SURV_CAP = 256;  // = the most clusters a chunk may hold (cell_jchunk)
constexpr uint32_t SURV_UNSAFE = 1u << 14;

template <typename T>
__host__ __device__ inline int cell_jchunk(bool sym = false) {
  // j atoms per shared-memory chunk: positions + (2 records per cluster of 8) within the budget;
  // the symmetric mode adds 3 integer accumulators + 1 word of wrap counters per atom and takes a
  // little more room
  const int per_atom =
      (int)sizeof(typename Vec4<T>::type) * (CELL_CLUSTER + 2) / CELL_CLUSTER + (sym ? 16 : 0);
  const int fit = ((sym ? CELL_SMEM_BUDGET_SYM : CELL_SMEM_BUDGET) / per_atom) / CELL_TILE * CELL_TILE;
  return fit < SURV_CAP * CELL_CLUSTER ? fit : SURV_CAP * CELL_CLUSTER;
}

template <typename T>
__host__ __device__ inline size_t cell_sweep_smem(bool sym = false) {
  using T4 = typename Vec4<T>::type;
  const int jchunk = cell_jchunk<T>(sym);
  const int batch_atoms = (sizeof(T) == 4 ? 16 : 8) * CELL_CLUSTER;
  return sizeof(T4) * ((size_t)jchunk + (size_t)(jchunk / CELL_CLUSTER) * 2 +
                       (size_t)(CELL_SWEEP_THREADS / 32) * batch_atoms) +
         (size_t)(CELL_SWEEP_THREADS / 32) * SURV_CAP * sizeof(uint16_t) +
         (sym ? (size_t)jchunk * 4 * sizeof(int) : 0) +
         (size_t)(CELL_SWEEP_THREADS / 32) * (batch_atoms / 32) * 32 * sizeof(uint32_t) +  // hit words
         (sizeof(T) == 4 ? (size_t)(CELL_SWEEP_THREADS / 32) * batch_atoms * sizeof(int) : 0);  // tags (SQ)
}

// Symmetric mode: exact fixed-point accumulation.  x = round(v * scale) (F2I, |v * scale| < 2^30 by
// the choice of the scale) is added to a 32-bit word that may wrap; the add returns the old value, so
// the lane knows whether *its* contribution carried the sum across +-2^31 and then (rarely) moves
// +-1 into the wrap counter of that component.  Three 10-bit counters, biased by 512, share one word
// per atom: |wraps| <= scale * max|dE/dr| * n_i / 2^32 < 500 is part of the host's choice of scale.
constexpr int SYM_WRAP_INIT = 512 | (512 << 10) | (512 << 20);
__device__ __forceinline__ unsigned sym_fixed(float v, float scale) {
  return (unsigned)__float2int_rn(v * scale);
}
// (acc: 32-bit shared-window address of the atom's three sums; the wrap word is found from it)
__device__ __forceinline__ unsigned atoms_add(uint32_t addr, unsigned x) {
  unsigned old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(x) : "memory");
  return old;
}
__device__ __forceinline__ void sym_add3_s(uint32_t acc, uint32_t jacc_s, uint32_t jwrap_s, float vx, float vy, float vz, float scale) {
  const unsigned x0 = sym_fixed(vx, scale), x1 = sym_fixed(vy, scale), x2 = sym_fixed(vz, scale);
  const unsigned o0 = atoms_add(acc, x0), o1 = atoms_add(acc + 4, x1), o2 = atoms_add(acc + 8, x2);
  const unsigned n0 = o0 + x0, n1 = o1 + x1, n2 = o2 + x2;
  const unsigned f0 = (o0 ^ n0) & (x0 ^ n0), f1 = (o1 ^ n1) & (x1 ^ n1), f2 = (o2 ^ n2) & (x2 ^ n2);
  if ((int)(f0 | f1 | f2) < 0) {
    int bump = 0;
    if ((int)f0 < 0) bump += (int)x0 < 0 ? -1 : 1;
    if ((int)f1 < 0) bump += ((int)x1 < 0 ? -1 : 1) << 10;
    if ((int)f2 < 0) bump += ((int)x2 < 0 ? -1 : 1) << 20;
    atoms_add(jwrap_s + (acc - jacc_s) / 12u * 4u, (unsigned)bump);
  }
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float sym_value(int lo, int wraps, int component, float inv_scale) {
  const int carry = ((wraps >> (10 * component)) & 1023) - 512;
  return (float)(((double)carry * 4294967296.0 + (double)lo) * (double)inv_scale);
}

// Where the symmetric sweep leaves its gradients: with `grad` set, straight in the caller's
// [B][N][3] array (every atom of the two disjoint groups is written exactly once: the i atoms when
// their tile has seen the last chunk, the j atoms when their chunk is flushed), scaled, added when
// `accumulate`; the packed arrays and the scatter pass are then not needed.
template <typename T>
struct DirectGradient {
  T* grad = nullptr;
  int64_t num_atoms = 0;
  T scale = T(1);
  int accumulate = 0;
};

// grid = B * nsplit CTAs; CTA (b, part) owns the i tiles part, part + nsplit, ... (strided by warp).
template <typename T, int KIND, int PBC, bool OVERLAP, bool GRAD, bool FAST6, bool SYM = false>
__global__ void __launch_bounds__(CELL_SWEEP_THREADS, 2)
    cell_sweep_kernel(const typename Vec4<T>::type* __restrict__ isorted,
                      const typename Vec4<T>::type* __restrict__ jsorted,
                      const typename Vec4<T>::type* __restrict__ jaabb,
                      const typename Vec4<T>::type* __restrict__ atom_params, int ni_pad, int nj_pad,
                      int nsplit, const T* __restrict__ box, int box_mode, PairConsts<T> pc,
                      const typename PairFn<T, KIND>::Acc* __restrict__ frame_coef,
                      double* __restrict__ energy_partials,
                      typename Vec4<T>::type* __restrict__ igrad,
                      typename Vec4<T>::type* __restrict__ jgrad, float sym_scale, int sym_passes,
                      int jchunk_atoms, int i_nchunk, DirectGradient<T> direct) {
  static_assert(!SYM || (GRAD && !OVERLAP && sizeof(T) == 4 && !PairFn<T, KIND>::HAS_PARAMS),
                "symmetric mode: FP32 gradients of disjoint groups without per-atom parameters");
  using T4 = typename Vec4<T>::type;
  using Fn = PairFn<T, KIND>;
  constexpr int NCH = Fn::NCH;
  constexpr bool PRM = Fn::HAS_PARAMS;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int NWARP = CELL_SWEEP_THREADS / 32;

  extern __shared__ __align__(128) unsigned char sweep_smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ double scratch[NCH * NWARP];
  // staged batch: 128 atoms (4 hit words) in FP32, 64 atoms in FP64 (same bytes)
  constexpr int BATCH_CL = sizeof(T) == 4 ? 16 : 8;
  constexpr int NW = BATCH_CL * CELL_CLUSTER / 32;
  constexpr int BATCH_ATOMS = BATCH_CL * CELL_CLUSTER;

  const int jchunk = cell_jchunk<T>(SYM);  // the shared-memory layout; chunks hold jchunk_atoms <= jchunk
  T4* jsm = reinterpret_cast<T4*>(sweep_smem);
  T4* bsm = jsm + jchunk;  // cluster boxes of the chunk: (centre, half) pairs
  // [NWARP] staging areas: BATCH_ATOMS records, then (SQ) BATCH_ATOMS tags
  constexpr bool SQ = sizeof(T) == 4;
  constexpr int STAGE_BYTES = BATCH_ATOMS * (int)sizeof(T4) + (SQ ? BATCH_ATOMS * 4 : 0);
  unsigned char* stage_all = reinterpret_cast<unsigned char*>(bsm + (jchunk / CELL_CLUSTER) * 2);
  T4* survivors_all = reinterpret_cast<T4*>(stage_all + NWARP * STAGE_BYTES);  // [NWARP][SURV_CAP] 16-bit entries
  uint32_t* hits_all = reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(survivors_all) + NWARP * SURV_CAP);  // [NWARP][NW][32]
  int* jacc = reinterpret_cast<int*>(hits_all + NWARP * NW * 32);  // SYM: [jchunk][3] wrapping sums
  int* jwrap = jacc + 3 * jchunk;                                  // SYM: [jchunk] packed wrap counters

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t b = blockIdx.x / nsplit;
  const int part = blockIdx.x - (int)b * nsplit;
  const T4* ip = isorted + b * ni_pad;
  const T4* jp = jsorted + b * nj_pad;
  const T4* jb = jaabb + b * (nj_pad / CELL_CLUSTER) * 2;
  T4* ig = GRAD ? igrad + b * ni_pad : nullptr;
  const int ntile = ni_pad / CELL_TILE;
  const int warp_global = part * NWARP + warp;
  const int warp_stride = nsplit * NWARP;

  Box<T> bx;
  T half_box[3] = {T(3e38), T(3e38), T(3e38)};
  if (PBC == 1) {
    bx = load_box(box, box_mode, b);
    half_box[0] = T(0.5) * bx.ax;
    half_box[1] = T(0.5) * bx.by;
    half_box[2] = T(0.5) * bx.cz;
  }
  const T rc = sqrt_t(pc.cutoff2);
  using Acc = typename Fn::Acc;
  const Acc* fc = (KIND == CVP_PAIR_SOFTMIN && GRAD) ? frame_coef + 2 * b : nullptr;

  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  Acc en[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) en[c] = Acc(0);

  uint32_t phase = 0;
  const int i_cpc = ni_pad / CELL_CLUSTER / i_nchunk;
  for (int j0 = 0; j0 < nj_pad; j0 += jchunk_atoms) {
    const int jcount = min(jchunk_atoms, nj_pad - j0);
    const int ncl = jcount / CELL_CLUSTER;
    __syncthreads();  // everyone is done with the previous chunk
    if (tid == 0) {
      const uint32_t pos_bytes = (uint32_t)(jcount * sizeof(T4));
      const uint32_t box_bytes = (uint32_t)(ncl * 2 * sizeof(T4));
      mbar_expect_tx(&bar, pos_bytes + box_bytes);
      // bulk copies of at most 32 KiB each
      for (uint32_t off = 0; off < pos_bytes; off += 32768u)
        tma_load_1d(reinterpret_cast<char*>(jsm) + off, reinterpret_cast<const char*>(jp + j0) + off,
                    min(32768u, pos_bytes - off), &bar);
      for (uint32_t off = 0; off < box_bytes; off += 32768u)
        tma_load_1d(reinterpret_cast<char*>(bsm) + off,
                    reinterpret_cast<const char*>(jb + (j0 / CELL_CLUSTER) * 2) + off,
                    min(32768u, box_bytes - off), &bar);
    }
    if (SYM) {
      for (int a = tid; a < jcount * 3; a += CELL_SWEEP_THREADS) jacc[a] = 0;
      for (int a = tid; a < jcount; a += CELL_SWEEP_THREADS) jwrap[a] = SYM_WRAP_INIT;
      __syncthreads();
    }
    mbar_wait(&bar, phase);
    phase ^= 1u;

    // symmetric mode: the i tiles are visited in `sym_passes` passes and the j accumulators are
    // flushed after each, so that at most ntile * 32 / sym_passes i atoms feed one accumulator
    const int tiles_per_pass = SYM ? (ntile + sym_passes - 1) / sym_passes : ntile;
    for (int tile_first = 0; tile_first < ntile; tile_first += tiles_per_pass) {
    const int tile_last = min(ntile, tile_first + tiles_per_pass);
    for (int tile = tile_first + warp_global; tile < tile_last; tile += warp_stride) {
      // (where the lane's atom of the tile lives: see CellLayout)
      const int ipos = cell_place(tile * (CELL_TILE / CELL_CLUSTER) + (lane >> 3), i_nchunk, i_cpc) * CELL_CLUSTER + (lane & 7);
      T4 pi = ip[ipos];
      const int tagi = tag_decode(pi.w);
      // padding lanes go to the far side of the padding atoms (+1e18), so no pair of them ever
      // passes a distance test
      if (tagi < 0) pi.x = pi.y = pi.z = T(-1e18);
      T4 qi;
      if (PRM) qi = atom_params[max(tagi, 0) & TAG_INDEX_MASK];
      T gx = T(0), gy = T(0), gz = T(0);
      if (GRAD && j0 > 0) {
        const T4 g = ig[ipos];
        gx = g.x;
        gy = g.y;
        gz = g.z;
      }
      // bounding box of the tile (padding lanes excluded)
      T mn[3], mx[3];
      {
        const bool real = tagi >= 0;
        mn[0] = real ? pi.x : T(3e38);
        mn[1] = real ? pi.y : T(3e38);
        mn[2] = real ? pi.z : T(3e38);
        mx[0] = real ? pi.x : T(-3e38);
        mx[1] = real ? pi.y : T(-3e38);
        mx[2] = real ? pi.z : T(-3e38);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
          for (int offset = 16; offset > 0; offset >>= 1) {
            mn[c] = fmin(mn[c], __shfl_xor_sync(FULL, mn[c], offset));
            mx[c] = fmax(mx[c], __shfl_xor_sync(FULL, mx[c], offset));
          }
        }
      }
      T ci[3], hi[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        ci[c] = T(0.5) * (mn[c] + mx[c]);
        hi[c] = T(0.5) * (mx[c] - mn[c]);
      }
      const T4& pr = pi;  // the lane's atom as the pair loops see it

      // ---- cull: all clusters of the chunk against the box of the tile ----
      // Survivors go to this warp's list as (cluster, shift code, needs-minimum-image flag); the list
      // is complete before anything is staged, because the batches below take their clusters from
      // all over it.
      // Clusters that one shift serves (*safe*, nearly all) fill the list from the front; the ones too
      // wide for that (box barely larger than twice the cutoff, or a sparse group whose tiles and
      // clusters are large) fill it from the back and get batches of their own below: a batch with a
      // single unsafe atom would take the per-pair minimum-image path as a whole.
      uint16_t* lst_all = reinterpret_cast<uint16_t*>(survivors_all) + warp * SURV_CAP;
      int nsafe = 0, nunsafe = 0;
      for (int c0 = 0; c0 < ncl; c0 += 32) {
        const int c = c0 + lane;
        bool near = false, safe = true;
        uint32_t code = 0u;
        if (c < ncl) {
          const T4 cj = bsm[2 * c], hj = bsm[2 * c + 1];
          const T cjv[3] = {cj.x, cj.y, cj.z}, hjv[3] = {hj.x, hj.y, hj.z};
          const T len[3] = {bx.ax, bx.by, bx.cz};
          T d2 = T(0);
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            T d = cjv[a] - ci[a];
            if (PBC == 1) {
              // positions are wrapped into [-L/2, L/2]: the image count is -1, 0 or +1
              const T k = rint_t(d / len[a]);
              d -= k * len[a];
              code |= (k > T(0) ? 1u : (k < T(0) ? 2u : 0u)) << (8 + 2 * a);
            }
            const T ext = hi[a] + hjv[a];
            const T gap = fmax(fabs(d) - ext, T(0));
            d2 += gap * gap;
            // One shift serves every pair of (tile, cluster) that can be in range iff no displacement
            // exceeds L - rc on this axis: the displacements lie within |d| + ext of zero, those up to
            // L/2 are minimum images, and those between L/2 and L - rc are out of range under either
            // image.  (|d| <= L/2, so ext <= L/2 - rc is sufficient whatever d.)
            if (PBC == 1 && fabs(d) + ext + rc > len[a]) safe = false;
          }
          // clusters made of padding only sit at 1e18: the periodic shift would fold them back
          near = d2 < pc.cutoff2 && cj.x < T(1e17);
        }
        const uint32_t smask = __ballot_sync(FULL, near && safe);
        const uint32_t umask = __ballot_sync(FULL, near && !safe);
        const uint32_t below = (1u << lane) - 1u;
        if (near && safe) lst_all[nsafe + __popc(smask & below)] = (uint16_t)((uint32_t)c | code);
        if (near && !safe)
          lst_all[SURV_CAP - 1 - (nunsafe + __popc(umask & below))] = (uint16_t)((uint32_t)c | code | SURV_UNSAFE);
        nsafe += __popc(smask);
        nunsafe += __popc(umask);
      }
      __syncwarp();

      // ---- batches ----
      // A *batch* = up to NW hit words of 32 staged atoms (128 atoms in FP32), copied -- already
      // shifted to the right periodic image -- into this warp's staging buffer.  Phase 1 tests every
      // staged atom against the lane's i atom and records hits in the hit words; phase 2 lets every
      // lane walk its own hits, so the expensive pair function runs for real contacts only and the
      // j side is accumulated in shared memory.  Phase 2 lasts as long as its busiest lane, so a
      // batch is a sample of the tile's whole neighbourhood (see the striding below), not one corner
      // of it: a batch of 16 *consecutive* clusters lies on one side of the tile, and the lanes on
      // that side had 2.3 x the average number of hits (ncu: 14 of 32 lanes active).
      T4* stg = reinterpret_cast<T4*>(stage_all + warp * STAGE_BYTES);
      int* side = reinterpret_cast<int*>(stg + BATCH_ATOMS);
      // SQ: what |x_j|^2 - 2 x_i . x_j is compared with.  Every quantity of a pair in range is at
      // most M = (|x_i| + rc)^2 <= 2 (|x_i|^2 + rc^2) in size and the two sides carry about 18
      // roundings of 2^-24 M between them: 2^-17 (|x_i|^2 + rc^2) is four times that.
      T sq_thr = T(0);
      if constexpr (SQ) {
        const T xi2 = fma(pi.z, pi.z, fma(pi.y, pi.y, pi.x * pi.x));
        sq_thr = (pc.cutoff2 - xi2) + T(7.62939453125e-6) * (xi2 + pc.cutoff2);
      }
      uint32_t stg_s = smem_addr(stg);
      asm volatile("" : "+r"(stg_s));  // opaque: kept in a register, not recomputed per hit
      const uint32_t jacc_s = smem_addr(jacc), jwrap_s = smem_addr(jwrap);
      for (int segment = 0; segment < (PBC == 1 ? 2 : 1); ++segment) {  // safe clusters, then the others
      const uint16_t* lst = segment == 0 ? lst_all : lst_all + (SURV_CAP - nunsafe);
      const int nsurv = segment == 0 ? nsafe : nunsafe;
      const int natoms = nsurv * CELL_CLUSTER;
      // Batches of whole hit words (32 staged atoms): `nq` words over `nb` batches, qmin or qmin + 1
      // each, so that phase 1 tests no word of padding.  The sample is strided by *atom*: number the
      // atoms of the surviving clusters t = 8 survivor + atom; slot a of batch b is atom b + a nb
      // while a < 32 qmin, and the extra word of the first `nq - qmin nb` batches comes from the end
      // of the list.  A lane's hits come in clumps -- a cluster is mostly inside or mostly outside
      // the cutoff sphere of an i atom -- so with whole clusters in a batch the busiest lane had
      // 1.9 x the mean number of hits (17 of 32 lanes active in phase 2); one or two atoms of
      // every cluster per batch bring the counts close to Poisson.
      const int nq = (natoms + 31) >> 5;
      const int nb = (nq + NW - 1) / NW;
      const int qmin = nb > 0 ? nq / nb : 0;
      const int strided = 32 * qmin * nb;
      for (int bi = 0; bi < nb; ++bi) {
        const int words = qmin + (bi < nq - qmin * nb ? 1 : 0);  // <= NW
        const int take = 4 * words;                               // clusters, padding included
        auto slot_atom = [&](int a) { return a < 32 * qmin ? bi + a * nb : strided + 32 * bi + (a - 32 * qmin); };
        bool unsafe = false;  // (set below: the segment decides)
        // (words that phase 1 does not test need no padding)
        for (int a = lane; a < take * CELL_CLUSTER; a += 32) {
          T4 pj;
          pj.x = pj.y = pj.z = T(1e18);  // padding of a short batch: never in range
          pj.w = tag_encode(TAG_DUMMY, T(0));
          const int t = slot_atom(a);
          if (a < take * CELL_CLUSTER && t < natoms) {
            const uint32_t ent = lst[t >> 3];
            const int cl = (int)(ent & 0xffu);
            pj = jsm[cl * CELL_CLUSTER + (t & 7)];
            if (PBC == 1) {
              // d = (x_j - shift) - x_i is the minimum image for a safe cluster.  Same image (the
              // common case): the stored coordinates go in as they are, so that x_j - x_i is the
              // exact difference of the inputs -- steep pair functions (ShortestDistance: d ln w /
              // d ln r = -100) feel every rounding of a coordinate.  Other image: x_j -+ L is formed
              // in double and rounded once (x_j - L computed in FP32 from a wrapped coordinate may
              // round twice); it lies near the tile, |.| <= L/2 + rc, so it is within
              // ulp(L/2 + rc)/2 of the exact image.
              const uint32_t kx = (ent >> 8) & 3u, ky = (ent >> 10) & 3u, kz = (ent >> 12) & 3u;
              if ((kx | ky | kz) != 0u) {
                const double sx = kx == 1u ? (double)bx.ax : (kx == 2u ? -(double)bx.ax : 0.0);
                const double sy = ky == 1u ? (double)bx.by : (ky == 2u ? -(double)bx.by : 0.0);
                const double sz = kz == 1u ? (double)bx.cz : (kz == 2u ? -(double)bx.cz : 0.0);
                pj.x = T((double)pj.x - sx);
                pj.y = T((double)pj.y - sy);
                pj.z = T((double)pj.z - sz);
              }
            }
            if (SYM) {
              // the shared-window address of the atom's accumulators (-1: padding)
              pj.w = tag_encode(tag_decode(pj.w) < 0 ? -1 : (int)(jacc_s + (uint32_t)(cl * CELL_CLUSTER + (t & 7)) * 12u), T(0));
            }
          }
          if constexpr (SQ) {
            side[a] = tag_decode(pj.w);
            pj.w = fma(pj.z, pj.z, fma(pj.y, pj.y, pj.x * pj.x));
            pj.x *= T(-2), pj.y *= T(-2), pj.z *= T(-2);
          }
          stg[a] = pj;
        }
        unsafe = PBC == 1 && segment == 1;
        __syncwarp();

        // one pair: energy, i side in registers, j side (SYM) in the shared accumulators
        // (ShortestDistance also gets the displacement in double, `ex`: with beta = 100 the weight
        // moves by 1e-5 when r moves by 1e-7 r, so neither the three roundings of an FP32 r^2 nor
        // the rounding of a staged image x_j - k L can be afforded)
        auto interact = [&](int tagj, T dx, T dy, T dz, bool checked, const double (&ex)[3]) {
          // (tagj: the tag of the j atom; SYM: its accumulator slot or address)
          const T r2 = fma(dz, dz, fma(dy, dy, dx * dx));
          if (SQ && !checked && !(r2 < pc.cutoff2)) return;  // passed the filter, not the test
          // padding never passes the shifted-image test (coordinates 1e18); the minimum-image
          // reduction of the unsafe path can fold it back, so that path checks the tags
          bool ok = checked ? (tagj >= 0 && tagi >= 0) : true;
          T w = T(1);
          if (OVERLAP) {
            if (((tagi ^ tagj) & TAG_INDEX_MASK) == 0) ok = false;
            if (tagi & tagj & TAG_BOTH) w = T(0.5);
          }
          if (!ok) return;
          Acc e[NCH];
          T coef;
          if constexpr (FAST6) {
            contacts_rational6<T, SYM>(pc, r2, e[0], coef);
          } else if constexpr (KIND == CVP_PAIR_SOFTMIN) {
            Fn::eval_d(pc, ex[0], ex[1], ex[2], fc, e, coef);
          } else {
            T4 qj;
            if (PRM) qj = atom_params[tagj & TAG_INDEX_MASK];
            Fn::eval(pc, r2, qi, qj, fc, e, coef);
          }
#pragma unroll
          for (int c = 0; c < NCH; ++c) en[c] += Acc(w) * e[c];
          if (GRAD) {
            coef *= w;
            const T fx = coef * dx, fy = coef * dy, fz = coef * dz;
            gx -= fx;
            gy -= fy;
            gz -= fz;
            if constexpr (SYM) {
              sym_add3_s((uint32_t)tagj, jacc_s, jwrap_s, (float)fx, (float)fy, (float)fz, sym_scale);
            }
          }
        };

        // exact displacement of staged atom `slot` from the lane's atom: stored coordinates and the
        // image count of its cluster, in double (only instantiated for ShortestDistance)
        auto exact_displacement = [&](int slot, double (&ex)[3]) {
          const int t = slot_atom(slot);
          const uint32_t ent = lst[t >> 3];
          const T4 raw = jsm[(int)(ent & 0xffu) * CELL_CLUSTER + (t & 7)];
          const uint32_t kx = (ent >> 8) & 3u, ky = (ent >> 10) & 3u, kz = (ent >> 12) & 3u;
          ex[0] = (double)raw.x - (double)pi.x - (kx == 1u ? (double)bx.ax : (kx == 2u ? -(double)bx.ax : 0.0));
          ex[1] = (double)raw.y - (double)pi.y - (ky == 1u ? (double)bx.by : (ky == 2u ? -(double)bx.by : 0.0));
          ex[2] = (double)raw.z - (double)pi.z - (kz == 1u ? (double)bx.cz : (kz == 2u ? -(double)bx.cz : 0.0));
        };

        if (unsafe) {
          // some (tile, cluster) pair of the batch is too wide for a single shift (box barely larger
          // than twice the cutoff): per-pair minimum image, straight loop, compact code
          for (int slot = 0; slot < take * CELL_CLUSTER; ++slot) {
            const T4 pj = stg[slot];
            T dx, dy, dz;
            if constexpr (SQ)
              dx = fma(T(-0.5), pj.x, -pr.x), dy = fma(T(-0.5), pj.y, -pr.y), dz = fma(T(-0.5), pj.z, -pr.z);
            else
              dx = pj.x - pr.x, dy = pj.y - pr.y, dz = pj.z - pr.z;
            pair_min_image<PBC>(dx, dy, dz, bx);
            if (fma(dz, dz, fma(dy, dy, dx * dx)) < pc.cutoff2) {
              double ex[3] = {(double)dx, (double)dy, (double)dz};
              if constexpr (KIND == CVP_PAIR_SOFTMIN && PBC == 1) {
                exact_displacement(slot, ex);
                ex[0] -= (double)bx.ax * rint(ex[0] / (double)bx.ax);
                ex[1] -= (double)bx.by * rint(ex[1] / (double)bx.by);
                ex[2] -= (double)bx.cz * rint(ex[2] / (double)bx.cz);
              }
              interact(SQ ? side[slot] : tag_decode(pj.w), dx, dy, dz, true, ex);
            }
          }
        } else {
          // phase 1: bit positions are compile-time constants
          uint32_t* hw = hits_all + warp * (NW * 32) + lane;  // word w of this lane at hw[32 w]
          int mine = 0;
          // its 32-bit shared address, kept opaque so that it lives in a register instead of being
          // recomputed from the thread index in the (divergent) refill path
          uint32_t hw_s = smem_addr(hw);
          asm volatile("" : "+r"(hw_s));
#pragma unroll
          for (int w = 0; w < NW; ++w) {
            uint32_t m = 0u;
            if (w < words) {  // warp-uniform
#pragma unroll
              for (int u = 0; u < 32; ++u) {
                const T4 pj = stg[w * 32 + u];
                if constexpr (SQ) {
                  if (fma(pr.x, pj.x, fma(pr.y, pj.y, fma(pr.z, pj.z, pj.w))) < sq_thr) m |= 1u << u;
                } else {
                  const T dx = pj.x - pr.x, dy = pj.y - pr.y, dz = pj.z - pr.z;
                  if (fma(dz, dz, fma(dy, dy, dx * dx)) < pc.cutoff2) m |= 1u << u;
                }
              }
            }
            hw[32 * w] = m;
            mine += __popc(m);
          }
          const int rounds = __reduce_max_sync(FULL, mine);
          // phase 2: every lane walks its own hit bits through a working copy `cur` of one mask
          // word and reads its next word back from shared memory when `cur` runs out (same thread
          // wrote it: no barrier).  Lanes start at different words, so that neighbouring i atoms
          // (similar hit lists) do not reach the same j accumulator in the same instruction.
          static_assert((NW & (NW - 1)) == 0, "NW must be a power of two");
          int wsel = lane & (NW - 1);
          uint32_t cur = hw[32 * wsel];
          for (int it = 0; it < rounds; ++it) {
            if (it < mine) {
              while (cur == 0u) {
                wsel = (wsel + 1) & (NW - 1);
                cur = lds_u32(hw_s + 128u * (uint32_t)wsel);
              }
              const int k = wsel * 32 + __ffs(cur) - 1;
              cur &= cur - 1u;
              double ex[3] = {0.0, 0.0, 0.0};
              if constexpr (KIND == CVP_PAIR_SOFTMIN) exact_displacement(k, ex);
              if constexpr (SQ) {
                const float4 pj = lds_f4(stg_s + (uint32_t)k * 16u);
                const int tagj = (int)lds_u32(stg_s + (uint32_t)(BATCH_ATOMS * 16) + (uint32_t)k * 4u);
                interact(tagj, fmaf(-0.5f, pj.x, -pr.x), fmaf(-0.5f, pj.y, -pr.y), fmaf(-0.5f, pj.z, -pr.z), false, ex);
              } else {
                const T4 pj = stg[k];
                interact(tag_decode(pj.w), pj.x - pr.x, pj.y - pr.y, pj.z - pr.z, false, ex);
              }
            }
          }
        }
        __syncwarp();  // the staging buffer is rewritten by the next batch
      }
      }  // segment
      if (GRAD) {
        T4 g;
        g.x = gx;
        g.y = gy;
        g.z = gz;
        g.w = T(0);
        if (SYM && direct.grad != nullptr && j0 + jchunk_atoms >= nj_pad) {
          if (tagi >= 0) {
            T* out = direct.grad + (b * direct.num_atoms + (tagi & TAG_INDEX_MASK)) * 3;
            const bool add = direct.accumulate != 0;
            out[0] = (add ? out[0] : T(0)) + direct.scale * gx;
            out[1] = (add ? out[1] : T(0)) + direct.scale * gy;
            out[2] = (add ? out[2] : T(0)) + direct.scale * gz;
          }
        } else {
          ig[ipos] = g;
        }
      }
    }
    if (SYM) {
      // flush the j accumulators of this pass: the first pass writes, the others add (every j atom
      // belongs to exactly one chunk, and the passes come in a fixed order)
      __syncthreads();
      T4* jg = jgrad + b * nj_pad + j0;
      const float inv_scale = 1.0f / sym_scale;
      for (int a = tid; a < jcount; a += CELL_SWEEP_THREADS) {
        const int wraps = jwrap[a];
        T4 g;
        g.x = T(sym_value(jacc[3 * a], wraps, 0, inv_scale));
        g.y = T(sym_value(jacc[3 * a + 1], wraps, 1, inv_scale));
        g.z = T(sym_value(jacc[3 * a + 2], wraps, 2, inv_scale));
        g.w = T(0);
        if (direct.grad != nullptr) {
          const int tag = tag_decode(jsm[a].w);
          if (tag >= 0) {
            T* out = direct.grad + (b * direct.num_atoms + (tag & TAG_INDEX_MASK)) * 3;
            const bool add = direct.accumulate != 0 || tile_first > 0;
            out[0] = (add ? out[0] : T(0)) + direct.scale * g.x;
            out[1] = (add ? out[1] : T(0)) + direct.scale * g.y;
            out[2] = (add ? out[2] : T(0)) + direct.scale * g.z;
          }
        } else {
          if (tile_first > 0) {
            const T4 old = jg[a];
            g.x += old.x;
            g.y += old.y;
            g.z += old.z;
          }
          jg[a] = g;
        }
        jacc[3 * a] = jacc[3 * a + 1] = jacc[3 * a + 2] = 0;
        jwrap[a] = SYM_WRAP_INIT;
      }
      __syncthreads();
    }
    }
  }

  if (energy_partials != nullptr) {
    double v[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) v[c] = (double)en[c];
    block_sum<NCH>(v, scratch);
    if (tid == 0) {
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        energy_partials[((size_t)b * nsplit + part) * NCH + c] = v[c];
    }
  }
}

// packed gradients in sorted order -> [B][N][3] through the atom index carried by the tag
template <typename T>
__global__ void cell_scatter_kernel(const typename Vec4<T>::type* __restrict__ sorted,
                                    const typename Vec4<T>::type* __restrict__ packed, int n_pad,
                                    int64_t num_atoms, T scale, T* __restrict__ grad,
                                    int64_t total) {
  using T4 = typename Vec4<T>::type;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int tag = tag_decode(sorted[t].w);
    if (tag < 0) continue;
    const int64_t b = t / n_pad;
    const T4 g = packed[t];
    T* out = grad + (b * num_atoms + (tag & TAG_INDEX_MASK)) * 3;
    out[0] += scale * g.x;
    out[1] += scale * g.y;
    out[2] += scale * g.z;
  }
}

}  // namespace
}  // namespace cvp
