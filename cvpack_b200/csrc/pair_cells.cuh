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
//                         gradient of the j atoms is accumulated in shared memory as 32-bit
//                         fixed-point integers -- integer addition is associative, so the result
//                         does not depend on the order of the atomics -- with a scale chosen on the
//                         host such that even `passes`-th of the i atoms pushing one j atom in the
//                         same direction with the largest possible |dE/dr| cannot overflow.
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
constexpr int CELL_SMEM_BUDGET_SYM = 72 * 1024;  // ... with the j gradient accumulators
constexpr int TAG_DUMMY = -1;

__host__ __device__ inline int cell_round_up(int n, int m) { return (n + m - 1) / m * m; }

// ---- 1. sort -------------------------------------------------------------------------------------
// One CTA per (frame, group).  Dynamic shared memory: uint16 cell_of[n], uint16 order[n],
// int start[CELL_MAX_CELLS + 1], int fill[CELL_MAX_CELLS].
template <typename T, int PBC>
__global__ void __launch_bounds__(CELL_SORT_THREADS)
    cell_sort_kernel(const T* __restrict__ pos, int64_t num_atoms, const int32_t* __restrict__ idx1,
                     const int32_t* __restrict__ tag1, int n1, int n1_pad,
                     const int32_t* __restrict__ idx2, const int32_t* __restrict__ tag2, int n2,
                     int n2_pad, const T* __restrict__ box, int box_mode,
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
    const Box<T> bx = load_box(box, box_mode, b);
    lo[0] = lo[1] = lo[2] = T(0);
    len[0] = bx.ax;
    len[1] = bx.by;
    len[2] = bx.cz;
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
      if (PBC == 1) x = fma(-len[c], floor(x / len[c]), x);  // single rounding: exact k*L inside
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
      sorted[s] = rec;
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
    aabb[2 * cl] = centre;
    aabb[2 * cl + 1] = half;
  }
}

// ---- 2. sweep ------------------------------------------------------------------------------------
template <typename T>
__host__ __device__ inline int cell_jchunk(bool sym = false) {
  // j atoms per shared-memory chunk: positions + (2 records per cluster of 8) within the budget;
  // the symmetric mode adds 3 integer accumulators per atom and takes a little more room
  const int per_atom =
      (int)sizeof(typename Vec4<T>::type) * (CELL_CLUSTER + 2) / CELL_CLUSTER + (sym ? 12 : 0);
  return ((sym ? CELL_SMEM_BUDGET_SYM : CELL_SMEM_BUDGET) / per_atom) / CELL_TILE * CELL_TILE;
}

template <typename T>
__host__ __device__ inline size_t cell_sweep_smem(bool sym = false) {
  using T4 = typename Vec4<T>::type;
  const int jchunk = cell_jchunk<T>(sym);
  const int batch_atoms = (sizeof(T) == 4 ? 16 : 8) * CELL_CLUSTER;
  return sizeof(T4) * ((size_t)jchunk + (size_t)(jchunk / CELL_CLUSTER) * 2 +
                       (size_t)(CELL_SWEEP_THREADS / 32) * (batch_atoms + 32)) +
         (sym ? (size_t)jchunk * 3 * sizeof(int) : 0);
}

// fixed-point conversion of the symmetric mode: round(v * scale) for |v * scale| < 2^22, without
// leaving the FMA/ALU pipes (the integer sits in the mantissa of v * scale + 1.5 * 2^23)
__device__ __forceinline__ int sym_fixed(float v, float scale) {
  return __float_as_int(fmaf(v, scale, 12582912.0f)) - 0x4B400000;
}

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
                      typename Vec4<T>::type* __restrict__ jgrad = nullptr, float sym_scale = 0.f,
                      int sym_passes = 1) {
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

  const int jchunk = cell_jchunk<T>(SYM);
  T4* jsm = reinterpret_cast<T4*>(sweep_smem);
  T4* bsm = jsm + jchunk;  // cluster boxes of the chunk: (centre, half) pairs
  T4* stage_all = bsm + (jchunk / CELL_CLUSTER) * 2;       // [NWARP][BATCH_ATOMS]
  T4* survivors_all = stage_all + NWARP * BATCH_ATOMS;    // [NWARP][32]
  int* jacc = reinterpret_cast<int*>(survivors_all + NWARP * 32);  // SYM: [jchunk][3]

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
  for (int j0 = 0; j0 < nj_pad; j0 += jchunk) {
    const int jcount = min(jchunk, nj_pad - j0);
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
      T4 pi = ip[tile * CELL_TILE + lane];
      const int tagi = tag_decode(pi.w);
      // padding lanes go to the far side of the padding atoms (+1e18), so no pair of them ever
      // passes a distance test
      if (tagi < 0) pi.x = pi.y = pi.z = T(-1e18);
      T4 qi;
      if (PRM) qi = atom_params[max(tagi, 0) & TAG_INDEX_MASK];
      T gx = T(0), gy = T(0), gz = T(0);
      if (GRAD && j0 > 0) {
        const T4 g = ig[tile * CELL_TILE + lane];
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

      // A *batch* = up to BATCH_CL surviving clusters (128 atoms in FP32), copied -- already shifted
      // to the right periodic image -- into this warp's staging buffer.  Phase 1 tests every staged
      // atom against the lane's i atom and records hits in NW 32-bit masks; phase 2 lets every lane
      // walk its own hits, so the expensive pair function runs for real contacts only.  The larger
      // the batch, the closer the busiest lane is to the average one.
      uint32_t hits[NW];
#pragma unroll
      for (int w = 0; w < NW; ++w) hits[w] = 0u;
      int nbatch = 0;  // clusters staged so far
      bool batch_unsafe = false;
      T4* stg = stage_all + warp * BATCH_ATOMS;
      T4* lst = survivors_all + warp * 32;
      auto run_phases = [&](auto unsafe_tag) {
        constexpr bool UNSAFE = decltype(unsafe_tag)::value;
        auto test = [&](int slot) -> bool {
          const T4 pj = stg[slot];
          T dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
          if (UNSAFE) pair_min_image<PBC>(dx, dy, dz, bx);
          return fma(dz, dz, fma(dy, dy, dx * dx)) < pc.cutoff2;
        };
        if (nbatch == BATCH_CL) {
          // full batch: bit positions are compile-time constants
#pragma unroll
          for (int w = 0; w < NW; ++w) {
#pragma unroll
            for (int u = 0; u < 32; ++u) {
              if (test(w * 32 + u)) hits[w] |= 1u << u;
            }
          }
        } else {
          for (int c = 0; c < nbatch; ++c) {
            uint32_t m = 0;
#pragma unroll
            for (int u = 0; u < CELL_CLUSTER; ++u) {
              if (test(c * CELL_CLUSTER + u)) m |= 1u << u;
            }
            m <<= (c & 3) * CELL_CLUSTER;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
              if ((c >> 2) == w) hits[w] |= m;
            }
          }
        }
        int mine = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) mine += __popc(hits[w]);
        const int rounds = __reduce_max_sync(FULL, mine);
        for (int it = 0; it < rounds; ++it) {
          int k = -1;
#pragma unroll
          for (int w = NW - 1; w >= 0; --w) {
            if (hits[w] != 0u) k = w;
          }
          if (k >= 0) {
#pragma unroll
            for (int w = 0; w < NW; ++w) {
              if (k == w) {
                k = w * 32 + __ffs(hits[w]) - 1;
                hits[w] &= hits[w] - 1u;
                break;
              }
            }
            const T4 pj = stg[k];
            const int tagj = tag_decode(pj.w);  // (SYM: the accumulator slot)
            T dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
            if (UNSAFE) pair_min_image<PBC>(dx, dy, dz, bx);
            const T r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            // padding never passes the shifted-image test (coordinates 1e18); the minimum-image
            // reduction of the unsafe path can fold it back, so that path checks the tags
            bool ok = UNSAFE ? (tagj >= 0 && tagi >= 0) : true;
            T w = T(1);
            if (OVERLAP) {
              if (((tagi ^ tagj) & TAG_INDEX_MASK) == 0) ok = false;
              if (tagi & tagj & TAG_BOTH) w = T(0.5);
            }
            if (ok) {
              Acc e[NCH];
              T coef;
              if constexpr (FAST6) {
                contacts_rational6(pc, r2, e[0], coef);
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
                if (SYM) {
                  int* acc = jacc + tagj * 3;
                  atomicAdd(acc, sym_fixed((float)fx, sym_scale));
                  atomicAdd(acc + 1, sym_fixed((float)fy, sym_scale));
                  atomicAdd(acc + 2, sym_fixed((float)fz, sym_scale));
                }
              }
            }
          }
        }
      };
      auto run_batch = [&]() {
        __syncwarp();
        // the per-pair minimum image is needed only when some (tile, cluster) pair of the batch is
        // too wide for a single shift; a warp-uniform branch keeps it out of the common path
        if (PBC == 1 && batch_unsafe)
          run_phases(std::true_type{});
        else
          run_phases(std::false_type{});
        __syncwarp();
        nbatch = 0;
        batch_unsafe = false;
      };

      for (int c0 = 0; c0 < ncl; c0 += 32) {
        // ---- cull: lane l tests cluster c0 + l against the tile box ----
        const int c = c0 + lane;
        bool near = false, safe = true;
        T sh[3] = {T(0), T(0), T(0)};
        if (c < ncl) {
          const T4 cj = bsm[2 * c], hj = bsm[2 * c + 1];
          const T cjv[3] = {cj.x, cj.y, cj.z}, hjv[3] = {hj.x, hj.y, hj.z};
          const T len[3] = {bx.ax, bx.by, bx.cz};
          T d2 = T(0);
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            T d = cjv[a] - ci[a];
            if (PBC == 1) {
              sh[a] = len[a] * rint_t(d / len[a]);
              d -= sh[a];
            }
            const T ext = hi[a] + hjv[a];
            const T gap = fmax(fabs(d) - ext, T(0));
            d2 += gap * gap;
            // one shift serves every pair of (tile, cluster) iff ext <= L/2 - rc on this axis
            if (PBC == 1 && ext + rc > half_box[a]) safe = false;
          }
          // clusters made of padding only sit at 1e18: the periodic shift would fold them back
          near = d2 < pc.cutoff2 && cj.x < T(1e17);
        }
        // survivors of this window, compacted in cluster order: (shift, cluster index)
        const uint32_t cmask = __ballot_sync(FULL, near);
        if (cmask == 0u) continue;
        const bool window_unsafe = __any_sync(FULL, near && !safe);
        if (near) {
          T4 ent;
          ent.x = sh[0];
          ent.y = sh[1];
          ent.z = sh[2];
          ent.w = tag_encode(c, T(0));
          lst[__popc(cmask & ((1u << lane) - 1u))] = ent;
        }
        __syncwarp();
        const int nsurv = __popc(cmask);
        for (int s0 = 0; s0 < nsurv;) {
          const int take = min(BATCH_CL - nbatch, nsurv - s0);
          // stage `take` clusters, 4 per pass (one atom per lane)
          for (int a = lane; a < take * CELL_CLUSTER; a += 32) {
            const T4 ent = lst[s0 + (a >> 3)];
            T4 pj = jsm[tag_decode(ent.w) * CELL_CLUSTER + (a & 7)];
            pj.x -= ent.x;  // d = (x_j - shift) - x_i is the minimum image for a safe cluster
            pj.y -= ent.y;
            pj.z -= ent.z;
            if (SYM)  // the accumulator slot of the atom within the chunk (-1: padding)
              pj.w = tag_encode(
                  tag_decode(pj.w) < 0 ? -1 : tag_decode(ent.w) * CELL_CLUSTER + (a & 7), T(0));
            stg[nbatch * CELL_CLUSTER + a] = pj;
          }
          batch_unsafe |= window_unsafe;
          nbatch += take;
          s0 += take;
          if (nbatch == BATCH_CL) run_batch();
        }
        __syncwarp();  // the survivor list is rewritten by the next window
      }
      if (nbatch > 0) run_batch();
      if (GRAD) {
        T4 g;
        g.x = gx;
        g.y = gy;
        g.z = gz;
        g.w = T(0);
        ig[tile * CELL_TILE + lane] = g;
      }
    }
    if (SYM) {
      // flush the j accumulators of this pass: the first pass writes, the others add (every j atom
      // belongs to exactly one chunk, and the passes come in a fixed order)
      __syncthreads();
      T4* jg = jgrad + b * nj_pad + j0;
      const T inv_scale = T(1) / T(sym_scale);
      for (int a = tid; a < jcount; a += CELL_SWEEP_THREADS) {
        T4 g;
        g.x = T(jacc[3 * a]) * inv_scale;
        g.y = T(jacc[3 * a + 1]) * inv_scale;
        g.z = T(jacc[3 * a + 2]) * inv_scale;
        g.w = T(0);
        if (tile_first > 0) {
          const T4 old = jg[a];
          g.x += old.x;
          g.y += old.y;
          g.z += old.z;
        }
        jg[a] = g;
        jacc[3 * a] = jacc[3 * a + 1] = jacc[3 * a + 2] = 0;
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
