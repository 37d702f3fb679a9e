// Many references, one atom set: the fast path of PathInRMSDSpace (path_in_rmsd_space.py:137-140
// builds one RMSD per milestone; when every milestone lists the same atoms in the same order, the
// G alignments of a frame share the frame's coordinates).
//
// The generic kernels treat the G x n (group, atom) entries of a frame independently: every entry
// re-reads its 24-byte reference row from L2 for the sums and again for the gradient -- 30 MB per
// frame at G = 64, n = 20 000 -- and both passes end up L2-latency bound.  Here the two passes are
// register-tiled double-precision contractions that stage frames and references in shared memory,
// so a reference row is read once per 32 frames:
//
//   covariance  S[f][g] (3x3) = sum_a x[f][a] (x) y^[g][a]          [3B x n] . [n x 3G]
//               CTA tile 32 frames x 32 groups, thread tile 2 x 2 (36 DFMA per atom and thread),
//               atoms in chunks of 8 through a double-buffered shared-memory stage, split over
//               the atoms (SK slices) to fill the machine.  The partial sums land in the layout
//               rmsd_solve_kernel reads; the moments sum x, sum |x|^2 do not depend on the group
//               and are produced once (group 0's slices).
//   gradient    g[f][a] = F_f (x[f][a] - c_f) - sum_g W[f][g] y^[g][a],   W = factor[f][g] U^T[f][g]
//               [3B x 3G] . [3G x n]: CTA tile 32 frames x 64 atoms, thread tile 2 frames x 4 atoms
//               (72 DFMA per group and thread), groups in chunks of 16.
// Double precision throughout, like the generic kernels: (x - c) - U^T y^ cancels three digits
// and the path weights cancel again across milestones.  Deterministic (fixed summation order).
#pragma once

#include "common.cuh"

namespace cvp {
namespace {

constexpr int MULTI_THREADS = 256;
constexpr int MULTI_TF = 32;      // frames per CTA tile
constexpr int MULTI_TG = 32;      // groups per CTA tile (covariance)
constexpr int MULTI_KA = 8;       // atoms per stage (covariance)
constexpr int MULTI_TA = 128;     // atoms per CTA tile (gradient)
#ifndef MULTI_GRAD_UNROLL
#define MULTI_GRAD_UNROLL 1
#endif
constexpr int MULTI_GRAD_UNROLL_N = MULTI_GRAD_UNROLL;  // group steps unrolled in the gradient kernel
constexpr int MULTI_KG = 8;       // groups per stage (gradient), two stages in flight

// ---- covariance ----------------------------------------------------------------------------------
// grid = (frame tiles, group tiles, SK); partials[(b * ns + g * SK + sk) * 13 + ...]
__global__ void __launch_bounds__(MULTI_THREADS, 2)
    rmsd_multi_cov_kernel(const float* __restrict__ pos, int64_t num_atoms, int first_atom, int n,
                          const double* __restrict__ ref /* [G][n][3], centred */, int num_groups,
                          int64_t num_frames, int SK, double* __restrict__ partials) {
  __shared__ __align__(16) double Xs[2][MULTI_KA][MULTI_TF][3];
  __shared__ __align__(16) double Ys[2][MULTI_KA][MULTI_TG][3];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;  // frames 2ty, 2ty+1; groups 2tx, 2tx+1 of the tile
  const int64_t f0 = (int64_t)blockIdx.x * MULTI_TF;
  const int g0 = blockIdx.y * MULTI_TG;
  const int sk = blockIdx.z;
  const int ns = num_groups * SK;
  // this CTA's atom range (multiples of MULTI_KA except at the end)
  const int chunks_total = (n + MULTI_KA - 1) / MULTI_KA;
  const int chunks_per = (chunks_total + SK - 1) / SK;
  const int a_begin = min(n, sk * chunks_per * MULTI_KA);
  const int a_end = min(n, (sk + 1) * chunks_per * MULTI_KA);

  double acc[2][2][9];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int k = 0; k < 9; ++k) acc[i][j][k] = 0.0;

  // loader roles: 32 rows x ROW values per tile and stage, NQ per thread.  X: (frame, atom, comp)
  // with the ROW floats of a frame's MULTI_KA atoms contiguous in global memory; Y: (group, atom,
  // comp) likewise.
  constexpr int ROW = MULTI_KA * 3;
  constexpr int NQ = MULTI_TF * ROW / MULTI_THREADS;
  static_assert(MULTI_TF * ROW % MULTI_THREADS == 0 && MULTI_TF == MULTI_TG, "loader shape");
  static_assert(MULTI_TA == 128, "the gradient's Y stage layout assumes 16 threads x 8 atoms");
  // (roles fixed per thread, so that a stage costs no index arithmetic: element q of this thread is
  // value `rest` of row `row` -- a frame for X, a group for Y -- of the stage's 8 atoms x 3)
  float xr[NQ];
  double yr[NQ];
  // 32-bit offsets from the tile's first frame / group (the host checks that they fit); an
  // invalid row (frame or group beyond the end) is encoded as an atom index that is never in range
  const float* x_base = pos + (f0 * num_atoms + first_atom) * 3;
  const double* y_base = ref + (int64_t)g0 * n * 3;
  uint32_t x_off[NQ], y_off[NQ];
  int s_dst[NQ], a_rel_x[NQ], a_rel_y[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int idx = tid + q * MULTI_THREADS;
    const int row = idx / ROW, rest = idx - row * ROW;
    const int a = rest / 3, c = rest - a * 3;
    s_dst[q] = (a * MULTI_TF + row) * 3 + c;
    a_rel_x[q] = f0 + row < num_frames ? a : (1 << 30);
    a_rel_y[q] = g0 + row < num_groups ? a : (1 << 30);
    x_off[q] = (uint32_t)row * (uint32_t)(num_atoms * 3) + (uint32_t)rest;
    y_off[q] = (uint32_t)row * (uint32_t)(n * 3) + (uint32_t)rest;
  }
  auto load_regs = [&](int a0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      xr[q] = a0 + a_rel_x[q] < a_end ? __ldg(x_base + x_off[q] + (int64_t)a0 * 3) : 0.f;
      yr[q] = a0 + a_rel_y[q] < a_end ? __ldg(y_base + y_off[q] + (int64_t)a0 * 3) : 0.0;
    }
  };
  auto store_regs = [&](int buf) {
    double* xs = &Xs[buf][0][0][0];
    double* ys = &Ys[buf][0][0][0];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      xs[s_dst[q]] = (double)xr[q];
      ys[s_dst[q]] = yr[q];
    }
  };

  int buf = 0;
  if (a_begin < a_end) {
    load_regs(a_begin);
    store_regs(0);
  }
  __syncthreads();
  for (int a0 = a_begin; a0 < a_end; a0 += MULTI_KA) {
    const bool more = a0 + MULTI_KA < a_end;
    if (more) load_regs(a0 + MULTI_KA);
#pragma unroll 4
    for (int k = 0; k < MULTI_KA; ++k) {
      const double2* px = reinterpret_cast<const double2*>(&Xs[buf][k][2 * ty][0]);
      const double2* py = reinterpret_cast<const double2*>(&Ys[buf][k][2 * tx][0]);
      const double2 x01 = px[0], x23 = px[1], x45 = px[2];  // frame 2ty: x y z, frame 2ty+1: x y z
      const double2 y01 = py[0], y23 = py[1], y45 = py[2];
      const double xv[2][3] = {{x01.x, x01.y, x23.x}, {x23.y, x45.x, x45.y}};
      const double yv[2][3] = {{y01.x, y01.y, y23.x}, {y23.y, y45.x, y45.y}};
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int d = 0; d < 3; ++d) acc[i][j][3 * c + d] = fma(xv[i][c], yv[j][d], acc[i][j][3 * c + d]);
    }
    if (more) store_regs(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int64_t f = f0 + 2 * ty + i;
    if (f >= num_frames) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int g = g0 + 2 * tx + j;
      if (g >= num_groups) continue;
      double* out = partials + ((size_t)f * ns + (size_t)g * SK + sk) * 13;
#pragma unroll
      for (int k = 0; k < 9; ++k) out[4 + k] = acc[i][j][k];
      // (out[0..3] of group 0 come from rmsd_multi_moments_kernel; the other groups' are unused)
    }
  }
}

// sum x, sum |x|^2 of a frame over the same atom slices; one warp per (frame, slice), written to
// group 0's slice records
__global__ void rmsd_multi_moments_kernel(const float* __restrict__ pos, int64_t num_atoms,
                                          int first_atom, int n, int num_groups,
                                          int64_t num_frames, int SK,
                                          double* __restrict__ partials) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (w >= num_frames * SK) return;
  const int64_t f = w / SK;
  const int sk = (int)(w - f * SK);
  const int chunks_total = (n + MULTI_KA - 1) / MULTI_KA;
  const int chunks_per = (chunks_total + SK - 1) / SK;
  const int a_begin = min(n, sk * chunks_per * MULTI_KA);
  const int a_end = min(n, (sk + 1) * chunks_per * MULTI_KA);
  const float* p = pos + (f * num_atoms + first_atom) * 3;
  double v[4] = {0, 0, 0, 0};
  for (int a = a_begin + lane; a < a_end; a += 32) {
    const double x = p[3 * (int64_t)a], y = p[3 * (int64_t)a + 1], z = p[3 * (int64_t)a + 2];
    v[0] += x;
    v[1] += y;
    v[2] += z;
    v[3] = fma(x, x, fma(y, y, fma(z, z, v[3])));
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
    double* out = partials + ((size_t)f * num_groups * SK + sk) * 13;
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = v[k];
  }
}

// ---- gradient ------------------------------------------------------------------------------------
// W[f][g][0..8] = factor[f][g] * U^T[f][g], W[f][g][9] = factor[f][g]: one thread per (frame, group),
// so that the W tiles of the gradient kernel are plain copies.
template <typename Solution>
__global__ void rmsd_multi_weights_kernel(const Solution* __restrict__ solution,
                                          const double* __restrict__ factor, int64_t total,
                                          double* __restrict__ weights /* [frame][group][10] */) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const double fac = factor[t];
  double2* out = reinterpret_cast<double2*>(weights + t * 10);
  const double* ut = solution[t].ut;
#pragma unroll
  for (int k = 0; k < 4; ++k) out[k] = make_double2(fac * ut[2 * k], fac * ut[2 * k + 1]);
  out[4] = make_double2(fac * ut[8], fac);
}

// 8- and 16-byte asynchronous copies global -> shared (LDGSTS); src_bytes = 0 fills with zeros
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gmem_src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_addr(smem_dst)),
               "l"(gmem_src), "r"(valid ? 8 : 0)
               : "memory");
}
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr(smem_dst)),
               "l"(gmem_src), "r"(valid ? 16 : 0)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// grid = (frame tiles, atom tiles).  Groups in stages of MULTI_KG through a double-buffered
// shared-memory ring filled by cp.async: the copies of stage k+1 are in flight while stage k is
// multiplied (the synchronous version stalled on its tile loads: FP64 pipe 28 % busy).
__global__ void __launch_bounds__(MULTI_THREADS, 1)
    rmsd_multi_gradient_kernel(const float* __restrict__ pos, int64_t num_atoms, int first_atom,
                               int n, const double* __restrict__ ref, int num_groups,
                               int64_t num_frames, const double* __restrict__ weights,
                               const double* __restrict__ centroid /* [frame][nseg][3], seg 0 */,
                               int nseg, float* __restrict__ grad, int accumulate) {
  extern __shared__ __align__(16) double multi_smem[];
  constexpr int W_STAGE = MULTI_KG * MULTI_TF * 10, Y_STAGE = MULTI_KG * MULTI_TA * 3;
  const int tid = threadIdx.x;
  // thread tile 2 frames x 8 atoms: 42 doubles from shared memory per 144 DFMA.  (A 128-bit
  // shared load costs 4 wavefronts, one per quarter warp, whatever is broadcast; with the 2 x 4
  // tile ncu showed the shared-memory pipe 79 % busy and the FP64 pipe 41 %.)
  const int ty = tid >> 4, tx = tid & 15;  // frames 2ty, 2ty+1; atoms 8tx .. 8tx+7 of the tile
  const int64_t f0 = (int64_t)blockIdx.x * MULTI_TF;
  const int a0 = blockIdx.y * MULTI_TA;

  // stage layout: Ws[gl][fl][10] (f U^T, f); Ys[gl]: the 24 doubles of thread tx (8 atoms x 3) as
  // 12 double2 pieces, piece k at double2 index k * 16 + tx: a half-warp reads 256 contiguous
  // bytes (in atom order the stride between threads is a multi-way bank conflict).
  // Copy roles, fixed per thread so that a stage costs no index arithmetic:
  //   W: thread -> (frame tid / MULTI_KG, group tid % MULTI_KG), its 80-byte record as 5 x 16 bytes;
  //   Y: a group's tile row is 3 * MULTI_TA doubles; thread tid < 3 * MULTI_TA / 2 copies double2
  //      `tid` of every group's row (n even: rows are 16-byte aligned), else two single doubles.
  static_assert(MULTI_TF * MULTI_KG == MULTI_THREADS && 3 * MULTI_TA / 2 <= MULTI_THREADS, "copy roles");
  const int w_fl = tid / MULTI_KG, w_gl = tid - w_fl * MULTI_KG;
  const bool w_frame_ok = f0 + w_fl < num_frames;
  const double* w_src = weights + ((w_frame_ok ? f0 + w_fl : 0) * num_groups + w_gl) * 10;
  const int w_dst = (w_gl * MULTI_TF + w_fl) * 10;
  const bool y_role = tid < 3 * MULTI_TA / 2;
  const bool y_aligned = (n & 1) == 0;
  // doubles 2 tid, 2 tid + 1 of the row: atoms (2 tid) / 3 and (2 tid + 1) / 3
  const bool y_ok0 = y_role && a0 + (2 * tid) / 3 < n, y_ok1 = y_role && a0 + (2 * tid + 1) / 3 < n;
  const double* y_src = ref + (int64_t)a0 * 3 + (y_ok0 ? 2 * tid : 0);
  int y_dst = 0;
  {
    const int d = 2 * tid, owner = d / 24, q = d - owner * 24;  // (q even: both doubles in piece q / 2)
    y_dst = ((q >> 1) * 16 + owner) << 1;
  }
  auto issue = [&](int g0, int buf) {
    double* Ws = multi_smem + (size_t)buf * (W_STAGE + Y_STAGE);
    double* Ys = Ws + W_STAGE;
    {
      const bool valid = w_frame_ok && g0 + w_gl < num_groups;
      const double* src = valid ? w_src + (int64_t)g0 * 10 : weights;
#pragma unroll
      for (int k = 0; k < 5; ++k) cp_async_16(Ws + w_dst + 2 * k, src + 2 * k, valid);
    }
    if (y_role) {
#pragma unroll
      for (int gl = 0; gl < MULTI_KG; ++gl) {
        const bool group_ok = g0 + gl < num_groups;
        const double* src = y_src + (int64_t)(g0 + gl) * n * 3;
        double* dst = Ys + gl * (MULTI_TA * 3) + y_dst;
        if (y_aligned) {
          // (n even: 3 (n - a0) is even, so both doubles are valid or neither)
          cp_async_16(dst, group_ok && y_ok0 ? src : ref, group_ok && y_ok0);
        } else {
          cp_async_8(dst, group_ok && y_ok0 ? src : ref, group_ok && y_ok0);
          cp_async_8(dst + 1, group_ok && y_ok1 ? src + 1 : ref, group_ok && y_ok1);
        }
      }
    }
    cp_async_commit();
  };

  double acc[2][8][3];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[i][j][c] = 0.0;
  double fsum[2] = {0.0, 0.0};

  issue(0, 0);
  int buf = 0;
  for (int g0 = 0; g0 < num_groups; g0 += MULTI_KG) {
    const bool more = g0 + MULTI_KG < num_groups;
    if (more) {
      issue(g0 + MULTI_KG, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const double* Ws = multi_smem + (size_t)buf * (W_STAGE + Y_STAGE);
    const double* Ys = Ws + W_STAGE;
#pragma unroll MULTI_GRAD_UNROLL_N
    for (int gl = 0; gl < MULTI_KG; ++gl) {
      double w[2][10];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double2* pw = reinterpret_cast<const double2*>(Ws + (gl * MULTI_TF + 2 * ty + i) * 10);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const double2 v = pw[k];
          w[i][2 * k] = v.x;
          w[i][2 * k + 1] = v.y;
        }
        fsum[i] += w[i][9];
      }
      const double2* py = reinterpret_cast<const double2*>(Ys + gl * (MULTI_TA * 3)) + tx;
      double y[24];
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const double2 v = py[k * 16];
        y[2 * k] = v.x;
        y[2 * k + 1] = v.y;
      }
      // (d outermost: 48 independent accumulators between two uses of the same one)
#pragma unroll
      for (int d = 2; d >= 0; --d)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[i][j][c] = fma(w[i][3 * c + d], y[3 * j + d], acc[i][j][c]);
    }
    __syncthreads();  // the stage is refilled by the copies issued in the next iteration
    buf ^= 1;
  }
  // g = F (x - c) - acc
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int64_t f = f0 + 2 * ty + i;
    if (f >= num_frames) continue;
    const double* cf = centroid + (size_t)f * nseg * 3;
    const double c[3] = {cf[0], cf[1], cf[2]};
    const float* px = pos + (f * num_atoms + first_atom) * 3;
    float* pg = grad + (f * num_atoms + first_atom) * 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int a = a0 + 8 * tx + j;
      if (a >= n) continue;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double v = fsum[i] * ((double)px[3 * (int64_t)a + d] - c[d]) - acc[i][j][d];
        if (accumulate)
          pg[3 * (int64_t)a + d] += (float)v;
        else
          pg[3 * (int64_t)a + d] = (float)v;
      }
    }
  }
}

// ---- FP64 tensor-core versions (mma.sync.m8n8k4.f64, DMMA) ----------------------------------------
// Both contractions above run at 41-51 % of the FP64 pipe: what limits them is shared memory, a
// 128-bit shared load costs four wavefronts whatever is broadcast (ncu, round 1).  With the tensor
// core instruction a warp multiplies an 8 x 4 by a 4 x 8 tile (256 DFMA) from ONE double per
// thread and operand: the operand tiles are stored in shared memory in fragment order (32
// consecutive doubles per 8 x 4 tile, lane l reads element l: conflict-free LDS.64) and a warp tile
// of 3 x 6 (covariance) / 3 x 8 (gradient) MMA tiles needs 9 / 11 loads per 18 / 24 instructions.
//   fragment layouts (PTX ISA, m8n8k4 .f64):  A[row = lane / 4][k = lane % 4],
//   B[k = lane % 4][col = lane / 4],  C[row = lane / 4][col = 2 (lane % 4) + {0, 1}]
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

constexpr int MMA_KA = 16;  // atoms per stage of the covariance (4 k-steps)
constexpr int MMA_ROWS = 3 * MULTI_TF;  // 96 = 12 tiles of 8 rows (frame, component)

// covariance: M = (frame, c) 96 rows, N = (group, d) 96 columns, K = atoms.
// grid = (frame tiles, group tiles, SK); same partial-sum records as rmsd_multi_cov_kernel.
__global__ void __launch_bounds__(MULTI_THREADS, 2)
    rmsd_multi_cov_mma_kernel(const float* __restrict__ pos, int64_t num_atoms, int first_atom, int n,
                              const double* __restrict__ ref /* [G][n][3], centred */, int num_groups,
                              int64_t num_frames, int SK, double* __restrict__ partials) {
  // [buffer][k-step][tile][fragment element]
  __shared__ __align__(16) double As[2][MMA_KA / 4][MMA_ROWS / 8][32];
  __shared__ __align__(16) double Bs[2][MMA_KA / 4][MMA_ROWS / 8][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;  // row tiles 3 wm .. 3 wm + 2, column tiles 6 wn .. 6 wn + 5
  const int64_t f0 = (int64_t)blockIdx.x * MULTI_TF;
  const int g0 = blockIdx.y * MULTI_TG;
  const int sk = blockIdx.z;
  const int ns = num_groups * SK;
  const int chunks_total = (n + MMA_KA - 1) / MMA_KA;
  const int chunks_per = (chunks_total + SK - 1) / SK;
  const int a_begin = min(n, sk * chunks_per * MMA_KA);
  const int a_end = min(n, (sk + 1) * chunks_per * MMA_KA);

  double acc[3][6][2];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // loader: a stage is 32 rows (frames for A, groups for B) x 48 values (16 atoms x 3), contiguous
  // in global memory per row.  Thread t takes the 6 consecutive values 6 (t % 8) .. of row t / 8:
  // atoms 2 (t % 8) and 2 (t % 8) + 1, three components each.  Value (atom a, row r = 3 row + comp)
  // goes to fragment position ((a / 4) * 12 + r / 8) * 32 + (r % 8) * 4 + a % 4
  //   = 384 (a / 4) + 4 r + a % 4 = dst + 4 comp + (a odd),
  // so the whole role of a thread is three numbers (round 2: six offsets, six destinations and
  // twelve guards per thread pushed the kernel over its 128 registers and the compiler recomputed
  // them in the main loop -- 82 % of the executed instructions were not DMMA).
  static_assert(MMA_KA == 16 && MULTI_THREADS == 256 && MULTI_TF == 32 && MULTI_TG == 32, "loader shape");
  constexpr int NQ = 6;
  float xr[NQ];
  double yr[NQ];
  const int l_row = tid >> 3, l_atom = 2 * (tid & 7);
  const float* x_src = pos + ((f0 + l_row) * num_atoms + first_atom + l_atom) * 3;
  const double* y_src = ref + ((int64_t)(g0 + l_row) * n + l_atom) * 3;
  const int dst = 384 * (l_atom >> 2) + 12 * l_row + (l_atom & 3);
  // an invalid row (frame or group beyond the end) is encoded as an atom that is never in range
  const int a_x = f0 + l_row < num_frames ? l_atom : (1 << 30);
  const int a_y = g0 + l_row < num_groups ? l_atom : (1 << 30);
  auto load_regs = [&](int a0) {
    const float* xs = x_src + (int64_t)a0 * 3;
    const double* ys = y_src + (int64_t)a0 * 3;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      xr[q] = a0 + a_x + q / 3 < a_end ? __ldg(xs + q) : 0.f;
      yr[q] = a0 + a_y + q / 3 < a_end ? __ldg(ys + q) : 0.0;
    }
  };
  auto store_regs = [&](int buf) {
    double* as = &As[buf][0][0][0] + dst;
    double* bs = &Bs[buf][0][0][0] + dst;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      as[4 * (q % 3) + q / 3] = (double)xr[q];
      bs[4 * (q % 3) + q / 3] = yr[q];
    }
  };

  int buf = 0;
  if (a_begin < a_end) {
    load_regs(a_begin);
    store_regs(0);
  }
  __syncthreads();
  for (int a0 = a_begin; a0 < a_end; a0 += MMA_KA) {
    const bool more = a0 + MMA_KA < a_end;
    if (more) load_regs(a0 + MMA_KA);
#pragma unroll
    for (int ks = 0; ks < MMA_KA / 4; ++ks) {
      double a[3], b[6];
#pragma unroll
      for (int i = 0; i < 3; ++i) a[i] = As[buf][ks][3 * wm + i][lane];
#pragma unroll
      for (int j = 0; j < 6; ++j) b[j] = Bs[buf][ks][6 * wn + j][lane];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) dmma(acc[i][j], a[i], b[j]);
    }
    if (more) store_regs(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  // C[row = lane / 4][col = 2 (lane % 4) + h] of tile (3 wm + i, 6 wn + j)
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int r = 8 * (3 * wm + i) + (lane >> 2);
    const int64_t f = f0 + r / 3;
    const int c = r % 3;
    if (f >= num_frames) continue;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int q = 8 * (6 * wn + j) + 2 * (lane & 3) + h;
        const int g = g0 + q / 3, d = q % 3;
        if (g >= num_groups) continue;
        partials[((size_t)f * ns + (size_t)g * SK + sk) * 13 + 4 + 3 * c + d] = acc[i][j][h];
      }
    }
  }
}

// gradient: M = (frame, c) 96 rows, N = 64 atoms, K = (group, d) in stages of 4 groups (12 = 3 k-steps)
//   g[f][a][c] = F_f (x[f][a][c] - c_f[c]) - sum_(g,d) W[f][g][3c+d] y^[g][a][d]
// grid = (frame tiles, atom tiles); dynamic shared memory MULTI_GRAD_MMA_SMEM; two CTAs per SM, so
// that the prologue (x tile, frame constants) and the epilogue (coalesced rows of the gradient) of
// one CTA run under the tensor-core loop of the other.
// ---- covariance, three buffers and no block-wide barrier -------------------------------------------
// The kernel above ends every stage (16 atoms: 72 DMMA per warp) in __syncthreads(): the warps of a
// CTA meet, the DMMA pipes of their SM quarters wait for the slowest of them, and with two CTAs per
// SM nobody else is there to fill in (ncu: the pipe is 79 % busy, top stall math_pipe_throttle --
// when a warp wants the pipe it is busy, but one fifth of the time nobody wants it).  Same loader,
// same fragment order, same summation order; but three stage buffers and two mbarriers per buffer
// arrived on *one stage ahead of the wait*: at the top of stage s a thread stores the values of
// stage s + 1 (loaded during stage s - 1) and arrives on full[s + 1], loads stage s + 2 into
// registers, and only then waits for full[s] -- which everybody arrived on a whole stage ago.  The
// buffer of stage s + 1 was last read in stage s - 2: empty[] is arrived on after the reads and
// waited on before the stores, again a stage apart.
constexpr int COV3_BUFS = 3;
constexpr int COV3_STAGE = (MMA_KA / 4) * (MMA_ROWS / 8) * 32;  // doubles per operand and buffer
constexpr size_t MULTI_COV3_SMEM = 2 * COV3_BUFS * COV3_STAGE * sizeof(double) + 2 * COV3_BUFS * sizeof(uint64_t);

__global__ void __launch_bounds__(MULTI_THREADS, 2)
    rmsd_multi_cov_ring_kernel(const float* __restrict__ pos, int64_t num_atoms, int first_atom, int n,
                               const double* __restrict__ ref /* [G][n][3], centred */, int num_groups,
                               int64_t num_frames, int SK, double* __restrict__ partials) {
  extern __shared__ __align__(16) double multi_smem[];
  double* As = multi_smem;                       // [COV3_BUFS][k-step][tile][32]
  double* Bs = As + COV3_BUFS * COV3_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(Bs + COV3_BUFS * COV3_STAGE);
  uint64_t* empty = full + COV3_BUFS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int64_t f0 = (int64_t)blockIdx.x * MULTI_TF;
  const int g0 = blockIdx.y * MULTI_TG;
  const int sk = blockIdx.z;
  const int ns = num_groups * SK;
  const int chunks_total = (n + MMA_KA - 1) / MMA_KA;
  const int chunks_per = (chunks_total + SK - 1) / SK;
  const int a_begin = min(n, sk * chunks_per * MMA_KA);
  const int a_end = min(n, (sk + 1) * chunks_per * MMA_KA);
  const int stages = (a_end - a_begin + MMA_KA - 1) / MMA_KA;

  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < COV3_BUFS; ++k) {
      mbar_init(&full[k], MULTI_THREADS / 32);
      mbar_init(&empty[k], MULTI_THREADS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  double acc[3][6][2];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // (loader roles: see rmsd_multi_cov_mma_kernel)
  constexpr int NQ = 6;
  float xr[NQ];
  double yr[NQ];
  const int l_row = tid >> 3, l_atom = 2 * (tid & 7);
  const float* x_src = pos + ((f0 + l_row) * num_atoms + first_atom + l_atom) * 3;
  const double* y_src = ref + ((int64_t)(g0 + l_row) * n + l_atom) * 3;
  const int dst = 384 * (l_atom >> 2) + 12 * l_row + (l_atom & 3);
  const int a_x = f0 + l_row < num_frames ? l_atom : (1 << 30);
  const int a_y = g0 + l_row < num_groups ? l_atom : (1 << 30);
  auto load_regs = [&](int a0) {
    const float* xs = x_src + (int64_t)a0 * 3;
    const double* ys = y_src + (int64_t)a0 * 3;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      xr[q] = a0 + a_x + q / 3 < a_end ? __ldg(xs + q) : 0.f;
      yr[q] = a0 + a_y + q / 3 < a_end ? __ldg(ys + q) : 0.0;
    }
  };
  // stage s -> buffer s % 3; its k-th use (k = s / 3) completes phase k of both barriers
  auto store_stage = [&](int s) {
    const int buf = s % COV3_BUFS;
    if (s >= COV3_BUFS) mbar_wait(&empty[buf], (uint32_t)(s / COV3_BUFS - 1) & 1u);  // reads of stage s - 3
    double* as = As + buf * COV3_STAGE + dst;
    double* bs = Bs + buf * COV3_STAGE + dst;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      as[4 * (q % 3) + q / 3] = (double)xr[q];
      bs[4 * (q % 3) + q / 3] = yr[q];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&full[buf]);
  };

  if (stages > 0) {
    load_regs(a_begin);
    store_stage(0);
    if (stages > 1) load_regs(a_begin + MMA_KA);
  }
  for (int s = 0; s < stages; ++s) {
    if (s + 1 < stages) store_stage(s + 1);
    if (s + 2 < stages) load_regs(a_begin + (s + 2) * MMA_KA);
    const int buf = s % COV3_BUFS;
    mbar_wait(&full[buf], (uint32_t)(s / COV3_BUFS) & 1u);
    const double* as = As + buf * COV3_STAGE;
    const double* bs = Bs + buf * COV3_STAGE;
#pragma unroll
    for (int ks = 0; ks < MMA_KA / 4; ++ks) {
      double a[3], b[6];
#pragma unroll
      for (int i = 0; i < 3; ++i) a[i] = as[((ks * (MMA_ROWS / 8) + 3 * wm + i) << 5) + lane];
#pragma unroll
      for (int j = 0; j < 6; ++j) b[j] = bs[((ks * (MMA_ROWS / 8) + 6 * wn + j) << 5) + lane];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) dmma(acc[i][j], a[i], b[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[buf]);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int r = 8 * (3 * wm + i) + (lane >> 2);
    const int64_t f = f0 + r / 3;
    const int c = r % 3;
    if (f >= num_frames) continue;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int q = 8 * (6 * wn + j) + 2 * (lane & 3) + h;
        const int g = g0 + q / 3, d = q % 3;
        if (g >= num_groups) continue;
        partials[((size_t)f * ns + (size_t)g * SK + sk) * 13 + 4 + 3 * c + d] = acc[i][j][h];
      }
    }
  }
}

constexpr int MMA_TA = 64;                     // atoms per CTA tile
constexpr int MMA_KG = 4;                      // groups per stage
constexpr int MMA_GSTEPS = MMA_KG * 3 / 4;     // k-steps per stage
constexpr int MMA_A_STAGE = MMA_GSTEPS * (MMA_ROWS / 8) * 32;    // doubles
constexpr int MMA_B_STAGE = MMA_GSTEPS * (MMA_TA / 8) * 32;
constexpr int MMA_X_PITCH = MMA_TA * 3 + 4;    // floats per frame row of the x / output tile
constexpr int MMA_STAGES = 2;                  // double buffer filled through registers
constexpr size_t MULTI_GRAD_MMA_SMEM = MMA_STAGES * sizeof(double) * (MMA_A_STAGE + MMA_B_STAGE) +
                                       sizeof(float) * MULTI_TF * MMA_X_PITCH + sizeof(double) * MULTI_TF * 4;

__global__ void __launch_bounds__(MULTI_THREADS, 2)
    rmsd_multi_gradient_mma_kernel(const float* __restrict__ pos, int64_t num_atoms, int first_atom,
                                   int n, const double* __restrict__ ref, int num_groups,
                                   int64_t num_frames, const double* __restrict__ weights,
                                   const double* __restrict__ centroid /* [frame][nseg][3], seg 0 */,
                                   int nseg, float* __restrict__ grad, int accumulate) {
  extern __shared__ __align__(16) double multi_smem[];
  double* stage = multi_smem;                                     // [MMA_STAGES][A stage | B stage]
  float* xs = reinterpret_cast<float*>(stage + MMA_STAGES * (MMA_A_STAGE + MMA_B_STAGE));  // [32][MMA_X_PITCH]
  double* fc = reinterpret_cast<double*>(xs + MULTI_TF * MMA_X_PITCH);           // [32][4]: c_f, F_f
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;  // row tiles 3 wm .. + 2, atom tiles 4 wn .. 4 wn + 3
  const int64_t f0 = (int64_t)blockIdx.x * MULTI_TF;
  const int a0 = blockIdx.y * MMA_TA;
  const int atoms_here = min(MMA_TA, n - a0);

  // x tile (coalesced rows) and the frame constants; consumed by the epilogue
  for (int e = tid; e < MULTI_TF * MMA_TA * 3; e += MULTI_THREADS) {
    const int fl = e / (MMA_TA * 3), rest = e - fl * (MMA_TA * 3);
    float v = 0.f;
    if (f0 + fl < num_frames && rest < atoms_here * 3)
      v = pos[((f0 + fl) * num_atoms + first_atom + a0) * 3 + rest];
    xs[fl * MMA_X_PITCH + rest] = v;
  }
  {
    // F_f = sum_g factor[f][g]: 8 threads per frame, each a strided eighth of the groups
    const int fl = tid >> 3, part = tid & 7;
    const int64_t f = f0 + fl;
    double sum = 0.0;
    if (f < num_frames)
      for (int g = part; g < num_groups; g += 8) sum += __ldg(weights + (f * num_groups + g) * 10 + 9);
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
    if (part < 4) {
      double v = sum;
      if (part < 3) v = f < num_frames ? centroid[(size_t)f * nseg * 3 + part] : 0.0;
      fc[4 * fl + part] = v;
    }
  }

  // loaders (roles fixed per thread).  A stage: 32 frames x 4 groups x 9 values; B stage: 4 groups
  // x 64 atoms x 3; fragment positions as in the covariance kernel with k = 3 * group + d.
  constexpr int NA = (MULTI_TF * MMA_KG * 9 + MULTI_THREADS - 1) / MULTI_THREADS;  // 5
  constexpr int NB = MMA_KG * MMA_TA * 3 / MULTI_THREADS;                          // 3
  static_assert(MMA_KG * MMA_TA * 3 % MULTI_THREADS == 0, "B loader shape");
  // (8-byte cp.async copies straight to the fragment positions were tried instead of the register
  // staging below: 1.24 ms against 1.08 ms per 1024 frames -- one LSU transaction per double)
  int a_dst[NA], b_dst[NB];
  int a_src[NA], b_src[NB];  // offsets for group 0 of the stage (32 bits are plenty); < 0: zero
  int a_gl[NA], b_gl[NB];
  const double* w_base = weights + f0 * num_groups * 10;
  const double* y_base = ref + (int64_t)a0 * 3;
#pragma unroll
  for (int q = 0; q < NA; ++q) {
    const int e = tid + q * MULTI_THREADS;
    const int fl = e / (MMA_KG * 9), rest = e - fl * (MMA_KG * 9);
    const int gl = rest / 9, v = rest - gl * 9, c = v / 3, d = v - 3 * c;
    const int r = 3 * fl + c, kk = 3 * gl + d;
    const bool ok = e < MULTI_TF * MMA_KG * 9 && f0 + fl < num_frames;
    a_dst[q] = e < MULTI_TF * MMA_KG * 9
                   ? ((((kk >> 2) * (MMA_ROWS / 8) + (r >> 3)) << 5) + ((r & 7) << 2) + (kk & 3))
                   : -1;
    a_src[q] = ok ? (fl * num_groups + gl) * 10 + v : -1;
    a_gl[q] = gl;
  }
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int e = tid + q * MULTI_THREADS;
    const int gl = e / (MMA_TA * 3), rest = e - gl * (MMA_TA * 3);
    const int al = rest / 3, d = rest - 3 * al, kk = 3 * gl + d;
    b_dst[q] = (((kk >> 2) * (MMA_TA / 8) + (al >> 3)) << 5) + ((al & 7) << 2) + (kk & 3);
    b_src[q] = al < atoms_here ? (gl * n + al) * 3 + d : -1;
    b_gl[q] = gl;
  }
  double ar[NA], br[NB];
  auto load_regs = [&](int g0) {
#pragma unroll
    for (int q = 0; q < NA; ++q)
      ar[q] = a_src[q] >= 0 && g0 + a_gl[q] < num_groups ? __ldg(w_base + a_src[q] + (int64_t)g0 * 10) : 0.0;
#pragma unroll
    for (int q = 0; q < NB; ++q)
      br[q] = b_src[q] >= 0 && g0 + b_gl[q] < num_groups ? __ldg(y_base + b_src[q] + (int64_t)g0 * n * 3) : 0.0;
  };
  auto store_regs = [&](int buf) {
    double* as = stage + (size_t)buf * (MMA_A_STAGE + MMA_B_STAGE);
    double* bs = as + MMA_A_STAGE;
#pragma unroll
    for (int q = 0; q < NA; ++q)
      if (a_dst[q] >= 0) as[a_dst[q]] = ar[q];
#pragma unroll
    for (int q = 0; q < NB; ++q) bs[b_dst[q]] = br[q];
  };

  double acc[3][4][2];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  load_regs(0);
  store_regs(0);
  __syncthreads();
  int buf = 0;
  for (int g0 = 0; g0 < num_groups; g0 += MMA_KG) {
    const bool more = g0 + MMA_KG < num_groups;
    if (more) load_regs(g0 + MMA_KG);
    const double* as = stage + (size_t)buf * (MMA_A_STAGE + MMA_B_STAGE);
    const double* bs = as + MMA_A_STAGE;
#pragma unroll
    for (int ks = 0; ks < MMA_GSTEPS; ++ks) {
      double a[3], b[4];
#pragma unroll
      for (int i = 0; i < 3; ++i) a[i] = as[((ks * (MMA_ROWS / 8) + 3 * wm + i) << 5) + lane];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = bs[((ks * (MMA_TA / 8) + 4 * wn + j) << 5) + lane];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j], a[i], b[j]);
    }
    if (more) store_regs(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  // epilogue: v = F (x - c) - acc in double, written over x in the shared tile, then coalesced rows
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int r = 8 * (3 * wm + i) + (lane >> 2);
    const int fl = r / 3, c = r - 3 * fl;
    const double cf = fc[4 * fl + c], big_f = fc[4 * fl + 3];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int al = 8 * (4 * wn + j) + 2 * (lane & 3) + h;
        float* slot = xs + fl * MMA_X_PITCH + al * 3 + c;
        *slot = (float)(big_f * ((double)*slot - cf) - acc[i][j][h]);
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < MULTI_TF * MMA_TA * 3; e += MULTI_THREADS) {
    const int fl = e / (MMA_TA * 3), rest = e - fl * (MMA_TA * 3);
    if (f0 + fl >= num_frames || rest >= atoms_here * 3) continue;
    float* out = grad + ((f0 + fl) * num_atoms + first_atom + a0) * 3 + rest;
    const float v = xs[fl * MMA_X_PITCH + rest];
    *out = accumulate ? *out + v : v;
  }
}

// ---- gradient, operands pre-packed in fragment order and fed by TMA ---------------------------------
// The register-staged kernel above reaches 60 % of the DMMA issue rate: K = 3 G is short, a stage
// holds 36 DMMA per warp, and the global loads of the next stage (scattered 8-byte loads with
// per-thread index tables, 24 registers) come back later than that.  Both operands are constants
// of something larger than a CTA -- the reference of the whole spec, the W records of a batch --
// so they are written ONCE in the order the fragments are read:
//   ref_pack [atom tile][stage][k-step][8 column tiles][32]   (host, at upload; zero padded)
//   w_pack   [frame tile][stage][k-step][12 row tiles][32]    (rmsd_multi_weights_pack_kernel)
// A stage of either is one contiguous block: one elected thread moves it with cp.async.bulk into a
// ring of PK_STAGES slots (full / empty mbarriers), no thread spends a register or an instruction
// on operand loads, and the main loop is LDS.64 + DMMA only.
#ifndef PK_KG_VALUE
#define PK_KG_VALUE 4
#endif
#ifndef PK_STAGES_VALUE
#define PK_STAGES_VALUE 4
#endif
constexpr int PK_KG = PK_KG_VALUE;            // groups per stage (3 PK_KG must be a multiple of 4)
constexpr int PK_STAGES = PK_STAGES_VALUE;    // ring slots
constexpr int PK_KS = PK_KG * 3 / 4;          // k-steps per stage
constexpr int PK_A_STAGE = PK_KS * (MMA_ROWS / 8) * 32;  // doubles
constexpr int PK_B_STAGE = PK_KS * (MMA_TA / 8) * 32;
static_assert(PK_KG * 3 % 4 == 0, "a stage is a whole number of k-steps");
constexpr size_t MULTI_GRAD_PK_SMEM = PK_STAGES * sizeof(double) * (PK_A_STAGE + PK_B_STAGE) +
                                      sizeof(float) * MULTI_TF * MMA_X_PITCH + sizeof(double) * MULTI_TF * 4 +
                                      2 * PK_STAGES * sizeof(uint64_t);

inline int pk_stages(int num_groups) { return (num_groups + PK_KG - 1) / PK_KG; }
// position of B[k = 3 gl + d][column = al] of a stage / A[row = 3 fl + c][k] in its block
__host__ __device__ inline int pk_b_index(int gl, int d, int al) {
  const int kk = 3 * gl + d;
  return (((kk >> 2) * (MMA_TA / 8) + (al >> 3)) << 5) + ((al & 7) << 2) + (kk & 3);
}
__host__ __device__ inline int pk_a_index(int gl, int d, int fl, int c) {
  const int kk = 3 * gl + d, r = 3 * fl + c;
  return (((kk >> 2) * (MMA_ROWS / 8) + (r >> 3)) << 5) + ((r & 7) << 2) + (kk & 3);
}

// host: the centred references [G][n][3] -> ref_pack
inline void pk_pack_reference(const double* ref, int num_groups, int n, std::vector<double>& out) {
  const int tiles = (n + MMA_TA - 1) / MMA_TA, ns = pk_stages(num_groups);
  out.assign((size_t)tiles * ns * PK_B_STAGE, 0.0);
  for (int g = 0; g < num_groups; ++g) {
    const int s = g / PK_KG, gl = g % PK_KG;
    for (int a = 0; a < n; ++a) {
      double* block = out.data() + ((size_t)(a / MMA_TA) * ns + s) * PK_B_STAGE;
      for (int d = 0; d < 3; ++d) block[pk_b_index(gl, d, a % MMA_TA)] = ref[((size_t)g * n + a) * 3 + d];
    }
  }
}

// W[f][g][3c+d] = factor[f][g] U^T[f][g][3c+d] -> w_pack; one thread per (padded frame, padded group)
template <typename Solution>
__global__ void rmsd_multi_weights_pack_kernel(const Solution* __restrict__ solution,
                                               const double* __restrict__ factor, int64_t num_frames,
                                               int num_groups, double* __restrict__ w_pack) {
  const int ns = (num_groups + PK_KG - 1) / PK_KG, gp = ns * PK_KG;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t frames_pad = (num_frames + MULTI_TF - 1) / MULTI_TF * MULTI_TF;
  if (t >= frames_pad * gp) return;
  const int64_t f = t / gp;
  const int g = (int)(t - f * gp);
  const bool real = f < num_frames && g < num_groups;
  const double fac = real ? factor[f * num_groups + g] : 0.0;
  const double* ut = solution[real ? f * num_groups + g : 0].ut;
  double* block = w_pack + ((f / MULTI_TF) * ns + g / PK_KG) * PK_A_STAGE;
  const int fl = (int)(f % MULTI_TF), gl = g % PK_KG;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int d = 0; d < 3; ++d) block[pk_a_index(gl, d, fl, c)] = real ? fac * ut[3 * c + d] : 0.0;
}

// grid = (frame tiles, atom tiles), MULTI_THREADS threads, dynamic shared memory MULTI_GRAD_PK_SMEM.
// BULK (positions and gradient 16-byte aligned, first_atom, N and n multiples of 4): the x tile comes
// in and the gradient tile goes out as one bulk copy per frame row (768 bytes), issued by the lanes
// of warp 0 -- in the scalar version the index arithmetic of those two loops was 70 % of the
// kernel's instructions (ncu source page, round 2).
template <bool BULK>
__global__ void __launch_bounds__(MULTI_THREADS, 2)
    rmsd_multi_gradient_pk_kernel(const float* __restrict__ pos, int64_t num_atoms, int first_atom, int n,
                                  const double* __restrict__ ref_pack, int num_groups, int64_t num_frames,
                                  const double* __restrict__ w_pack, const double* __restrict__ factor,
                                  const double* __restrict__ centroid /* [frame][nseg][3], seg 0 */,
                                  int nseg, float* __restrict__ grad, int accumulate) {
  extern __shared__ __align__(16) double multi_smem[];
  double* ring = multi_smem;  // [PK_STAGES][A stage | B stage]
  float* xs = reinterpret_cast<float*>(ring + PK_STAGES * (PK_A_STAGE + PK_B_STAGE));  // [32][MMA_X_PITCH]
  double* fc = reinterpret_cast<double*>(xs + MULTI_TF * MMA_X_PITCH);                  // [32][4]: c_f, F_f
  uint64_t* full = reinterpret_cast<uint64_t*>(fc + MULTI_TF * 4);
  uint64_t* empty = full + PK_STAGES;
  __shared__ __align__(8) uint64_t xbar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int64_t f0 = (int64_t)blockIdx.x * MULTI_TF;
  const int a0 = blockIdx.y * MMA_TA;
  const int atoms_here = min(MMA_TA, n - a0);
  const int frames_here = (int)min((int64_t)MULTI_TF, num_frames - f0);
  const int ns = (num_groups + PK_KG - 1) / PK_KG;
  const double* a_src = w_pack + (size_t)blockIdx.x * ns * PK_A_STAGE;
  const double* b_src = ref_pack + (size_t)blockIdx.y * ns * PK_B_STAGE;
  constexpr uint32_t STAGE_BYTES = (PK_A_STAGE + PK_B_STAGE) * sizeof(double);

  auto issue = [&](int s) {  // one thread: stage s -> slot s % PK_STAGES
    const int slot = s % PK_STAGES;
    double* dst = ring + (size_t)slot * (PK_A_STAGE + PK_B_STAGE);
    mbar_expect_tx(&full[slot], STAGE_BYTES);
    tma_load_1d(dst, a_src + (size_t)s * PK_A_STAGE, PK_A_STAGE * sizeof(double), &full[slot]);
    tma_load_1d(dst + PK_A_STAGE, b_src + (size_t)s * PK_B_STAGE, PK_B_STAGE * sizeof(double), &full[slot]);
  };
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < PK_STAGES; ++k) {
      mbar_init(&full[k], 1);
      mbar_init(&empty[k], MULTI_THREADS / 32);
    }
    if (BULK) mbar_init(&xbar, 1);
    mbar_fence_init();
    for (int s = 0; s < min(ns, PK_STAGES); ++s) issue(s);
    if (BULK) mbar_expect_tx(&xbar, (uint32_t)(frames_here * atoms_here * 12));
  }

  // x tile and the frame constants, consumed by the epilogue -- under the first copies
  if (BULK) {
    // (rows of frames past the batch and columns past the group stay undefined: never stored)
    __syncthreads();  // xbar initialised and armed
    if (warp == 0 && lane < frames_here)
      tma_load_1d(xs + lane * MMA_X_PITCH, pos + ((f0 + lane) * num_atoms + first_atom + a0) * 3,
                  (uint32_t)(atoms_here * 12), &xbar);
  } else {
    for (int e = tid; e < MULTI_TF * MMA_TA * 3; e += MULTI_THREADS) {
      const int fl = e / (MMA_TA * 3), rest = e - fl * (MMA_TA * 3);
      float v = 0.f;
      if (f0 + fl < num_frames && rest < atoms_here * 3)
        v = pos[((f0 + fl) * num_atoms + first_atom + a0) * 3 + rest];
      xs[fl * MMA_X_PITCH + rest] = v;
    }
  }
  {
    // F_f = sum_g factor[f][g]: 8 threads per frame, each a strided eighth of the groups
    const int fl = tid >> 3, part = tid & 7;
    const int64_t f = f0 + fl;
    double sum = 0.0;
    if (f < num_frames)
      for (int g = part; g < num_groups; g += 8) sum += __ldg(factor + f * num_groups + g);
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
    if (part < 4) {
      double v = sum;
      if (part < 3) v = f < num_frames ? centroid[(size_t)f * nseg * 3 + part] : 0.0;
      fc[4 * fl + part] = v;
    }
  }
  __syncthreads();  // barriers initialised, x tile and constants in place

  double acc[3][4][2];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int s = 0; s < ns; ++s) {
    const int slot = s % PK_STAGES;
    const uint32_t parity = (uint32_t)(s / PK_STAGES) & 1u;
    mbar_wait(&full[slot], parity);
    const double* as = ring + (size_t)slot * (PK_A_STAGE + PK_B_STAGE);
    const double* bs = as + PK_A_STAGE;
#pragma unroll
    for (int ks = 0; ks < PK_KS; ++ks) {
      double a[3], b[4];
#pragma unroll
      for (int i = 0; i < 3; ++i) a[i] = as[((ks * (MMA_ROWS / 8) + 3 * wm + i) << 5) + lane];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = bs[((ks * (MMA_TA / 8) + 4 * wn + j) << 5) + lane];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j], a[i], b[j]);
    }
    if (s + PK_STAGES < ns) {  // the slot is refilled once every warp has read it
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
      if (tid == 0) {
        mbar_wait(&empty[slot], parity);
        issue(s + PK_STAGES);
      }
    }
  }
  // epilogue: v = F (x - c) - acc in double, written over x in the shared tile, then whole rows
  if (BULK) mbar_wait(&xbar, 0);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int r = 8 * (3 * wm + i) + (lane >> 2);
    const int fl = r / 3, c = r - 3 * fl;
    const double cf = fc[4 * fl + c], big_f = fc[4 * fl + 3];
    float* row = xs + fl * MMA_X_PITCH + (8 * 4 * wn + 2 * (lane & 3)) * 3 + c;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float* slot = row + (8 * j + h) * 3;
        *slot = (float)(big_f * ((double)*slot - cf) - acc[i][j][h]);
      }
    }
  }
  if (BULK) {
    fence_proxy_async_smem();  // the generic-proxy writes above are read by the bulk stores below
    __syncthreads();
    if (warp == 0) {
      if (lane < frames_here) {
        float* out = grad + ((f0 + lane) * num_atoms + first_atom + a0) * 3;
        if (accumulate)
          tma_reduce_add_f32_1d(out, xs + lane * MMA_X_PITCH, (uint32_t)(atoms_here * 12));
        else
          tma_store_1d(out, xs + lane * MMA_X_PITCH, (uint32_t)(atoms_here * 12), l2_policy_evict_first());
      }
      tma_commit_group();
      tma_wait_group_read<0>();  // the tile must outlive the reads
    }
    return;
  }
  __syncthreads();
  for (int e = tid; e < MULTI_TF * MMA_TA * 3; e += MULTI_THREADS) {
    const int fl = e / (MMA_TA * 3), rest = e - fl * (MMA_TA * 3);
    if (f0 + fl >= num_frames || rest >= atoms_here * 3) continue;
    float* out = grad + ((f0 + fl) * num_atoms + first_atom + a0) * 3 + rest;
    const float v = xs[fl * MMA_X_PITCH + rest];
    *out = accumulate ? *out + v : v;
  }
}

// ---- covariance, TMA-fed ----------------------------------------------------------------------------
// Same idea for the forward contraction.  The references are packed once, per tile of MULTI_TG groups
// and chunk of MMA_KA atoms, in B-fragment order (ref_pack_cov [group tile][chunk][k-step][12][32]);
// the frames cannot be packed (they are the input), but their rows can be copied as they are: 32 bulk
// copies of 16 atoms x 12 bytes per stage, one per lane of warp 0.  The A fragments are then read
// from the raw float rows and widened in registers (3 LDS.32 + 3 F2F per 18 DMMA).  MMA row tile
// 3 fb + c holds component c of frames 8 fb .. 8 fb + 7, which makes those loads conflict-free with
// a row pitch of 52 floats: bank = 20 (lane / 4) + 3 (lane % 4) + const (mod 32), all distinct.
// Needs first_atom, N and n multiples of 4 and aligned positions; otherwise the register-staged
// kernel above runs.  Same partial-sum records, same summation order per record.
constexpr int CPK_STAGES = 4;
constexpr int CPK_XP = 52;                                  // floats per frame row of a stage
constexpr int CPK_A_STAGE = MULTI_TF * CPK_XP;              // floats
constexpr int CPK_B_STAGE = (MMA_KA / 4) * (MMA_ROWS / 8) * 32;  // doubles
constexpr size_t MULTI_COV_PK_SMEM = CPK_STAGES * (sizeof(double) * CPK_B_STAGE + sizeof(float) * CPK_A_STAGE) +
                                     2 * CPK_STAGES * sizeof(uint64_t);

inline int cpk_chunks(int n) { return (n + MMA_KA - 1) / MMA_KA; }
// host: the centred references [G][n][3] -> ref_pack_cov
inline void cpk_pack_reference(const double* ref, int num_groups, int n, std::vector<double>& out) {
  const int tiles = (num_groups + MULTI_TG - 1) / MULTI_TG, chunks = cpk_chunks(n);
  out.assign((size_t)tiles * chunks * CPK_B_STAGE, 0.0);
  for (int g = 0; g < num_groups; ++g)
    for (int a = 0; a < n; ++a) {
      double* block = out.data() + ((size_t)(g / MULTI_TG) * chunks + a / MMA_KA) * CPK_B_STAGE;
      const int al = a % MMA_KA;
      for (int d = 0; d < 3; ++d) {
        const int q = 3 * (g % MULTI_TG) + d;
        block[(((al >> 2) * (MMA_ROWS / 8) + (q >> 3)) << 5) + ((q & 7) << 2) + (al & 3)] =
            ref[((size_t)g * n + a) * 3 + d];
      }
    }
}

// grid = (frame tiles, group tiles, SK), MULTI_THREADS threads, dynamic shared memory MULTI_COV_PK_SMEM
__global__ void __launch_bounds__(MULTI_THREADS, 2)
    rmsd_multi_cov_pk_kernel(const float* __restrict__ pos, int64_t num_atoms, int first_atom, int n,
                             const double* __restrict__ ref_pack_cov, int num_groups, int64_t num_frames,
                             int SK, double* __restrict__ partials) {
  extern __shared__ __align__(16) double multi_smem[];
  double* bring = multi_smem;                                              // [CPK_STAGES][CPK_B_STAGE]
  float* aring = reinterpret_cast<float*>(bring + CPK_STAGES * CPK_B_STAGE);  // [CPK_STAGES][32][CPK_XP]
  uint64_t* full = reinterpret_cast<uint64_t*>(aring + CPK_STAGES * CPK_A_STAGE);
  uint64_t* empty = full + CPK_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;  // frames 8 wm .. 8 wm + 7 (3 row tiles), column tiles 6 wn .. 6 wn + 5
  const int64_t f0 = (int64_t)blockIdx.x * MULTI_TF;
  const int g0 = blockIdx.y * MULTI_TG;
  const int sk = blockIdx.z;
  const int ns = num_groups * SK;
  const int frames_here = (int)min((int64_t)MULTI_TF, num_frames - f0);
  const int chunks_total = (n + MMA_KA - 1) / MMA_KA;
  const int chunks_per = (chunks_total + SK - 1) / SK;
  const int c_begin = min(chunks_total, sk * chunks_per);
  const int c_end = min(chunks_total, (sk + 1) * chunks_per);
  const int stages = c_end - c_begin;
  const float* x_row = pos + ((f0 + min(lane, frames_here - 1)) * num_atoms + first_atom) * 3;
  const double* b_src = ref_pack_cov + (size_t)blockIdx.y * chunks_total * CPK_B_STAGE;

  // warp 0: stage s -> slot s % CPK_STAGES (lane = frame row; lane 0 also moves the reference block)
  auto issue = [&](int s) {
    const int slot = s % CPK_STAGES;
    const int a = (c_begin + s) * MMA_KA;
    const uint32_t row_bytes = (uint32_t)(min(MMA_KA, n - a) * 12);
    if (lane == 0) mbar_expect_tx(&full[slot], (uint32_t)(CPK_B_STAGE * sizeof(double)) + row_bytes * frames_here);
    __syncwarp();
    if (lane == 0)
      tma_load_1d(bring + (size_t)slot * CPK_B_STAGE, b_src + (size_t)(c_begin + s) * CPK_B_STAGE,
                  CPK_B_STAGE * sizeof(double), &full[slot]);
    if (lane < frames_here)
      tma_load_1d(aring + (size_t)slot * CPK_A_STAGE + lane * CPK_XP, x_row + (int64_t)a * 3, row_bytes, &full[slot]);
  };
  // rows of frames past the batch and columns past the group are never copied: they must hold
  // finite numbers (they meet zeros of the packed reference or rows that are not stored)
  for (int e = tid; e < CPK_STAGES * CPK_A_STAGE; e += MULTI_THREADS) aring[e] = 0.f;
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < CPK_STAGES; ++k) {
      mbar_init(&full[k], 1);
      mbar_init(&empty[k], MULTI_THREADS / 32);
    }
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (warp == 0)
    for (int s = 0; s < min(stages, CPK_STAGES); ++s) issue(s);

  double acc[3][6][2];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int a_off = (8 * wm + (lane >> 2)) * CPK_XP + 3 * (lane & 3);
  for (int s = 0; s < stages; ++s) {
    const int slot = s % CPK_STAGES;
    const uint32_t parity = (uint32_t)(s / CPK_STAGES) & 1u;
    mbar_wait(&full[slot], parity);
    const float* as = aring + (size_t)slot * CPK_A_STAGE + a_off;
    const double* bs = bring + (size_t)slot * CPK_B_STAGE;
#pragma unroll
    for (int ks = 0; ks < MMA_KA / 4; ++ks) {
      double a[3], b[6];
#pragma unroll
      for (int i = 0; i < 3; ++i) a[i] = (double)as[12 * ks + i];
#pragma unroll
      for (int j = 0; j < 6; ++j) b[j] = bs[((ks * (MMA_ROWS / 8) + 6 * wn + j) << 5) + lane];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) dmma(acc[i][j], a[i], b[j]);
    }
    if (s + CPK_STAGES < stages) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
      if (warp == 0) {
        mbar_wait(&empty[slot], parity);
        issue(s + CPK_STAGES);
      }
    }
  }
  // C[row = lane / 4][col = 2 (lane % 4) + h] of tile (component i of frames 8 wm .., column tile 6 wn + j)
  const int64_t f = f0 + 8 * wm + (lane >> 2);
  if (f >= num_frames) return;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int q = 8 * (6 * wn + j) + 2 * (lane & 3) + h;
        const int g = g0 + q / 3, d = q % 3;
        if (g >= num_groups) continue;
        partials[((size_t)f * ns + (size_t)g * SK + sk) * 13 + 4 + 3 * i + d] = acc[i][j][h];
      }
    }
  }
}

constexpr size_t MULTI_GRAD_SMEM =
    2 * sizeof(double) * (MULTI_KG * MULTI_TF * 10 + MULTI_KG * MULTI_TA * 3);

}  // namespace
}  // namespace cvp
