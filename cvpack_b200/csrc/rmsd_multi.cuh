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
constexpr int MULTI_TA = 64;      // atoms per CTA tile (gradient)
constexpr int MULTI_KG = 16;      // groups per stage (gradient)

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
  float xr[NQ];
  double yr[NQ];
  auto load_regs = [&](int a0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int idx = tid + q * MULTI_THREADS;
      const int row = idx / ROW, rest = idx - row * ROW;  // row: frame / group of the tile
      const int a = a0 + rest / 3;
      const int64_t f = f0 + row;
      xr[q] = (f < num_frames && a < a_end)
                  ? __ldg(pos + (f * num_atoms + first_atom) * 3 + (int64_t)a0 * 3 + rest)
                  : 0.f;
      const int g = g0 + row;
      yr[q] = (g < num_groups && a < a_end)
                  ? __ldg(ref + ((int64_t)g * n + a0) * 3 + rest)
                  : 0.0;
    }
  };
  auto store_regs = [&](int buf) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int idx = tid + q * MULTI_THREADS;
      const int row = idx / ROW, rest = idx - row * ROW;
      const int a = rest / 3, c = rest - a * 3;
      Xs[buf][a][row][c] = (double)xr[q];
      Ys[buf][a][row][c] = yr[q];
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
// grid = (frame tiles, atom tiles).  `solution` / `factor` are indexed [frame][group].
template <typename Solution>
__global__ void __launch_bounds__(MULTI_THREADS, 2)
    rmsd_multi_gradient_kernel(const float* __restrict__ pos, int64_t num_atoms, int first_atom,
                               int n, const double* __restrict__ ref, int num_groups,
                               int64_t num_frames, const Solution* __restrict__ solution,
                               const double* __restrict__ factor,
                               const double* __restrict__ centroid /* [frame][nseg][3], seg 0 */,
                               int nseg, float* __restrict__ grad, int accumulate) {
  extern __shared__ __align__(16) double multi_smem[];
  double (*Ws)[MULTI_TF][10] = reinterpret_cast<double (*)[MULTI_TF][10]>(multi_smem);  // f U^T, f
  double (*Ys)[MULTI_TA][3] =
      reinterpret_cast<double (*)[MULTI_TA][3]>(multi_smem + MULTI_KG * MULTI_TF * 10);
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;  // frames 2ty, 2ty+1; atoms 4tx .. 4tx+3 of the tile
  const int64_t f0 = (int64_t)blockIdx.x * MULTI_TF;
  const int a0 = blockIdx.y * MULTI_TA;

  double acc[2][4][3];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[i][j][c] = 0.0;
  double fsum[2] = {0.0, 0.0};

  for (int g0 = 0; g0 < num_groups; g0 += MULTI_KG) {
    __syncthreads();
    // W tile: MULTI_KG groups x 32 frames, one (frame, group) pair per thread and pass; consecutive
    // threads take consecutive groups of a frame (contiguous 80-byte solution records)
    for (int idx = tid; idx < MULTI_KG * MULTI_TF; idx += MULTI_THREADS) {
      const int fl = idx / MULTI_KG, gl = idx - fl * MULTI_KG;
      const int g = g0 + gl;
      const int64_t f = f0 + fl;
      double w[10];
      if (g < num_groups && f < num_frames) {
        const double fac = factor[f * num_groups + g];
        const Solution& sol = solution[f * num_groups + g];
#pragma unroll
        for (int k = 0; k < 9; ++k) w[k] = fac * sol.ut[k];
        w[9] = fac;
      } else {
#pragma unroll
        for (int k = 0; k < 10; ++k) w[k] = 0.0;
      }
#pragma unroll
      for (int k = 0; k < 10; ++k) Ws[gl][fl][k] = w[k];
    }
    // Y tile: MULTI_KG groups x 64 atoms x 3 (192 contiguous doubles per group)
    for (int idx = tid; idx < MULTI_KG * MULTI_TA * 3; idx += MULTI_THREADS) {
      const int gl = idx / (MULTI_TA * 3), rest = idx - gl * (MULTI_TA * 3);
      const int g = g0 + gl;
      const int a = a0 + rest / 3;
      (&Ys[gl][0][0])[rest] =
          (g < num_groups && a < n) ? __ldg(ref + ((int64_t)g * n + a0) * 3 + rest) : 0.0;
    }
    __syncthreads();
#pragma unroll 2
    for (int gl = 0; gl < MULTI_KG; ++gl) {
      double w[2][10];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double2* pw = reinterpret_cast<const double2*>(&Ws[gl][2 * ty + i][0]);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const double2 v = pw[k];
          w[i][2 * k] = v.x;
          w[i][2 * k + 1] = v.y;
        }
        fsum[i] += w[i][9];
      }
      const double2* py = reinterpret_cast<const double2*>(&Ys[gl][4 * tx][0]);
      double y[12];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const double2 v = py[k];
        y[2 * k] = v.x;
        y[2 * k + 1] = v.y;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            acc[i][j][c] = fma(w[i][3 * c], y[3 * j],
                               fma(w[i][3 * c + 1], y[3 * j + 1],
                                   fma(w[i][3 * c + 2], y[3 * j + 2], acc[i][j][c])));
    }
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
    for (int j = 0; j < 4; ++j) {
      const int a = a0 + 4 * tx + j;
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

constexpr size_t MULTI_GRAD_SMEM =
    sizeof(double) * (MULTI_KG * MULTI_TF * 10 + MULTI_KG * MULTI_TA * 3);

}  // namespace
}  // namespace cvp
