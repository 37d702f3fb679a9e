// RMSD family, value + gradient, for a batch of frames: RMSD, CompositeRMSD, Helix/SheetRMSDContent
// and PathInRMSDSpace.
//
// Replaces openmm.RMSDForce (cvpack/rmsd.py:96), openmmcppforces.CompositeRMSDForce
// (composite_rmsd.py:135-137) and the CustomCVForce compositions over RMSD CVs
// (base_rmsd_content.py:44-79, base_path_cv.py:43-62).
//
// A *group* is one optimal-superposition problem.  Its entries (atom, reference row) are sorted by
// *body*; a (group, body) run is a *segment*, centred on its own centroid in both structures, all
// segments of a group sharing one rotation (plain RMSD: one segment per group).
//
//   R = sum_i x^_i (x) y^_i          (3x3 correlation; y^ pre-centred on the host, so x needs no
//                                     centring here: sum x^ (x) y^ = sum x (x) y^)
//   Gx = sum |x|^2 - sum_seg |sum_seg x|^2 / n_seg,  Gy = sum |y^|^2
//   F(R) Horn's 4x4 matrix, lambda_max, q -> U (rotates x^ onto y^)
//   msd = max(0, (Gx + Gy - 2 lambda_max)/n);  msd < 1e-20 -> rmsd = 0 with zero gradient
//   d rmsd/d x_i = (x_i - c_seg(i) - U^T y^_i) / (n rmsd)
//
// Stages (one stream, deterministic):
//   sums      one warp per (frame, slice of <= 1024 entries): 13 double sums
//   solve     one thread per (frame, group): reduce slices, Jacobi eigen-solve in double
//   combine   one thread per frame: CV value and d CV / d rmsd_g
//   gradient  one thread per (frame, active atom): sum over the atom's entries (CSR), single writer
#include "common.cuh"
#include "stream_fused.cuh"
#include "rmsd_multi.cuh"

namespace cvp {
namespace {

constexpr int RMSD_SLICE = 1024;
constexpr int RMSD_WARPS = 8;

struct RmsdSlice {
  int32_t begin, end;  // entry range
};

// ---- sums --------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(RMSD_WARPS * 32)
    rmsd_sums_kernel(const T* __restrict__ pos, int64_t num_atoms,
                     const int32_t* __restrict__ entry_atom, const double* __restrict__ ref,
                     const RmsdSlice* __restrict__ slices, int nslice, int64_t num_frames,
                     double* __restrict__ partials) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (w >= num_frames * nslice) return;
  const int64_t b = w / nslice;
  const int s = (int)(w - b * nslice);
  const RmsdSlice sl = slices[s];
  const T* frame = pos + b * num_atoms * 3;
  double v[13];
#pragma unroll
  for (int k = 0; k < 13; ++k) v[k] = 0.0;
  for (int e = sl.begin + lane; e < sl.end; e += 32) {
    const int64_t atom = entry_atom[e];
    const double x = frame[3 * atom], y = frame[3 * atom + 1], z = frame[3 * atom + 2];
    const double rx = ref[3 * (int64_t)e], ry = ref[3 * (int64_t)e + 1], rz = ref[3 * (int64_t)e + 2];
    v[0] += x;
    v[1] += y;
    v[2] += z;
    v[3] += x * x + y * y + z * z;
    v[4] += x * rx;
    v[5] += x * ry;
    v[6] += x * rz;
    v[7] += y * rx;
    v[8] += y * ry;
    v[9] += y * rz;
    v[10] += z * rx;
    v[11] += z * ry;
    v[12] += z * rz;
  }
#pragma unroll
  for (int k = 0; k < 13; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 13; ++k) partials[(size_t)w * 13 + k] = v[k];
  }
}

// ---- streaming fast path: one group = a contiguous, 16-byte aligned run of atoms (FP32) ----------
// The common large case (RMSD of a whole protein, BASELINE config 3) needs no index gather: the
// frame slice is read as float4 (4 atoms = 3 x float4), the centred reference as float4 from L2.
constexpr int RMSD_FAST_THREADS = 256;
constexpr int RMSD_FAST_SLICE = 4096;  // atoms per block

__device__ __forceinline__ void unpack4(const float4& a0, const float4& a1, const float4& a2,
                                        float (&x)[4], float (&y)[4], float (&z)[4]) {
  x[0] = a0.x; y[0] = a0.y; z[0] = a0.z;
  x[1] = a0.w; y[1] = a1.x; z[1] = a1.y;
  x[2] = a1.z; y[2] = a1.w; z[2] = a2.x;
  x[3] = a2.y; y[3] = a2.z; z[3] = a2.w;
}

__global__ void __launch_bounds__(RMSD_FAST_THREADS)
    rmsd_sums_fast_kernel(const float* __restrict__ pos, int64_t num_atoms, int first_atom, int n,
                          const double* __restrict__ ref, int nslice, double* __restrict__ partials) {
  __shared__ double scratch[13 * (RMSD_FAST_THREADS / 32)];
  const int64_t b = blockIdx.x / nslice;
  const int s = blockIdx.x - (int)b * nslice;
  const float4* px = reinterpret_cast<const float4*>(pos + (b * num_atoms + first_atom) * 3);
  // the centred reference in double: rounded to float it is no longer centred, and the
  // single-pass covariance sum x (x) y^ (x not centred) picks up c (x) sum y^ -- an error of 1e-9 in
  // the mean square deviation, i.e. of 5e-5 nm in an RMSD close to zero
  const double2* py = reinterpret_cast<const double2*>(ref);
  const int g_begin = s * (RMSD_FAST_SLICE / 4);
  const int g_end = min(n / 4, g_begin + RMSD_FAST_SLICE / 4);
  double v[13];
#pragma unroll
  for (int k = 0; k < 13; ++k) v[k] = 0.0;
  for (int g = g_begin + threadIdx.x; g < g_end; g += RMSD_FAST_THREADS) {
    const float4 a0 = __ldg(px + 3 * g), a1 = __ldg(px + 3 * g + 1), a2 = __ldg(px + 3 * g + 2);
    double2 r[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) r[k] = __ldg(py + 6 * g + k);
    const double rx[4] = {r[0].x, r[1].y, r[3].x, r[4].y};
    const double ry[4] = {r[0].y, r[2].x, r[3].y, r[5].x};
    const double rz[4] = {r[1].x, r[2].y, r[4].x, r[5].y};
    float x[4], y[4], z[4];
    unpack4(a0, a1, a2, x, y, z);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double xd = x[k], yd = y[k], zd = z[k], rxd = rx[k], ryd = ry[k], rzd = rz[k];
      v[0] += xd;
      v[1] += yd;
      v[2] += zd;
      v[3] += xd * xd + yd * yd + zd * zd;
      v[4] += xd * rxd;
      v[5] += xd * ryd;
      v[6] += xd * rzd;
      v[7] += yd * rxd;
      v[8] += yd * ryd;
      v[9] += yd * rzd;
      v[10] += zd * rxd;
      v[11] += zd * ryd;
      v[12] += zd * rzd;
    }
  }
  block_sum<13>(v, scratch);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 13; ++k) partials[((size_t)b * nslice + s) * 13 + k] = v[k];
  }
}

// ---- solve -------------------------------------------------------------------------------------
// Largest eigenpair of a symmetric 4x4 matrix by cyclic Jacobi rotations (double precision; plays
// the role of the JAMA eigen-decomposition in OpenMM's reference RMSD kernel).
__device__ inline void jacobi4_max(double (&a)[4][4], double& lambda, double (&q)[4]) {
  double v[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) v[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 24; ++sweep) {
    double off = 0.0, diag = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      diag += a[i][i] * a[i][i];
#pragma unroll
      for (int j = i + 1; j < 4; ++j) off += a[i][j] * a[i][j];
    }
    if (off <= 1e-32 * diag || off == 0.0) break;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
#pragma unroll
      for (int r = p + 1; r < 4; ++r) {
        const double apr = a[p][r];
        if (apr == 0.0) continue;
        const double theta = (a[r][r] - a[p][p]) / (2.0 * apr);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double akp = a[k][p], akr = a[k][r];
          a[k][p] = c * akp - s * akr;
          a[k][r] = s * akp + c * akr;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double apk = a[p][k], ark = a[r][k];
          a[p][k] = c * apk - s * ark;
          a[r][k] = s * apk + c * ark;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double vkp = v[k][p], vkr = v[k][r];
          v[k][p] = c * vkp - s * vkr;
          v[k][r] = s * vkp + c * vkr;
        }
      }
    }
  }
  int best = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (a[i][i] > a[best][best]) best = i;
  lambda = a[best][best];
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = v[k][best];
}

// Largest eigenpair of Horn's matrix the fast way (Theobald's QCP idea): Newton iteration on the
// characteristic polynomial from lambda0 = (Gx+Gy)/2 >= lambda_max (monotone convergence), with the
// polynomial coefficients taken from power traces (F is traceless), then the eigenvector as a
// column of adj(F - lambda I).  Returns false when the eigenvalue is (nearly) degenerate and the
// adjugate vanishes; the caller then falls back to Jacobi.
__device__ inline double det3(double a, double b, double c, double d, double e, double f, double g,
                              double h, double i) {
  return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}

// cofactor (ROW, COL) of a 4x4 matrix with compile-time indices (registers only)
template <int ROW, int COL>
__device__ __forceinline__ double cofactor4(const double (&a)[4][4]) {
  constexpr int r0 = ROW == 0 ? 1 : 0, r1 = ROW <= 1 ? 2 : 1, r2 = ROW <= 2 ? 3 : 2;
  constexpr int c0 = COL == 0 ? 1 : 0, c1 = COL <= 1 ? 2 : 1, c2 = COL <= 2 ? 3 : 2;
  const double d = det3(a[r0][c0], a[r0][c1], a[r0][c2], a[r1][c0], a[r1][c1], a[r1][c2], a[r2][c0],
                        a[r2][c1], a[r2][c2]);
  return ((ROW + COL) & 1) ? -d : d;
}
template <int ROW>
__device__ __forceinline__ void cofactor_row4(const double (&a)[4][4], double (&q)[4]) {
  q[0] = cofactor4<ROW, 0>(a);
  q[1] = cofactor4<ROW, 1>(a);
  q[2] = cofactor4<ROW, 2>(a);
  q[3] = cofactor4<ROW, 3>(a);
}

__device__ inline bool qcp4_max(const double (&f)[4][4], double lambda0, double& lambda,
                                double (&q)[4]) {
  double f2[4][4];
  double t2 = 0.0, t3 = 0.0, t4 = 0.0, scale = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) s += f[i][k] * f[k][j];
      f2[i][j] = f2[j][i] = s;
      scale = fmax(scale, fabs(f[i][j]));
    }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      t2 += f[i][j] * f[i][j];
      t3 += f2[i][j] * f[i][j];
      t4 += f2[i][j] * f2[i][j];
    }
  if (scale == 0.0) return false;
  const double c2 = -0.5 * t2, c1 = -t3 / 3.0, c0 = 0.25 * (0.5 * t2 * t2 - t4);
  // Newton from above: p > 0 and p' > 0 all the way down to the largest root, so the iterates
  // decrease monotonically; stop when the step drowns in rounding (or changes sign).
  double x = lambda0;
  for (int it = 0; it < 48; ++it) {
    const double x2 = x * x;
    const double p = (x2 + c2) * x2 + c1 * x + c0;
    const double dp = (4.0 * x2 + 2.0 * c2) * x + c1;
    if (!(dp > 0.0)) break;
    const double step = p / dp;
    if (!(step > 4e-16 * x)) break;
    x -= step;
  }
  lambda = x;
  double a[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = f[i][j] - (i == j ? x : 0.0);
  // adj(A) = c q q^T for a rank-3 A: the row with the largest diagonal cofactor is the best
  // conditioned copy of the eigenvector
  double row[4][4];
  cofactor_row4<0>(a, row[0]);
  cofactor_row4<1>(a, row[1]);
  cofactor_row4<2>(a, row[2]);
  cofactor_row4<3>(a, row[3]);
  double diag_best = fabs(row[0][0]);
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = row[0][k];
#pragma unroll
  for (int j = 1; j < 4; ++j) {
    const bool better = fabs(row[j][j]) > diag_best;
    diag_best = better ? fabs(row[j][j]) : diag_best;
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = better ? row[j][k] : q[k];
  }
  return diag_best > 1e-10 * scale * scale * scale;
}

// rmsd and U^T from the 13 sums of a single-segment group
__device__ inline void rmsd_from_sums(const double* sums, double gy, int n, bool allow_qcp,
                                      double& rmsd, double (&ut)[9], double (&centroid)[3]) {
  const double inv = 1.0 / n;
  const double gx = sums[3] - (sums[0] * sums[0] + sums[1] * sums[1] + sums[2] * sums[2]) * inv;
  centroid[0] = sums[0] * inv;
  centroid[1] = sums[1] * inv;
  centroid[2] = sums[2] * inv;
  const double Sxx = sums[4], Sxy = sums[5], Sxz = sums[6], Syx = sums[7], Syy = sums[8],
               Syz = sums[9], Szx = sums[10], Szy = sums[11], Szz = sums[12];
  double f[4][4] = {{Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
                    {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
                    {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
                    {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
  double lambda, q[4];
  if (!allow_qcp || !qcp4_max(f, 0.5 * (gx + gy), lambda, q)) jacobi4_max(f, lambda, q);
  const double norm = 1.0 / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double q0 = q[0] * norm, q1 = q[1] * norm, q2 = q[2] * norm, q3 = q[3] * norm;
  const double msd = (gx + gy - 2.0 * lambda) * inv;
  rmsd = msd < 1e-20 ? 0.0 : sqrt(msd);
  ut[0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3;
  ut[1] = 2 * (q1 * q2 + q0 * q3);
  ut[2] = 2 * (q1 * q3 - q0 * q2);
  ut[3] = 2 * (q1 * q2 - q0 * q3);
  ut[4] = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3;
  ut[5] = 2 * (q2 * q3 + q0 * q1);
  ut[6] = 2 * (q1 * q3 + q0 * q2);
  ut[7] = 2 * (q2 * q3 - q0 * q1);
  ut[8] = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
}

// One-read fused evaluation of a contiguous FP32 group (stream_fused.cuh).
struct RmsdFusedPolicy {
  static constexpr int NS = 13;
  static constexpr int AUX = 3;   // per-atom auxiliary data: the centred reference, float [n][3]
  static constexpr int AUXD = 3;  // ... and in double for the sums (exactly centred)
  static constexpr int FRAME_FLOATS = 13;
  struct Params {
    double gy;  // sum |y^|^2 of that reference
    int n;
  };
  struct Frame {
    float c[3];   // centroid
    float ut[9];  // U^T
    float f;      // 1 / (n rmsd)
    float pad[3];
  };
  __device__ static __forceinline__ void accumulate(const Params&, const float*, const double2* auxd,
                                                    const float (&x)[4], const float (&y)[4],
                                                    const float (&z)[4], double (&v)[NS]) {
    // shared memory, conflict-free: (r0x r0y) (r0z r1x) (r1y r1z) (r2x r2y) (r2z r3x) (r3y r3z)
    const double2 p0 = auxd[0], p1 = auxd[32], p2 = auxd[64], p3 = auxd[96], p4 = auxd[128],
                  p5 = auxd[160];
    const double rx[4] = {p0.x, p1.y, p3.x, p4.y};
    const double ry[4] = {p0.y, p2.x, p3.y, p5.x};
    const double rz[4] = {p1.x, p2.y, p4.x, p5.y};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double xd = x[k], yd = y[k], zd = z[k];
      v[0] += xd;
      v[1] += yd;
      v[2] += zd;
      v[3] = fma(xd, xd, v[3]);
      v[3] = fma(yd, yd, v[3]);
      v[3] = fma(zd, zd, v[3]);
      v[4] = fma(xd, rx[k], v[4]);
      v[5] = fma(xd, ry[k], v[5]);
      v[6] = fma(xd, rz[k], v[6]);
      v[7] = fma(yd, rx[k], v[7]);
      v[8] = fma(yd, ry[k], v[8]);
      v[9] = fma(yd, rz[k], v[9]);
      v[10] = fma(zd, rx[k], v[10]);
      v[11] = fma(zd, ry[k], v[11]);
      v[12] = fma(zd, rz[k], v[12]);
    }
  }
  // kept out of line so that the solver's registers do not burden the streaming loops
  __device__ static __noinline__ void complete(const Params& p, const double (&sums)[NS],
                                               Frame* frame, float* value, int accumulate) {
    double rmsd, ut[9], c[3];
    rmsd_from_sums(sums, p.gy, p.n, true, rmsd, ut, c);
    const double f = rmsd > 0.0 ? 1.0 / (p.n * rmsd) : 0.0;
    Frame fr;
#pragma unroll
    for (int k = 0; k < 9; ++k) fr.ut[k] = (float)ut[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) fr.c[k] = (float)c[k];
    fr.f = (float)f;
    fr.pad[0] = fr.pad[1] = fr.pad[2] = 0.f;
    *frame = fr;
    *value = accumulate ? *value + (float)rmsd : (float)rmsd;
  }
  __device__ static __forceinline__ void gradient(const Params&, const Frame& fr, const float* aux,
                                                  const float (&x)[4], const float (&y)[4],
                                                  const float (&z)[4], float (&gx)[4],
                                                  float (&gy)[4], float (&gz)[4]) {
    const float4* py = reinterpret_cast<const float4*>(aux);  // shared memory
    const float4 r0 = py[0], r1 = py[1], r2 = py[2];
    float rx[4], ry[4], rz[4];
    fused_unpack4(r0, r1, r2, rx, ry, rz);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // f ((x - c) - U^T y): one rounding per step, the products are not rounded on their own
      gx[k] = fr.f * fmaf(-fr.ut[0], rx[k], fmaf(-fr.ut[1], ry[k], fmaf(-fr.ut[2], rz[k], x[k] - fr.c[0])));
      gy[k] = fr.f * fmaf(-fr.ut[3], rx[k], fmaf(-fr.ut[4], ry[k], fmaf(-fr.ut[5], rz[k], y[k] - fr.c[1])));
      gz[k] = fr.f * fmaf(-fr.ut[6], rx[k], fmaf(-fr.ut[7], ry[k], fmaf(-fr.ut[8], rz[k], z[k] - fr.c[2])));
    }
  }
};

// Per (frame, group) result.
struct RmsdSolution {
  double rmsd;
  double ut[9];  // U^T, row-major: (U^T y)_c = sum_d ut[3c+d] y_d
};

__global__ void rmsd_solve_kernel(const double* __restrict__ partials, int nslice,
                                  const int32_t* __restrict__ group_seg_offsets,
                                  const int32_t* __restrict__ seg_slice_offsets,
                                  const int32_t* __restrict__ seg_size,
                                  const int32_t* __restrict__ group_size,
                                  const double* __restrict__ group_gy, int num_groups,
                                  int64_t num_frames, RmsdSolution* __restrict__ solution,
                                  double* __restrict__ seg_centroid, int nseg,
                                  int shared_moments = 0) {
  // shared_moments = SK > 0: every group covers the same atoms (rmsd_multi.cuh); segment s owns the
  // slices [s * SK, (s + 1) * SK) and sum x, sum |x|^2 are stored once, in the slices of group 0
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= num_frames * num_groups) return;
  const int64_t b = t / num_groups;
  const int g = (int)(t - b * num_groups);
  double r[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double gx = 0.0;
  for (int seg = group_seg_offsets[g]; seg < group_seg_offsets[g + 1]; ++seg) {
    double sx = 0, sy = 0, sz = 0, s2 = 0;
    const int s_begin = shared_moments ? seg * shared_moments : seg_slice_offsets[seg];
    const int s_end = shared_moments ? (seg + 1) * shared_moments : seg_slice_offsets[seg + 1];
    for (int s = s_begin; s < s_end; ++s) {
      const double* p = partials + ((size_t)b * nslice + s) * 13;
      const double* pm = shared_moments ? partials + ((size_t)b * nslice + (s - s_begin)) * 13 : p;
      sx += pm[0];
      sy += pm[1];
      sz += pm[2];
      s2 += pm[3];
#pragma unroll
      for (int k = 0; k < 9; ++k) r[k] += p[4 + k];
    }
    const double inv = 1.0 / seg_size[seg];
    gx += s2 - (sx * sx + sy * sy + sz * sz) * inv;
    double* c = seg_centroid + ((size_t)b * nseg + seg) * 3;
    c[0] = sx * inv;
    c[1] = sy * inv;
    c[2] = sz * inv;
  }
  const double Sxx = r[0], Sxy = r[1], Sxz = r[2], Syx = r[3], Syy = r[4], Syz = r[5], Szx = r[6],
               Szy = r[7], Szz = r[8];
  double f[4][4] = {{Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
                    {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
                    {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
                    {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
  double lambda, q[4];
  jacobi4_max(f, lambda, q);
  const double norm = 1.0 / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double q0 = q[0] * norm, q1 = q[1] * norm, q2 = q[2] * norm, q3 = q[3] * norm;
  const int n = group_size[g];
  const double msd = (gx + group_gy[g] - 2.0 * lambda) / n;
  RmsdSolution sol;
  sol.rmsd = msd < 1e-20 ? 0.0 : sqrt(msd);
  // U rotates x^ onto y^; store its transpose
  const double u00 = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, u01 = 2 * (q1 * q2 - q0 * q3),
               u02 = 2 * (q1 * q3 + q0 * q2), u10 = 2 * (q1 * q2 + q0 * q3),
               u11 = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, u12 = 2 * (q2 * q3 - q0 * q1),
               u20 = 2 * (q1 * q3 - q0 * q2), u21 = 2 * (q2 * q3 + q0 * q1),
               u22 = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
  sol.ut[0] = u00; sol.ut[1] = u10; sol.ut[2] = u20;
  sol.ut[3] = u01; sol.ut[4] = u11; sol.ut[5] = u21;
  sol.ut[6] = u02; sol.ut[7] = u12; sol.ut[8] = u22;
  solution[t] = sol;
}

// ---- combine -----------------------------------------------------------------------------------
struct CombineParams {
  int mode;
  StepParams step;
  double inv_r0, scale, lambda;  // lambda = 1/(2 sigma^2)
};

// One thread per frame.  Writes the CV value and, per group, the factor multiplying
// (x_i - c - U^T y_i) in the gradient: dCV/drmsd_g / (n_g rmsd_g)  (0 for a degenerate group).
template <typename T>
__global__ void rmsd_combine_kernel(const RmsdSolution* __restrict__ solution,
                                    const int32_t* __restrict__ group_size, int num_groups,
                                    CombineParams cp, T* __restrict__ value,
                                    double* __restrict__ group_factor, int accumulate,
                                    int64_t num_frames) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= num_frames) return;
  const RmsdSolution* sol = solution + b * num_groups;
  double* factor = group_factor + b * num_groups;
  double result = 0.0;
  if (cp.mode == CVP_COMBINE_IDENTITY) {
    result = sol[0].rmsd;
    factor[0] = result > 0.0 ? 1.0 / (group_size[0] * result) : 0.0;
  } else if (cp.mode == CVP_COMBINE_STEP_SUM) {
    for (int g = 0; g < num_groups; ++g) {
      const double d = sol[g].rmsd;
      double s, ds;
      step_function(cp.step, d * cp.inv_r0, s, ds);
      result += s;
      factor[g] = d > 0.0 ? cp.scale * ds * cp.inv_r0 / (group_size[g] * d) : 0.0;
    }
    result *= cp.scale;
  } else {
    // x_g = rmsd_g^2, w_g = exp(lambda (xmin - x_g))  (base_path_cv.py:44-59)
    double xmin = sol[0].rmsd * sol[0].rmsd;
    for (int g = 1; g < num_groups; ++g) xmin = fmin(xmin, sol[g].rmsd * sol[g].rmsd);
    double wsum = 0.0, iwsum = 0.0;
    for (int g = 0; g < num_groups; ++g) {
      const double w = exp(cp.lambda * (xmin - sol[g].rmsd * sol[g].rmsd));
      wsum += w;
      iwsum += g * w;
    }
    const double s = iwsum / ((num_groups - 1) * wsum);
    result = cp.mode == CVP_COMBINE_PATH_PROGRESS ? s : xmin - log(wsum) / cp.lambda;
    for (int g = 0; g < num_groups; ++g) {
      const double w = exp(cp.lambda * (xmin - sol[g].rmsd * sol[g].rmsd));
      // dCV/dx_g, then dx_g/drmsd_g = 2 rmsd_g, then 1/(n_g rmsd_g)
      const double dx = cp.mode == CVP_COMBINE_PATH_PROGRESS
                            ? -cp.lambda * w * ((double)g / (num_groups - 1) - s) / wsum
                            : w / wsum;
      factor[g] = sol[g].rmsd > 0.0 ? 2.0 * dx / group_size[g] : 0.0;
    }
  }
  value[b] = accumulate ? T((double)value[b] + result) : T(result);
}

// ---- gradient ----------------------------------------------------------------------------------
template <typename T>
__global__ void rmsd_gradient_kernel(const T* __restrict__ pos, int64_t num_atoms,
                                     const int32_t* __restrict__ active_atom,
                                     const int32_t* __restrict__ atom_entry_offsets,
                                     const int32_t* __restrict__ atom_entries, int num_active,
                                     const int32_t* __restrict__ entry_group,
                                     const int32_t* __restrict__ entry_seg,
                                     const double* __restrict__ ref,
                                     const RmsdSolution* __restrict__ solution,
                                     const double* __restrict__ seg_centroid,
                                     const double* __restrict__ group_factor, int num_groups,
                                     int nseg, T* __restrict__ grad, int accumulate,
                                     int64_t total) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int64_t b = t / num_active;
  const int k = (int)(t - b * num_active);
  const int64_t atom = active_atom[k];
  const T* p = pos + (b * num_atoms + atom) * 3;
  const double x = p[0], y = p[1], z = p[2];
  double gx = 0, gy = 0, gz = 0;
  for (int m = atom_entry_offsets[k]; m < atom_entry_offsets[k + 1]; ++m) {
    const int e = atom_entries[m];
    const int g = entry_group[e];
    const double f = group_factor[b * num_groups + g];
    if (f == 0.0) continue;
    const double* c = seg_centroid + ((size_t)b * nseg + entry_seg[e]) * 3;
    const double* ut = solution[b * num_groups + g].ut;
    const double rx = ref[3 * (int64_t)e], ry = ref[3 * (int64_t)e + 1], rz = ref[3 * (int64_t)e + 2];
    gx += f * (x - c[0] - (ut[0] * rx + ut[1] * ry + ut[2] * rz));
    gy += f * (y - c[1] - (ut[3] * rx + ut[4] * ry + ut[5] * rz));
    gz += f * (z - c[2] - (ut[6] * rx + ut[7] * ry + ut[8] * rz));
  }
  T* out = grad + (b * num_atoms + atom) * 3;
  if (accumulate) {
    out[0] += T(gx);
    out[1] += T(gy);
    out[2] += T(gz);
  } else {
    out[0] = T(gx);
    out[1] = T(gy);
    out[2] = T(gz);
  }
}

// ---- many small groups over few atoms: one CTA per frame -------------------------------------------
// HelixRMSDContent / SheetRMSDContent (base_rmsd_content.py:56-102): hundreds of overlapping
// 30-atom blocks over the backbone of one chain, value = scale * sum_g S(rmsd_g / r0).  The generic
// path runs four kernels and moves every block through global memory (partials, solutions,
// factors); here the frame's active atoms are staged once in shared memory, every thread owns whole
// blocks (13 FP64 sums, QCP solve, step function), the per-block constants stay in shared memory
// and the gradient is gathered per atom from them (single writer, fixed order: deterministic).
constexpr int BLOCKS_THREADS = 256;
constexpr int BLOCKS_SMEM_BUDGET = 200 * 1024;  // (two CTAs per SM up to half of it)
constexpr int BLOCKS_SOL = 13;                  // doubles per block: c[3], U^T[9], factor

template <typename T>
__global__ void __launch_bounds__(BLOCKS_THREADS, 2)
    rmsd_blocks_kernel(const T* __restrict__ pos, int64_t num_atoms,
                       const int32_t* __restrict__ active_atom, int num_active,
                       const int32_t* __restrict__ entry_slot, const int32_t* __restrict__ entry_group,
                       const int32_t* __restrict__ group_entry_offsets, const double* __restrict__ ref,
                       int shared_ref_entries, const double* __restrict__ group_gy, int num_groups,
                       const int32_t* __restrict__ atom_entry_offsets,
                       const int32_t* __restrict__ atom_entries, CombineParams cp,
                       T* __restrict__ value, T* __restrict__ grad, int accumulate) {
  extern __shared__ __align__(16) unsigned char blocks_smem[];
  double* s_sol = reinterpret_cast<double*>(blocks_smem);                 // [num_groups][13]
  double* s_ref = s_sol + (size_t)num_groups * BLOCKS_SOL;                // [shared_ref_entries][3]
  T* s_x = reinterpret_cast<T*>(s_ref + (size_t)shared_ref_entries * 3);  // [num_active][3]
  __shared__ double s_red[BLOCKS_THREADS];
  const int64_t b = blockIdx.x;
  const T* frame = pos + b * num_atoms * 3;
  for (int k = threadIdx.x; k < num_active; k += BLOCKS_THREADS) {
    const T* p = frame + 3 * (int64_t)active_atom[k];
    s_x[3 * k] = p[0];
    s_x[3 * k + 1] = p[1];
    s_x[3 * k + 2] = p[2];
  }
  for (int k = threadIdx.x; k < shared_ref_entries * 3; k += BLOCKS_THREADS) s_ref[k] = ref[k];
  __syncthreads();

  double local = 0.0;
  for (int g = threadIdx.x; g < num_groups; g += BLOCKS_THREADS) {
    const int begin = group_entry_offsets[g], end = group_entry_offsets[g + 1];
    double v[13];
#pragma unroll
    for (int k = 0; k < 13; ++k) v[k] = 0.0;
#pragma unroll 6
    for (int e = begin; e < end; ++e) {  // (unrolled: the slot loads of several entries in flight)
      const int slot = entry_slot[e];
      const double x = s_x[3 * slot], y = s_x[3 * slot + 1], z = s_x[3 * slot + 2];
      const double* r = shared_ref_entries ? s_ref + 3 * (e - begin) : ref + 3 * (int64_t)e;
      const double rx = r[0], ry = r[1], rz = r[2];
      v[0] += x;
      v[1] += y;
      v[2] += z;
      v[3] += x * x + y * y + z * z;
      v[4] += x * rx;
      v[5] += x * ry;
      v[6] += x * rz;
      v[7] += y * rx;
      v[8] += y * ry;
      v[9] += y * rz;
      v[10] += z * rx;
      v[11] += z * ry;
      v[12] += z * rz;
    }
    double rmsd, ut[9], c[3];
    // QCP (Newton on the characteristic polynomial) in FP32 mode, cyclic Jacobi in FP64 mode, as
    // the generic solve kernel: the 1e-10 bar of the FP64 mode leaves no room for the adjugate
    rmsd_from_sums(v, group_gy[g], end - begin, sizeof(T) == 4, rmsd, ut, c);
    double sv, ds;
    step_function(cp.step, rmsd * cp.inv_r0, sv, ds);
    local += sv;
    double* sol = s_sol + (size_t)g * BLOCKS_SOL;
    sol[0] = c[0];
    sol[1] = c[1];
    sol[2] = c[2];
#pragma unroll
    for (int k = 0; k < 9; ++k) sol[3 + k] = ut[k];
    sol[12] = rmsd > 0.0 ? cp.scale * ds * cp.inv_r0 / ((end - begin) * rmsd) : 0.0;
  }
  s_red[threadIdx.x] = local;
  __syncthreads();
  for (int stride = BLOCKS_THREADS / 2; stride > 0; stride >>= 1) {
    if (threadIdx.x < stride) s_red[threadIdx.x] += s_red[threadIdx.x + stride];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double result = cp.scale * s_red[0];
    value[b] = accumulate ? value[b] + T(result) : T(result);
  }
  if (grad == nullptr) return;
  for (int k = threadIdx.x; k < num_active; k += BLOCKS_THREADS) {
    const double x = s_x[3 * k], y = s_x[3 * k + 1], z = s_x[3 * k + 2];
    double gx = 0, gy = 0, gz = 0;
    for (int m = atom_entry_offsets[k]; m < atom_entry_offsets[k + 1]; ++m) {
      const int e = atom_entries[m];
      const int g = entry_group[e];
      const double* sol = s_sol + (size_t)g * BLOCKS_SOL;
      const double f = sol[12];
      if (f == 0.0) continue;
      const double* r = shared_ref_entries ? s_ref + 3 * (e - group_entry_offsets[g]) : ref + 3 * (int64_t)e;
      const double rx = r[0], ry = r[1], rz = r[2];
      gx += f * (x - sol[0] - (sol[3] * rx + sol[4] * ry + sol[5] * rz));
      gy += f * (y - sol[1] - (sol[6] * rx + sol[7] * ry + sol[8] * rz));
      gz += f * (z - sol[2] - (sol[9] * rx + sol[10] * ry + sol[11] * rz));
    }
    T* out = grad + (b * num_atoms + active_atom[k]) * 3;
    if (accumulate) {
      out[0] += T(gx);
      out[1] += T(gy);
      out[2] += T(gz);
    } else {
      out[0] = T(gx);
      out[1] = T(gy);
      out[2] = T(gz);
    }
  }
}

// gradient of the streaming fast path: g = factor * (x - c - U^T y), 4 atoms per thread iteration
__global__ void __launch_bounds__(RMSD_FAST_THREADS)
    rmsd_gradient_fast_kernel(const float* __restrict__ pos, int64_t num_atoms, int first_atom,
                              int n, const float* __restrict__ ref,
                              const RmsdSolution* __restrict__ solution,
                              const double* __restrict__ seg_centroid,
                              const double* __restrict__ group_factor, int nslice,
                              float* __restrict__ grad, int accumulate) {
  const int64_t b = blockIdx.x / nslice;
  const int s = blockIdx.x - (int)b * nslice;
  const float4* px = reinterpret_cast<const float4*>(pos + (b * num_atoms + first_atom) * 3);
  const float4* py = reinterpret_cast<const float4*>(ref);
  float4* pg = reinterpret_cast<float4*>(grad + (b * num_atoms + first_atom) * 3);
  const int g_begin = s * (RMSD_FAST_SLICE / 4);
  const int g_end = min(n / 4, g_begin + RMSD_FAST_SLICE / 4);
  const float f = (float)group_factor[b];
  float ut[9], c[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) ut[k] = (float)solution[b].ut[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) c[k] = (float)seg_centroid[3 * b + k];
  for (int g = g_begin + threadIdx.x; g < g_end; g += RMSD_FAST_THREADS) {
    const float4 a0 = __ldg(px + 3 * g), a1 = __ldg(px + 3 * g + 1), a2 = __ldg(px + 3 * g + 2);
    const float4 r0 = __ldg(py + 3 * g), r1 = __ldg(py + 3 * g + 1), r2 = __ldg(py + 3 * g + 2);
    float x[4], y[4], z[4], rx[4], ry[4], rz[4], gx[4], gy[4], gz[4];
    unpack4(a0, a1, a2, x, y, z);
    unpack4(r0, r1, r2, rx, ry, rz);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      gx[k] = f * ((x[k] - c[0]) - (ut[0] * rx[k] + ut[1] * ry[k] + ut[2] * rz[k]));
      gy[k] = f * ((y[k] - c[1]) - (ut[3] * rx[k] + ut[4] * ry[k] + ut[5] * rz[k]));
      gz[k] = f * ((z[k] - c[2]) - (ut[6] * rx[k] + ut[7] * ry[k] + ut[8] * rz[k]));
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

struct RmsdCV : cvp_cv {
  // streaming fast path (see rmsd_sums_fast_kernel)
  bool fast = false;
  int use_fused = 1;    // option "fused": one-read persistent kernel for the streaming fast path
  // measured on the 100k-atom batch: the per-frame solve (record gather + QCP) takes a few rounds,
  // and waiting for it costs more than the L2 misses of a long lag
  int fused_lag = 2;    // option "fused_lag": rounds between the sums and the gradient of a frame
  int fused_slice = 640;  // option "fused_slice": atoms per sub-slice (0 = chosen by the planner)
  int fused_slots = 0;  // option "fused_slots": auxiliary-data slots per CTA (0 = planner)
  int fused_hints = 3;  // option "fused_hints": L2 eviction hints (see FusedGeom::l2_hints)
  int fast_first = 0, fast_nslice = 0;
  // many references over one contiguous atom set (rmsd_multi.cuh)
  bool multi = false;
  int use_multi = 1;  // option "multi"
  // option "multi_mma": FP64 tensor-core (DMMA) versions of the multi kernels; 1 = gradient operands
  // pre-packed in fragment order and fed by TMA, 2 = register-staged gradient, 0 = SIMT DFMA,
  // 3 = like 1 with the covariance operands fed by TMA as well (measured: 0.93 ms against 0.89 ms
  // for the register-staged covariance kernel -- kept for the record, not the default).  The
  // covariance of mode 1 is the three-buffer kernel without block-wide barriers, of modes 2 and 3
  // (when the rows are not aligned) the double-buffered one.
  int use_multi_mma = 1;
  std::vector<double> ref_pack;  // references in the fragment order of rmsd_multi_gradient_pk_kernel
  std::vector<double> ref_pack_cov;  // ... and of rmsd_multi_cov_pk_kernel
  DeviceBuffer d_ref_pack, d_ref_pack_cov;
  int multi_first = 0, multi_n = 0, multi_sk_max = 1;
  std::vector<float> ref_f32;
  std::vector<int32_t> fast_seg_slice_offsets;
  DeviceBuffer d_ref_f32, d_ref_perm, d_fast_seg_slice_offsets;
  std::vector<double> ref_perm;  // the centred reference in the layout of fused_permute_doubles()
  int num_groups = 0, nseg = 0, nslice = 0;
  int64_t num_entries = 0;
  CombineParams cp{};
  std::vector<int32_t> entry_atom, entry_group, entry_seg;
  std::vector<double> ref;  // centred per segment
  std::vector<RmsdSlice> slices;
  std::vector<int32_t> group_seg_offsets, seg_slice_offsets, seg_size, group_size;
  std::vector<double> group_gy;
  std::vector<int32_t> atom_entry_offsets, atom_entries;
  bool dense = false;  // every atom of the system is active
  // many small single-body groups summed through a step function (rmsd_blocks_kernel)
  bool blocks = false;
  int use_blocks = 1;          // option "blocks"
  int blocks_shared_ref = 0;   // entries of the reference all groups share (0: per-group references)
  std::vector<int32_t> entry_slot, group_entry_offsets;
  DeviceBuffer d_entry_slot, d_group_entry_offsets;
  DeviceBuffer d_entry_atom, d_entry_group, d_entry_seg, d_ref, d_slices, d_group_seg_offsets,
      d_seg_slice_offsets, d_seg_size, d_group_size, d_group_gy, d_active, d_atom_entry_offsets,
      d_atom_entries;

  ~RmsdCV() override {
    for (DeviceBuffer* buf :
         {&d_entry_atom, &d_entry_group, &d_entry_seg, &d_ref, &d_slices, &d_group_seg_offsets,
          &d_seg_slice_offsets, &d_seg_size, &d_group_size, &d_group_gy, &d_active,
          &d_atom_entry_offsets, &d_atom_entries, &d_ref_f32, &d_ref_perm, &d_fast_seg_slice_offsets,
          &d_entry_slot, &d_group_entry_offsets, &d_ref_pack, &d_ref_pack_cov})
      buf->release();
  }
  int upload() override {
    int status;
#define CVP_UP(buf, vec) \
  if ((status = cvp::upload(buf, vec)) != CVP_OK) return status
    CVP_UP(d_entry_atom, entry_atom);
    CVP_UP(d_entry_group, entry_group);
    CVP_UP(d_entry_seg, entry_seg);
    CVP_UP(d_ref, ref);
    CVP_UP(d_slices, slices);
    CVP_UP(d_group_seg_offsets, group_seg_offsets);
    CVP_UP(d_seg_slice_offsets, seg_slice_offsets);
    CVP_UP(d_seg_size, seg_size);
    CVP_UP(d_group_size, group_size);
    CVP_UP(d_group_gy, group_gy);
    CVP_UP(d_active, active_atoms);
    CVP_UP(d_atom_entry_offsets, atom_entry_offsets);
    CVP_UP(d_atom_entries, atom_entries);
    CVP_UP(d_ref_f32, ref_f32);
    CVP_UP(d_ref_perm, ref_perm);
    CVP_UP(d_fast_seg_slice_offsets, fast_seg_slice_offsets);
    CVP_UP(d_entry_slot, entry_slot);
    CVP_UP(d_group_entry_offsets, group_entry_offsets);
    CVP_UP(d_ref_pack, ref_pack);
    CVP_UP(d_ref_pack_cov, ref_pack_cov);
#undef CVP_UP
    return CVP_OK;
  }
  size_t workspace_per_frame(int, bool) const override {
    // (the fused kernel's records only exist for the single contiguous group it serves)
    return (fast ? FusedLaunch<RmsdFusedPolicy>::workspace_per_frame((int)num_entries) : 0) +
           (size_t)std::max(nslice, fast_nslice) * 13 * sizeof(double) + (size_t)num_groups * sizeof(RmsdSolution) +
           (size_t)nseg * 3 * sizeof(double) + (size_t)num_groups * sizeof(double) + 64 + w_pack_bytes(1);
  }
  size_t workspace_fixed(int) const override {
    // (the packed W records are padded to whole tiles of MULTI_TF frames)
    return 7 * 256 + FusedLaunch<RmsdFusedPolicy>::workspace_fixed() + w_pack_bytes(MULTI_TF);
  }
  size_t w_pack_bytes(int64_t frames) const {
    return multi ? (size_t)frames * pk_stages(num_groups) * PK_KG * 9 * sizeof(double) : 0;
  }
  int set_option(const char* name, int value) override {
    const std::string key(name);
    if (key == "fused") {
      std::swap(use_fused, value);
      return value;
    }
    if (key == "multi") {
      std::swap(use_multi, value);
      return value;
    }
    if (key == "multi_mma") {
      std::swap(use_multi_mma, value);
      return value;
    }
    if (key == "blocks") {
      std::swap(use_blocks, value);
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
  int64_t max_frames_per_run() const override {
    int64_t per_frame = std::max<int64_t>(nslice, (int64_t)active_atoms.size() / 64 + 1);
    return std::max<int64_t>(1, (int64_t(1) << 30) / per_frame);
  }

  size_t blocks_smem_bytes(size_t coordinate_bytes) const {
    return ((size_t)num_groups * BLOCKS_SOL + (size_t)blocks_shared_ref * 3) * sizeof(double) +
           active_atoms.size() * 3 * coordinate_bytes;
  }

  template <typename T>
  int run_type(const EvalArgs& a, void* ws) {
    const int64_t B = a.num_frames, N = a.num_atoms;
    const bool accumulate = (a.flags & CVP_FLAG_ACCUMULATE) != 0;
    const T* pos = static_cast<const T*>(a.positions);
    T* value = static_cast<T*>(a.value);
    T* grad = static_cast<T*>(a.grad);
    cudaStream_t st = a.stream;
    Carver carve(ws);
    double* partials = carve.take<double>((size_t)B * nslice * 13);
    RmsdSolution* solution = carve.take<RmsdSolution>((size_t)B * num_groups);
    double* seg_centroid = carve.take<double>((size_t)B * nseg * 3);
    double* group_factor = carve.take<double>((size_t)B * num_groups);
    double* w_pack = multi ? carve.take<double>(w_pack_bytes(div_up(B, MULTI_TF) * MULTI_TF) / sizeof(double))
                           : nullptr;

    // streaming fast path: FP32, one contiguous aligned group
    const bool use_fast = fast && std::is_same<T, float>::value && (N % 4) == 0 &&
                          aligned16(a.positions, a.grad);
    FusedGeom geom;
    if (use_fast && use_fused && cp.mode == CVP_COMBINE_IDENTITY && B >= 64 &&
        FusedLaunch<RmsdFusedPolicy>::plan((int)num_entries, N, fast_first, B, grad != nullptr,
                                           fused_lag, fused_slice, fused_slots, fused_hints, sm_count(), &geom)) {
      RmsdFusedPolicy::Params params;
      params.gy = group_gy[0];
      params.n = (int)num_entries;
      if (grad != nullptr && !accumulate && !dense)
        CVP_CUDA_CHECK(cudaMemsetAsync(grad, 0, (size_t)B * N * 3 * sizeof(T), st));
      timer.begin("rmsd_fused", st);
      int status = FusedLaunch<RmsdFusedPolicy>::launch(
          reinterpret_cast<const float*>(pos), static_cast<const float*>(d_ref_f32.ptr),
          static_cast<const double*>(d_ref_perm.ptr), geom, sm_count(), params, ws, reinterpret_cast<float*>(value), reinterpret_cast<float*>(grad),
          accumulate, st);
      timer.end(st);
      return status;
    }
    if (blocks && use_blocks && blocks_smem_bytes(sizeof(T)) <= (size_t)BLOCKS_SMEM_BUDGET) {
      const size_t smem = blocks_smem_bytes(sizeof(T));
      static bool blocks_configured = false;
      if (!blocks_configured) {
        CVP_CUDA_CHECK(cudaFuncSetAttribute(rmsd_blocks_kernel<T>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            BLOCKS_SMEM_BUDGET));
        blocks_configured = true;
      }
      if (grad != nullptr && !accumulate && !dense)
        CVP_CUDA_CHECK(cudaMemsetAsync(grad, 0, (size_t)B * N * 3 * sizeof(T), st));
      timer.begin("rmsd_blocks", st);
      rmsd_blocks_kernel<T><<<(unsigned)B, BLOCKS_THREADS, smem, st>>>(
          pos, N, static_cast<const int32_t*>(d_active.ptr), (int)active_atoms.size(),
          static_cast<const int32_t*>(d_entry_slot.ptr), static_cast<const int32_t*>(d_entry_group.ptr),
          static_cast<const int32_t*>(d_group_entry_offsets.ptr), static_cast<const double*>(d_ref.ptr),
          blocks_shared_ref, static_cast<const double*>(d_group_gy.ptr), num_groups,
          static_cast<const int32_t*>(d_atom_entry_offsets.ptr),
          static_cast<const int32_t*>(d_atom_entries.ptr), cp, value, grad, accumulate ? 1 : 0);
      timer.end(st);
      CVP_LAUNCH_CHECK();
      return CVP_OK;
    }
    const bool use_multi_path = multi && use_multi && std::is_same<T, float>::value;
    if (use_multi_path) {
      // slices over the atoms: about two waves of CTAs (2 per SM)
      const int64_t tiles = div_up(B, MULTI_TF) * div_up(num_groups, MULTI_TG);
      const int multi_sk =
          (int)std::max<int64_t>(1, std::min<int64_t>(multi_sk_max, (4 * sm_count() + tiles / 2) / tiles));
      const int ns_multi = num_groups * multi_sk;
      const dim3 cov_grid((unsigned)div_up(B, MULTI_TF), (unsigned)div_up(num_groups, MULTI_TG),
                          (unsigned)multi_sk);
      // whole rows of the tiles as bulk copies when every row starts and ends on 16 bytes
      const bool bulk_rows = multi_first % 4 == 0 && N % 4 == 0 && multi_n % 4 == 0 &&
                             (reinterpret_cast<uintptr_t>(pos) & 15) == 0;
      static bool cov_configured = false;
      if (!cov_configured) {
        CVP_CUDA_CHECK(cudaFuncSetAttribute(rmsd_multi_cov_pk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)MULTI_COV_PK_SMEM));
        CVP_CUDA_CHECK(cudaFuncSetAttribute(rmsd_multi_cov_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)MULTI_COV3_SMEM));
        cov_configured = true;
      }
      timer.begin("rmsd_multi_cov", st);
      if (use_multi_mma == 1)
        rmsd_multi_cov_ring_kernel<<<cov_grid, MULTI_THREADS, MULTI_COV3_SMEM, st>>>(
            reinterpret_cast<const float*>(pos), N, multi_first, multi_n,
            static_cast<const double*>(d_ref.ptr), num_groups, B, multi_sk, partials);
      else if (use_multi_mma == 3 && bulk_rows)
        rmsd_multi_cov_pk_kernel<<<cov_grid, MULTI_THREADS, MULTI_COV_PK_SMEM, st>>>(
            reinterpret_cast<const float*>(pos), N, multi_first, multi_n,
            static_cast<const double*>(d_ref_pack_cov.ptr), num_groups, B, multi_sk, partials);
      else if (use_multi_mma)
        rmsd_multi_cov_mma_kernel<<<cov_grid, MULTI_THREADS, 0, st>>>(
            reinterpret_cast<const float*>(pos), N, multi_first, multi_n,
            static_cast<const double*>(d_ref.ptr), num_groups, B, multi_sk, partials);
      else
        rmsd_multi_cov_kernel<<<cov_grid, MULTI_THREADS, 0, st>>>(
            reinterpret_cast<const float*>(pos), N, multi_first, multi_n,
            static_cast<const double*>(d_ref.ptr), num_groups, B, multi_sk, partials);
      CVP_LAUNCH_CHECK();
      rmsd_multi_moments_kernel<<<(unsigned)div_up(B * multi_sk * 32, 256), 256, 0, st>>>(
          reinterpret_cast<const float*>(pos), N, multi_first, multi_n, num_groups, B, multi_sk,
          partials);
      timer.end(st);
      CVP_LAUNCH_CHECK();
      rmsd_solve_kernel<<<(unsigned)div_up(B * num_groups, 64), 64, 0, st>>>(
          partials, ns_multi, static_cast<const int32_t*>(d_group_seg_offsets.ptr),
          nullptr, static_cast<const int32_t*>(d_seg_size.ptr),
          static_cast<const int32_t*>(d_group_size.ptr), static_cast<const double*>(d_group_gy.ptr),
          num_groups, B, solution, seg_centroid, nseg, multi_sk);
      CVP_LAUNCH_CHECK();
      rmsd_combine_kernel<T><<<(unsigned)div_up(B, 128), 128, 0, st>>>(
          solution, static_cast<const int32_t*>(d_group_size.ptr), num_groups, cp, value,
          group_factor, accumulate ? 1 : 0, B);
      CVP_LAUNCH_CHECK();
      if (grad == nullptr) return CVP_OK;
      if (!accumulate && !dense)
        CVP_CUDA_CHECK(cudaMemsetAsync(grad, 0, (size_t)B * N * 3 * sizeof(T), st));
      const dim3 grad_grid((unsigned)div_up(B, MULTI_TF),
                           (unsigned)div_up(multi_n, use_multi_mma ? MMA_TA : MULTI_TA));
      static bool multi_configured = false;
      if (!multi_configured) {
        CVP_CUDA_CHECK(cudaFuncSetAttribute(rmsd_multi_gradient_kernel,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)MULTI_GRAD_SMEM));
        CVP_CUDA_CHECK(cudaFuncSetAttribute(rmsd_multi_gradient_mma_kernel,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)MULTI_GRAD_MMA_SMEM));
        CVP_CUDA_CHECK(cudaFuncSetAttribute(rmsd_multi_gradient_pk_kernel<true>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)MULTI_GRAD_PK_SMEM));
        CVP_CUDA_CHECK(cudaFuncSetAttribute(rmsd_multi_gradient_pk_kernel<false>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)MULTI_GRAD_PK_SMEM));
        multi_configured = true;
      }
      if (use_multi_mma == 1 || use_multi_mma == 3) {
        const int64_t padded = div_up(B, MULTI_TF) * MULTI_TF * pk_stages(num_groups) * PK_KG;
        rmsd_multi_weights_pack_kernel<RmsdSolution><<<(unsigned)div_up(padded, 256), 256, 0, st>>>(
            solution, group_factor, B, num_groups, w_pack);
        CVP_LAUNCH_CHECK();
        // whole rows of the tiles as bulk copies when every row starts and ends on 16 bytes
        const bool bulk = multi_first % 4 == 0 && N % 4 == 0 && multi_n % 4 == 0 &&
                          ((reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(grad)) & 15) == 0;
        auto kernel = bulk ? rmsd_multi_gradient_pk_kernel<true> : rmsd_multi_gradient_pk_kernel<false>;
        timer.begin("rmsd_multi_gradient", st);
        kernel<<<grad_grid, MULTI_THREADS, MULTI_GRAD_PK_SMEM, st>>>(
            reinterpret_cast<const float*>(pos), N, multi_first, multi_n,
            static_cast<const double*>(d_ref_pack.ptr), num_groups, B, w_pack, group_factor,
            seg_centroid, nseg, reinterpret_cast<float*>(grad), accumulate ? 1 : 0);
        timer.end(st);
        CVP_LAUNCH_CHECK();
        return CVP_OK;
      }
      // the slice records are dead after the solve: their memory holds the W = factor * U^T records
      double* weights = partials;
      static_assert(sizeof(double) * 10 <= sizeof(double) * 13, "W records fit the slice records");
      rmsd_multi_weights_kernel<RmsdSolution><<<(unsigned)div_up(B * num_groups, 256), 256, 0, st>>>(
          solution, group_factor, B * (int64_t)num_groups, weights);
      CVP_LAUNCH_CHECK();
      timer.begin("rmsd_multi_gradient", st);
      if (use_multi_mma)
        rmsd_multi_gradient_mma_kernel<<<grad_grid, MULTI_THREADS, MULTI_GRAD_MMA_SMEM, st>>>(
            reinterpret_cast<const float*>(pos), N, multi_first, multi_n,
            static_cast<const double*>(d_ref.ptr), num_groups, B, weights,
            seg_centroid, nseg, reinterpret_cast<float*>(grad), accumulate ? 1 : 0);
      else
        rmsd_multi_gradient_kernel<<<grad_grid, MULTI_THREADS, MULTI_GRAD_SMEM, st>>>(
            reinterpret_cast<const float*>(pos), N, multi_first, multi_n,
            static_cast<const double*>(d_ref.ptr), num_groups, B, weights,
            seg_centroid, nseg, reinterpret_cast<float*>(grad), accumulate ? 1 : 0);
      timer.end(st);
      CVP_LAUNCH_CHECK();
      return CVP_OK;
    }
    const int ns = use_fast ? fast_nslice : nslice;
    if (use_fast) {
      timer.begin("rmsd_sums", st);
      rmsd_sums_fast_kernel<<<(unsigned)(B * ns), RMSD_FAST_THREADS, 0, st>>>(
          reinterpret_cast<const float*>(pos), N, fast_first, (int)num_entries,
          static_cast<const double*>(d_ref.ptr), ns, partials);
      timer.end(st);
      CVP_LAUNCH_CHECK();
    } else {
      int64_t warps = B * nslice;
      timer.begin("rmsd_sums", st);
      rmsd_sums_kernel<T><<<(unsigned)div_up(warps, RMSD_WARPS), RMSD_WARPS * 32, 0, st>>>(
          pos, N, static_cast<const int32_t*>(d_entry_atom.ptr),
          static_cast<const double*>(d_ref.ptr), static_cast<const RmsdSlice*>(d_slices.ptr),
          nslice, B, partials);
      timer.end(st);
      CVP_LAUNCH_CHECK();
    }
    {
      int64_t total = B * num_groups;
      timer.begin("rmsd_solve", st);
      rmsd_solve_kernel<<<(unsigned)div_up(total, 64), 64, 0, st>>>(
          partials, ns, static_cast<const int32_t*>(d_group_seg_offsets.ptr),
          static_cast<const int32_t*>(use_fast ? d_fast_seg_slice_offsets.ptr
                                               : d_seg_slice_offsets.ptr),
          static_cast<const int32_t*>(d_seg_size.ptr),
          static_cast<const int32_t*>(d_group_size.ptr),
          static_cast<const double*>(d_group_gy.ptr), num_groups, B,
          solution, seg_centroid, nseg);
      timer.end(st);
      CVP_LAUNCH_CHECK();
    }
    rmsd_combine_kernel<T><<<(unsigned)div_up(B, 128), 128, 0, st>>>(
        solution, static_cast<const int32_t*>(d_group_size.ptr), num_groups, cp, value,
        group_factor, accumulate ? 1 : 0, B);
    CVP_LAUNCH_CHECK();
    if (grad == nullptr) return CVP_OK;
    if (!accumulate && !dense)
      CVP_CUDA_CHECK(cudaMemsetAsync(grad, 0, (size_t)B * N * 3 * sizeof(T), st));
    if (use_fast) {
      timer.begin("rmsd_gradient", st);
      rmsd_gradient_fast_kernel<<<(unsigned)(B * ns), RMSD_FAST_THREADS, 0, st>>>(
          reinterpret_cast<const float*>(pos), N, fast_first, (int)num_entries,
          static_cast<const float*>(d_ref_f32.ptr), solution, seg_centroid, group_factor, ns,
          reinterpret_cast<float*>(grad), accumulate ? 1 : 0);
      timer.end(st);
      CVP_LAUNCH_CHECK();
      return CVP_OK;
    }
    {
      const int num_active = (int)active_atoms.size();
      int64_t total = B * (int64_t)num_active;
      timer.begin("rmsd_gradient", st);
      rmsd_gradient_kernel<T><<<(unsigned)div_up(total, 256), 256, 0, st>>>(
          pos, N, static_cast<const int32_t*>(d_active.ptr),
          static_cast<const int32_t*>(d_atom_entry_offsets.ptr),
          static_cast<const int32_t*>(d_atom_entries.ptr), num_active,
          static_cast<const int32_t*>(d_entry_group.ptr),
          static_cast<const int32_t*>(d_entry_seg.ptr), static_cast<const double*>(d_ref.ptr),
          solution, seg_centroid, group_factor, num_groups, nseg, grad, accumulate ? 1 : 0, total);
      timer.end(st);
      CVP_LAUNCH_CHECK();
    }
    return CVP_OK;
  }
  int run(const EvalArgs& a, void* ws) override {
    return a.dtype == CVP_F32 ? run_type<float>(a, ws) : run_type<double>(a, ws);
  }
};

}  // namespace

static bool B_blocks_fit(const RmsdCV& cv) {
  // (FP32 coordinates, a shared reference of up to 64 entries; checked again per dtype at run time)
  return ((size_t)cv.num_groups * BLOCKS_SOL + 64 * 3) * sizeof(double) +
             cv.active_atoms.size() * 3 * sizeof(float) <=
         (size_t)BLOCKS_SMEM_BUDGET;
}

int rmsd_create(const cvp_rmsd_desc* desc, cvp_cv** out) {
  CVP_REQUIRE(desc != nullptr && out != nullptr, "null argument");
  CVP_REQUIRE(desc->num_atoms > 0 && desc->num_groups > 0 && desc->group_offsets && desc->atoms &&
                  desc->reference,
              "empty RMSD description");
  CVP_REQUIRE(desc->combine >= CVP_COMBINE_IDENTITY && desc->combine <= CVP_COMBINE_PATH_DEVIATION,
              "unknown combine rule");
  CVP_REQUIRE(desc->combine != CVP_COMBINE_IDENTITY || desc->num_groups == 1,
              "identity combine needs exactly one group");
  if (desc->combine >= CVP_COMBINE_PATH_PROGRESS)
    CVP_REQUIRE(desc->num_groups >= 2 && desc->sigma > 0, "path needs >= 2 groups and sigma > 0");
  if (desc->combine == CVP_COMBINE_STEP_SUM)
    CVP_REQUIRE(desc->r0 > 0 && desc->step_kind >= CVP_STEP_RATIONAL &&
                    desc->step_kind <= CVP_STEP_PLUMED,
                "bad step function");
  const int G = desc->num_groups;
  const int64_t E = desc->group_offsets[G];
  CVP_REQUIRE(desc->group_offsets[0] == 0 && E > 0 && E < (int64_t(1) << 31), "bad group offsets");

  auto cv = new RmsdCV();
  cv->num_atoms = desc->num_atoms;
  cv->num_groups = G;
  cv->num_entries = E;
  cv->cp.mode = desc->combine;
  cv->cp.step.kind = desc->step_kind;
  cv->cp.step.n = (int)desc->step_p0;
  cv->cp.step.m = (int)desc->step_p1;
  cv->cp.inv_r0 = desc->r0 > 0 ? 1.0 / desc->r0 : 1.0;
  cv->cp.scale = desc->scale;
  cv->cp.lambda = desc->sigma > 0 ? 1.0 / (2.0 * desc->sigma * desc->sigma) : 0.0;

  auto fail = [&](const char* message) {
    delete cv;
    set_error(message);
    return CVP_ERR_INVALID;
  };
  cv->group_seg_offsets.push_back(0);
  cv->seg_slice_offsets.push_back(0);
  std::vector<char> seen(desc->num_atoms, 0);
  for (int g = 0; g < G; ++g) {
    const int begin = desc->group_offsets[g], end = desc->group_offsets[g + 1];
    if (end <= begin) return fail("empty RMSD group");
    // stable order of entries by body id
    std::vector<int32_t> order(end - begin);
    for (int k = 0; k < end - begin; ++k) order[k] = begin + k;
    if (desc->body_ids)
      std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
        return desc->body_ids[x] < desc->body_ids[y];
      });
    std::fill(seen.begin(), seen.end(), 0);
    double gy = 0.0;
    size_t k = 0;
    while (k < order.size()) {
      const int body = desc->body_ids ? desc->body_ids[order[k]] : 0;
      size_t k2 = k;
      double c[3] = {0, 0, 0};
      while (k2 < order.size() && (desc->body_ids ? desc->body_ids[order[k2]] : 0) == body) {
        for (int d = 0; d < 3; ++d) c[d] += desc->reference[3 * (size_t)order[k2] + d];
        ++k2;
      }
      const int count = (int)(k2 - k);
      for (int d = 0; d < 3; ++d) c[d] /= count;
      const int seg = (int)cv->seg_size.size();
      const int seg_begin = (int)cv->entry_atom.size();
      for (size_t m = k; m < k2; ++m) {
        const int32_t atom = desc->atoms[order[m]];
        if (atom < 0 || atom >= desc->num_atoms) return fail("atom index out of range in RMSD group");
        if (seen[atom]) return fail("repeated atom in RMSD group");
        seen[atom] = 1;
        cv->entry_atom.push_back(atom);
        cv->entry_group.push_back(g);
        cv->entry_seg.push_back(seg);
        for (int d = 0; d < 3; ++d) {
          const double y = desc->reference[3 * (size_t)order[m] + d] - c[d];
          cv->ref.push_back(y);
          gy += y * y;
        }
      }
      cv->seg_size.push_back(count);
      for (int s = seg_begin; s < seg_begin + count; s += RMSD_SLICE)
        cv->slices.push_back({s, std::min(s + RMSD_SLICE, seg_begin + count)});
      cv->seg_slice_offsets.push_back((int32_t)cv->slices.size());
      k = k2;
    }
    cv->group_seg_offsets.push_back((int32_t)cv->seg_size.size());
    cv->group_size.push_back(end - begin);
    cv->group_gy.push_back(gy);
  }
  cv->nseg = (int)cv->seg_size.size();
  cv->nslice = (int)cv->slices.size();
  // atom -> entries (CSR over sorted active atoms)
  {
    std::vector<std::vector<int32_t>> by_atom(desc->num_atoms);
    for (int64_t e = 0; e < E; ++e) by_atom[cv->entry_atom[e]].push_back((int32_t)e);
    cv->atom_entry_offsets.push_back(0);
    for (int a = 0; a < desc->num_atoms; ++a) {
      if (by_atom[a].empty()) continue;
      cv->active_atoms.push_back(a);
      for (int32_t e : by_atom[a]) cv->atom_entries.push_back(e);
      cv->atom_entry_offsets.push_back((int32_t)cv->atom_entries.size());
    }
    cv->dense = (int)cv->active_atoms.size() == desc->num_atoms;
  }
  // streaming fast path: a single group/segment whose entries are atoms first, first+1, ... with a
  // multiple-of-4 count and a 16-byte aligned start within the frame
  if (G == 1 && cv->nseg == 1 && E % 4 == 0) {
    bool contiguous = true;
    for (int64_t e = 0; e < E && contiguous; ++e)
      contiguous = cv->entry_atom[e] == cv->entry_atom[0] + (int32_t)e;
    if (contiguous && cv->entry_atom[0] % 4 == 0) {
      cv->fast = true;
      cv->fast_first = cv->entry_atom[0];
      cv->fast_nslice = (int)div_up(E, RMSD_FAST_SLICE);
      // the sums use the centred reference in double (see rmsd_sums_fast_kernel), the FP32
      // gradient its float copy
      cv->ref_f32.resize((size_t)E * 3);
      for (size_t k = 0; k < cv->ref_f32.size(); ++k) cv->ref_f32[k] = (float)cv->ref[k];
      cv->ref_perm = fused_permute_doubles(cv->ref.data(), (size_t)E);
      cv->fast_seg_slice_offsets = {0, cv->fast_nslice};
    }
  }
  // many small single-body groups summed through a step function (helix / sheet content): one CTA
  // per frame when the frame's active atoms and the per-group constants fit in shared memory
  if (desc->combine == CVP_COMBINE_STEP_SUM && cv->nseg == G && B_blocks_fit(*cv)) {
    cv->blocks = true;
    std::vector<int32_t> slot_of_atom(desc->num_atoms, -1);
    for (size_t k = 0; k < cv->active_atoms.size(); ++k) slot_of_atom[cv->active_atoms[k]] = (int32_t)k;
    cv->entry_slot.resize((size_t)E);
    for (int64_t e = 0; e < E; ++e) cv->entry_slot[e] = slot_of_atom[cv->entry_atom[e]];
    cv->group_entry_offsets.assign(1, 0);
    for (int g = 0; g < G; ++g)
      cv->group_entry_offsets.push_back(cv->group_entry_offsets.back() + cv->group_size[g]);
    // one centred reference shared by all groups (bitwise), as the content CVs build them?
    const int n0 = cv->group_size[0];
    bool shared = n0 <= 64;
    for (int g = 1; g < G && shared; ++g) {
      shared = cv->group_size[g] == n0 &&
               std::equal(cv->ref.begin(), cv->ref.begin() + 3 * (size_t)n0,
                          cv->ref.begin() + 3 * (size_t)cv->group_entry_offsets[g]);
    }
    cv->blocks_shared_ref = shared ? n0 : 0;
  }
  // many references, one atom set: G >= 2 single-segment groups that all list atoms first,
  // first + 1, ..., first + n - 1 in this order
  if (G >= 2 && cv->nseg == G && E % G == 0) {
    const int64_t n = E / G;
    // (the covariance kernel addresses a tile's rows with 32-bit offsets: 32 rows of 3 N values)
    bool same = n >= 64 && (int64_t)desc->num_atoms * 3 * MULTI_TF < (int64_t(1) << 32);
    for (int g = 0; g < G && same; ++g) same = desc->group_offsets[g] == g * n;
    for (int64_t e = 0; e < E && same; ++e)
      same = cv->entry_atom[e] == cv->entry_atom[0] + (int32_t)(e % n);
    if (same) {
      cv->multi = true;
      cv->multi_first = cv->entry_atom[0];
      cv->multi_n = (int)n;
      // slices over the atoms: at least 512 atoms each, never more than the generic path reserves
      // workspace for (ceil(n / 1024) per group)
      cv->multi_sk_max = (int)std::max<int64_t>(1, std::min<int64_t>({16, n / 512, div_up(n, 1024)}));
      pk_pack_reference(cv->ref.data(), G, (int)n, cv->ref_pack);
      cpk_pack_reference(cv->ref.data(), G, (int)n, cv->ref_pack_cov);
    }
  }
  *out = cv;
  return CVP_OK;
}

}  // namespace cvp

#ifdef FUSED_TIMING
// diagnostics build only (make EXTRA=-DFUSED_TIMING): where the RMSD instantiation of
// stream_fused_kernel writes its per-frame timestamps ([frames][6] unsigned long long, device memory
// owned by the caller; scripts/fused_timing.py)
extern "C" int cvp_debug_set_fused_timing(void* device_buffer) {
  CVP_CUDA_CHECK(cudaMemcpyToSymbol(cvp::g_fused_timing, &device_buffer, sizeof(void*)));
  return CVP_OK;
}
#endif
