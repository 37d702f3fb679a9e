/*
 * CPU oracle, C restatement (double precision, OpenMP over frames) -- TEST INFRASTRUCTURE ONLY.
 *
 * Used as the checker at sizes where the NumPy oracle (oracle/cv_oracle.py) is too slow, and as
 * the timed CPU baseline of bench.py.  Nothing under cvpack_b200/ links or loads this file.
 *
 * What it restates: the per-frame evaluation OpenMM's *Reference* platform performs for the Force
 * objects cvpack configures -- OpenMM is a third-party dependency of the reference, unpinned
 * (pyproject.toml:23-27; CI 7.7/8.0/8.1) and absent from /root/reference and from this image, so
 * the algorithm is restated from its published behaviour and anchored on cvpack's call sites:
 *   - contacts:  CustomNonbondedForce with one interaction group, all |g1| x |g2| pairs visited,
 *                cutoff + switching, minimum image (cvpack/number_of_contacts.py:143-161)
 *   - rmsd:      RMSDForce: centroids, 3x3 correlation, Horn 4x4 eigenproblem (cvpack/rmsd.py:96)
 *   - gyration:  CustomCentroidBondForce distance(g_i, g_all)^2 terms
 *                (cvpack/radius_of_gyration_sq.py:87-89)
 * Pinned against oracle/cv_oracle.py (itself pinned by the reference's known answers, see
 * tests/test_oracle.py) in tests/test_oracle.py::test_c_oracle_matches_numpy.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static void min_image(double* d, const double* box) {
  /* box: 9 numbers, rows a, b, c in reduced form; reduce along c, then b, then a */
  if (!box) return;
  double s = floor(d[2] / box[8] + 0.5);
  d[0] -= s * box[6]; d[1] -= s * box[7]; d[2] -= s * box[8];
  s = floor(d[1] / box[4] + 0.5);
  d[0] -= s * box[3]; d[1] -= s * box[4];
  s = floor(d[0] / box[0] + 0.5);
  d[0] -= s * box[0];
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants every host thread */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* step_kind: 0 = 1/(1+x^n), 1 = step(1-x) */
static void step_function(int kind, int n, double x, double* s, double* ds) {
  if (kind == 1) {
    *s = x <= 1.0 ? 1.0 : 0.0;
    *ds = 0.0;
    return;
  }
  double xn1 = 1.0; /* x^(n-1) by repeated squaring (pow() dominated the in-range pair cost) */
  {
    double base = x;
    for (int e = n - 1; e > 0; e >>= 1) {
      if (e & 1) xn1 *= base;
      base *= base;
    }
  }
  double inv = 1.0 / (1.0 + xn1 * x);
  *s = inv;
  *ds = -n * xn1 * inv * inv;
}

/*
 * NumberOfContacts for `num_frames` frames.  in1/in2: membership flags per atom.  ex_offsets /
 * ex_partners: CSR list of excluded partners per atom (may be NULL).  rs <= 0: no switching.
 * value[B], grad[B][N][3] (grad may be NULL).
 */
void oracle_contacts_batch(const double* x, int64_t num_frames, int num_atoms, const int32_t* g1,
                           int n1, const int32_t* g2, int n2, const int32_t* ex_offsets,
                           const int32_t* ex_partners, const double* box, int box_per_frame,
                           double r0, double rc, double rs, int step_kind, int step_n,
                           double reference, double* value, double* grad) {
  char* in1 = calloc(num_atoms, 1);
  char* in2 = calloc(num_atoms, 1);
  for (int i = 0; i < n1; ++i) in1[g1[i]] = 1;
  for (int j = 0; j < n2; ++j) in2[g2[j]] = 1;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t b = 0; b < num_frames; ++b) {
    const double* xf = x + b * (int64_t)num_atoms * 3;
    const double* bx = box ? box + (box_per_frame ? b * 9 : 0) : NULL;
    double* gf = grad ? grad + b * (int64_t)num_atoms * 3 : NULL;
    if (gf) memset(gf, 0, sizeof(double) * 3 * num_atoms);
    double total = 0.0;
    for (int a = 0; a < n1; ++a) {
      const int i = g1[a];
      for (int c = 0; c < n2; ++c) {
        const int j = g2[c];
        if (i == j) continue;
        if (in1[j] && in2[i] && j < i) continue; /* unordered pair seen from both sides: once */
        if (ex_offsets) {
          int excluded = 0;
          for (int e = ex_offsets[i]; e < ex_offsets[i + 1]; ++e)
            if (ex_partners[e] == j) { excluded = 1; break; }
          if (excluded) continue;
        }
        double d[3] = {xf[3 * j] - xf[3 * i], xf[3 * j + 1] - xf[3 * i + 1],
                       xf[3 * j + 2] - xf[3 * i + 2]};
        min_image(d, bx);
        const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        if (r2 >= rc * rc) continue;
        const double r = sqrt(r2);
        double s, ds;
        step_function(step_kind, step_n, r / r0, &s, &ds);
        double de = ds / r0;
        if (rs > 0.0 && r > rs) {
          const double t = (r - rs) / (rc - rs);
          const double sw = 1.0 + t * t * t * (-10.0 + t * (15.0 - 6.0 * t));
          const double dsw = t * t * (-30.0 + t * (60.0 - 30.0 * t)) / (rc - rs);
          de = de * sw + s * dsw;
          s *= sw;
        }
        total += s;
        if (gf) {
          const double f = de / r / reference;
          for (int k = 0; k < 3; ++k) {
            gf[3 * j + k] += f * d[k];
            gf[3 * i + k] -= f * d[k];
          }
        }
      }
    }
    value[b] = total / reference;
  }
  free(in1);
  free(in2);
}

/* largest eigenpair of a symmetric 4x4 matrix by cyclic Jacobi rotations */
static void jacobi4(double a[4][4], double* lambda, double q[4]) {
  double v[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 50; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < 4; ++i)
      for (int j = i + 1; j < 4; ++j) off += a[i][j] * a[i][j];
    if (off < 1e-300) break;
    for (int p = 0; p < 3; ++p)
      for (int r = p + 1; r < 4; ++r) {
        if (a[p][r] == 0.0) continue;
        const double theta = (a[r][r] - a[p][p]) / (2.0 * a[p][r]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) {
          const double akp = a[k][p], akr = a[k][r];
          a[k][p] = c * akp - s * akr;
          a[k][r] = s * akp + c * akr;
        }
        for (int k = 0; k < 4; ++k) {
          const double apk = a[p][k], ark = a[r][k];
          a[p][k] = c * apk - s * ark;
          a[r][k] = s * apk + c * ark;
        }
        for (int k = 0; k < 4; ++k) {
          const double vkp = v[k][p], vkr = v[k][r];
          v[k][p] = c * vkp - s * vkr;
          v[k][r] = s * vkp + c * vkr;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < 4; ++i)
    if (a[i][i] > a[best][best]) best = i;
  *lambda = a[best][best];
  for (int k = 0; k < 4; ++k) q[k] = v[k][best];
}

/* RMSD of one group after optimal superposition, `num_frames` frames; ref[n][3] pairs with group */
void oracle_rmsd_batch(const double* x, int64_t num_frames, int num_atoms, const int32_t* group,
                       int n, const double* ref, double* value, double* grad) {
  double yc[3] = {0, 0, 0};
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) yc[k] += ref[3 * i + k] / n;
  double gy = 0.0;
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) gy += (ref[3 * i + k] - yc[k]) * (ref[3 * i + k] - yc[k]);
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < num_frames; ++b) {
    const double* xf = x + b * (int64_t)num_atoms * 3;
    double* gf = grad ? grad + b * (int64_t)num_atoms * 3 : NULL;
    if (gf) memset(gf, 0, sizeof(double) * 3 * num_atoms);
    double xc[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) xc[k] += xf[3 * group[i] + k];
    for (int k = 0; k < 3; ++k) xc[k] /= n;
    double r[3][3] = {{0}}, gx = 0.0;
    for (int i = 0; i < n; ++i) {
      double xh[3], yh[3];
      for (int k = 0; k < 3; ++k) {
        xh[k] = xf[3 * group[i] + k] - xc[k];
        yh[k] = ref[3 * i + k] - yc[k];
        gx += xh[k] * xh[k];
      }
      for (int c = 0; c < 3; ++c)
        for (int d = 0; d < 3; ++d) r[c][d] += xh[c] * yh[d];
    }
    double f[4][4] = {
        {r[0][0] + r[1][1] + r[2][2], r[1][2] - r[2][1], r[2][0] - r[0][2], r[0][1] - r[1][0]},
        {r[1][2] - r[2][1], r[0][0] - r[1][1] - r[2][2], r[0][1] + r[1][0], r[2][0] + r[0][2]},
        {r[2][0] - r[0][2], r[0][1] + r[1][0], -r[0][0] + r[1][1] - r[2][2], r[1][2] + r[2][1]},
        {r[0][1] - r[1][0], r[2][0] + r[0][2], r[1][2] + r[2][1], -r[0][0] - r[1][1] + r[2][2]}};
    double lambda, q[4];
    jacobi4(f, &lambda, q);
    const double msd = (gx + gy - 2.0 * lambda) / n;
    if (msd < 1e-20) {
      value[b] = 0.0;
      continue;
    }
    const double d = sqrt(msd);
    value[b] = d;
    if (!gf) continue;
    const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double q0 = q[0] / nq, q1 = q[1] / nq, q2 = q[2] / nq, q3 = q[3] / nq;
    const double u[3][3] = {
        {q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2)},
        {2 * (q1 * q2 + q0 * q3), q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2 * (q2 * q3 - q0 * q1)},
        {2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3}};
    for (int i = 0; i < n; ++i) {
      double yh[3];
      for (int k = 0; k < 3; ++k) yh[k] = ref[3 * i + k] - yc[k];
      for (int c = 0; c < 3; ++c) {
        /* (U^T y)_c = sum_d U[d][c] y_d */
        const double uty = u[0][c] * yh[0] + u[1][c] * yh[1] + u[2][c] * yh[2];
        gf[3 * group[i] + c] = (xf[3 * group[i] + c] - xc[c] - uty) / (n * d);
      }
    }
  }
}

/* radius of gyration (squared != 0: its square), weights[n] normalised, non-periodic */
void oracle_gyration_batch(const double* x, int64_t num_frames, int num_atoms,
                           const int32_t* group, int n, const double* weights, int squared,
                           double* value, double* grad) {
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < num_frames; ++b) {
    const double* xf = x + b * (int64_t)num_atoms * 3;
    double* gf = grad ? grad + b * (int64_t)num_atoms * 3 : NULL;
    if (gf) memset(gf, 0, sizeof(double) * 3 * num_atoms);
    double c[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) c[k] += weights[i] * xf[3 * group[i] + k];
    double s[3] = {0, 0, 0}, sum = 0.0;
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) {
        const double d = xf[3 * group[i] + k] - c[k];
        s[k] += d;
        sum += d * d;
      }
    const double rg2 = sum / n, rg = sqrt(rg2);
    value[b] = squared ? rg2 : rg;
    if (!gf) continue;
    const double factor = squared ? 2.0 / n : (rg > 0 ? 1.0 / (n * rg) : 0.0);
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k)
        gf[3 * group[i] + k] = factor * (xf[3 * group[i] + k] - c[k] - weights[i] * s[k]);
  }
}

/*
 * NumberOfContacts with a cell list -- the CPU analogue of what OpenMM's *CPU* platform does for a
 * CustomNonbondedForce with a cutoff (a neighbour list instead of the Reference platform's loop over
 * all |g1| x |g2| pairs).  Same pair set and arithmetic as oracle_contacts_batch (verified against
 * it in tests/test_oracle.py); rectangular periodic box or no box.  The group-2 atoms of a frame
 * are binned in cells of edge >= rc; every group-1 atom visits the 27 surrounding cells.  This is
 * the fair CPU arm of bench.py: O(N) per frame, OpenMP over frames.
 * Returns 0, or -1 if the geometry is not supported (triclinic box, fewer than 3 cells per axis
 * in a periodic box: callers fall back to oracle_contacts_batch).
 */
int oracle_contacts_cells_batch(const double* x, int64_t num_frames, int num_atoms,
                                const int32_t* g1, int n1, const int32_t* g2, int n2,
                                const int32_t* ex_offsets, const int32_t* ex_partners,
                                const double* box, int box_per_frame, double r0, double rc,
                                double rs, int step_kind, int step_n, double reference,
                                double* value, double* grad) {
  if (box) {
    const int64_t boxes = box_per_frame ? num_frames : 1;
    for (int64_t b = 0; b < boxes; ++b) {
      const double* bx = box + 9 * b;
      if (bx[1] != 0 || bx[2] != 0 || bx[3] != 0 || bx[5] != 0 || bx[6] != 0 || bx[7] != 0) return -1;
      if (bx[0] < 3 * rc || bx[4] < 3 * rc || bx[8] < 3 * rc) return -1;
    }
  }
  char* in1 = calloc(num_atoms, 1);
  char* in2 = calloc(num_atoms, 1);
  for (int i = 0; i < n1; ++i) in1[g1[i]] = 1;
  for (int j = 0; j < n2; ++j) in2[g2[j]] = 1;
#pragma omp parallel
  {
    int* cell_of = malloc(sizeof(int) * (size_t)n2);
    int* order = malloc(sizeof(int) * (size_t)n2);
    double* wrapped = malloc(sizeof(double) * 3 * (size_t)n2);
    int* start = NULL;
    int start_cap = 0;
#pragma omp for schedule(dynamic, 1)
    for (int64_t b = 0; b < num_frames; ++b) {
      const double* xf = x + b * (int64_t)num_atoms * 3;
      const double* bx = box ? box + (box_per_frame ? b * 9 : 0) : NULL;
      double* gf = grad ? grad + b * (int64_t)num_atoms * 3 : NULL;
      if (gf) memset(gf, 0, sizeof(double) * 3 * num_atoms);
      /* grid: the periodic box, or the bounding box of group 2 */
      double lo[3], len[3];
      if (bx) {
        lo[0] = lo[1] = lo[2] = 0.0;
        len[0] = bx[0]; len[1] = bx[4]; len[2] = bx[8];
      } else {
        for (int k = 0; k < 3; ++k) { lo[k] = 1e300; len[k] = -1e300; }
        for (int c = 0; c < n2; ++c)
          for (int k = 0; k < 3; ++k) {
            const double v = xf[3 * g2[c] + k];
            if (v < lo[k]) lo[k] = v;
            if (v > len[k]) len[k] = v;
          }
        for (int k = 0; k < 3; ++k) len[k] = fmax(len[k] - lo[k], 1e-9) * (1.0 + 1e-12) + 1e-12;
      }
      /* cells of edge >= rc/m, neighbours within +-m cells (m = 2 when the box allows it: the
         candidate volume drops from 27 (rc)^3-cells to the cells that can reach the cutoff sphere) */
      int dims[3], m = 2;
      for (int k = 0; k < 3; ++k)
        if (bx && (int)floor(len[k] / (0.5 * rc)) < 5) m = 1;
      for (int k = 0; k < 3; ++k) {
        dims[k] = (int)floor(len[k] / (rc / m));
        if (dims[k] < 1) dims[k] = 1;
        if (dims[k] > 128) dims[k] = 128;
      }
      double edge[3];
      for (int k = 0; k < 3; ++k) edge[k] = len[k] / dims[k];
      const int ncells = dims[0] * dims[1] * dims[2];
      if (ncells + 1 > start_cap) {
        free(start);
        start_cap = ncells + 1;
        start = malloc(sizeof(int) * (size_t)start_cap);
      }
      memset(start, 0, sizeof(int) * (size_t)(ncells + 1));
      for (int c = 0; c < n2; ++c) {
        int cc[3];
        for (int k = 0; k < 3; ++k) {
          double v = xf[3 * g2[c] + k];
          if (bx) v -= len[k] * floor(v / len[k]);
          wrapped[3 * c + k] = v;
          int q = (int)((v - lo[k]) / len[k] * dims[k]);
          cc[k] = q < 0 ? 0 : (q >= dims[k] ? dims[k] - 1 : q);
        }
        cell_of[c] = (cc[2] * dims[1] + cc[1]) * dims[0] + cc[0];
        start[cell_of[c] + 1]++;
      }
      for (int c = 0; c < ncells; ++c) start[c + 1] += start[c];
      /* stable counting sort: inside a cell the group order is kept, so sums have a fixed order */
      {
        int* fill = malloc(sizeof(int) * (size_t)ncells);
        memcpy(fill, start, sizeof(int) * (size_t)ncells);
        for (int c = 0; c < n2; ++c) order[fill[cell_of[c]]++] = c;
        free(fill);
      }
      double total = 0.0;
      for (int a = 0; a < n1; ++a) {
        const int i = g1[a];
        double xi[3];
        int ci[3];
        for (int k = 0; k < 3; ++k) {
          double v = xf[3 * i + k];
          if (bx) v -= len[k] * floor(v / len[k]);
          xi[k] = v;
          ci[k] = (int)floor((v - lo[k]) / len[k] * dims[k]);
          /* no box: an atom outside the bounding box of group 2 is clamped to the layer just
             outside it -- it then sees the border cells (extra candidates fail the distance test) */
          if (!bx) ci[k] = ci[k] < -m ? -m : (ci[k] > dims[k] + m - 1 ? dims[k] + m - 1 : ci[k]);
        }
        for (int dz = -m; dz <= m; ++dz)
          for (int dy = -m; dy <= m; ++dy)
            for (int dx = -m; dx <= m; ++dx) {
              int cc[3] = {ci[0] + dx, ci[1] + dy, ci[2] + dz};
              /* cells that cannot reach the cutoff sphere of any atom of the home cell */
              const double gx = (abs(dx) > 1 ? abs(dx) - 1 : 0) * edge[0];
              const double gy = (abs(dy) > 1 ? abs(dy) - 1 : 0) * edge[1];
              const double gz = (abs(dz) > 1 ? abs(dz) - 1 : 0) * edge[2];
              if (gx * gx + gy * gy + gz * gz >= rc * rc) continue;
              int skip = 0;
              for (int k = 0; k < 3; ++k) {
                if (bx) cc[k] = (cc[k] % dims[k] + dims[k]) % dims[k];  /* >= 2m+1 cells: all distinct */
                else if (cc[k] < 0 || cc[k] >= dims[k]) skip = 1;
              }
              if (skip) continue;
              const int cell = (cc[2] * dims[1] + cc[1]) * dims[0] + cc[0];
              for (int s = start[cell]; s < start[cell + 1]; ++s) {
                const int c = order[s];
                const int j = g2[c];
                if (i == j) continue;
                if (in1[j] && in2[i] && j < i) continue;
                if (ex_offsets) {
                  int excluded = 0;
                  for (int e = ex_offsets[i]; e < ex_offsets[i + 1]; ++e)
                    if (ex_partners[e] == j) { excluded = 1; break; }
                  if (excluded) continue;
                }
                double d[3];
                for (int k = 0; k < 3; ++k) {
                  d[k] = wrapped[3 * c + k] - xi[k];
                  if (bx) d[k] -= len[k] * floor(d[k] / len[k] + 0.5);
                }
                const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                if (r2 >= rc * rc) continue;
                const double r = sqrt(r2);
                double sv, ds;
                step_function(step_kind, step_n, r / r0, &sv, &ds);
                double de = ds / r0;
                if (rs > 0.0 && r > rs) {
                  const double t = (r - rs) / (rc - rs);
                  const double sw = 1.0 + t * t * t * (-10.0 + t * (15.0 - 6.0 * t));
                  const double dsw = t * t * (-30.0 + t * (60.0 - 30.0 * t)) / (rc - rs);
                  de = de * sw + sv * dsw;
                  sv *= sw;
                }
                total += sv;
                if (gf && r > 0.0) {
                  const double f = de / r / reference;
                  for (int k = 0; k < 3; ++k) {
                    gf[3 * j + k] += f * d[k];
                    gf[3 * i + k] -= f * d[k];
                  }
                }
              }
            }
      }
      value[b] = total / reference;
    }
    free(cell_of);
    free(order);
    free(wrapped);
    free(start);
  }
  free(in1);
  free(in2);
  return 0;
}
