// Sums of bonded terms (distances, angles, dihedrals, pairs of dihedrals), value + gradient, for a
// batch of frames: Distance, Angle, Torsion, Helix{Torsion,Angle,HBond}Content, TorsionSimilarity.
//
// Replaces CustomBondForce / CustomAngleForce / CustomTorsionForce with one expression for all
// terms (cvpack/distance.py:57-59, angle.py:73-75, torsion.py:80-82,
// helix_torsion_content.py:126-145, helix_angle_content.py:118-128, helix_hbond_content.py:104-114)
// and the 8-particle CustomCompoundBondForce of torsion_similarity.py:89-93.
//
// Stages (one stream, deterministic):
//   terms     one thread per (frame, term): q, f(q) and the gradient on the term's 2-4 atoms
//   value     one warp per frame: fixed-order sum of the term values
//   gradient  one thread per (frame, active atom): sums the term slots it belongs to (CSR)
#include "common.cuh"

namespace cvp {
namespace {

struct TermParams {
  int function;
  double inv_tolerance;
  int exponent;  // 2m
  double coefficient;
};

template <int PBC>
__device__ __forceinline__ void delta(const double* a, const double* b, const Box<double>& box,
                                      double (&d)[3]) {
  d[0] = a[0] - b[0];
  d[1] = a[1] - b[1];
  d[2] = a[2] - b[2];
  min_image<PBC>(d[0], d[1], d[2], box);
}
__device__ __forceinline__ double dot3(const double (&a)[3], const double (&b)[3]) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
__device__ __forceinline__ void cross3(const double (&a)[3], const double (&b)[3], double (&c)[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// dihedral of four atoms and its gradient; sign convention of cvpack/torsion.py:22-33
template <int PBC>
__device__ __forceinline__ double torsion_geometry(const double (&r0)[3], const double (&r1)[3],
                                                   const double (&r2)[3], const double (&r3)[3],
                                                   const Box<double>& box, double (&g0)[3],
                                                   double (&g1)[3], double (&g2)[3],
                                                   double (&g3)[3]) {
  double f[3], gv[3], h[3], fxh[3], a[3], b[3];
  delta<PBC>(r0, r1, box, f);
  delta<PBC>(r2, r1, box, gv);
  delta<PBC>(r3, r2, box, h);
  const double gg = dot3(gv, gv);
  const double ng = sqrt(gg);
  cross3(f, h, fxh);
  const double fu = dot3(f, gv) / ng, hu = dot3(h, gv) / ng;
  const double q = atan2(dot3(fxh, gv) / ng, dot3(f, h) - fu * hu);
  cross3(f, gv, a);
  cross3(h, gv, b);
  const double ca = ng / dot3(a, a), cb = -ng / dot3(b, b);
  const double pf = dot3(f, gv) / gg, ph = dot3(h, gv) / gg;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double d1 = ca * a[c], d4 = cb * b[c];
    g0[c] = d1;
    g3[c] = d4;
    g1[c] = -d1 + pf * d1 + ph * d4;
    g2[c] = -d4 - pf * d1 - ph * d4;
  }
  return q;
}

// q and dq/dr for the A atoms of one term; coordinates promoted to double (a handful of terms per
// frame: accuracy is free, and acos/atan2 near their singular points need it).
template <int A, int PBC>
__device__ __forceinline__ double term_geometry(const double (&r)[A][3], const Box<double>& box,
                                                double (&g)[A][3]) {
  if (A == 2) {
    double d[3];
    delta<PBC>(r[1], r[0], box, d);
    const double q = sqrt(dot3(d, d));
    const double inv = 1.0 / q;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      g[1][c] = d[c] * inv;
      g[0][c] = -d[c] * inv;
    }
    return q;
  } else if (A == 3) {
    double u[3], v[3];
    delta<PBC>(r[0], r[1], box, u);
    delta<PBC>(r[2], r[1], box, v);
    const double nu = sqrt(dot3(u, u)), nv = sqrt(dot3(v, v));
    double cosine = dot3(u, v) / (nu * nv);
    cosine = fmin(1.0, fmax(-1.0, cosine));
    const double sine = sqrt(fmax(1.0 - cosine * cosine, 1e-300));
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double g1 = -(v[c] / (nu * nv) - cosine * u[c] / (nu * nu)) / sine;
      const double g3 = -(u[c] / (nu * nv) - cosine * v[c] / (nv * nv)) / sine;
      g[0][c] = g1;
      g[2][c] = g3;
      g[1][c] = -(g1 + g3);
    }
    return acos(cosine);
  } else if (A == 4) {
    return torsion_geometry<PBC>(r[0], r[1], r[2], r[3], box, g[0], g[1], g[2], g[3]);
  } else {
    // A == 8: q = dihedral(p1..p4) - dihedral(p5..p8)  (torsion_similarity.py:90)
    const double q1 = torsion_geometry<PBC>(r[0], r[1], r[2], r[3], box, g[0], g[1], g[2], g[3]);
    const double q2 = torsion_geometry<PBC>(r[4], r[5], r[6], r[7], box, g[4], g[5], g[6], g[7]);
#pragma unroll
    for (int a = 4; a < 8; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) g[a][c] = -g[a][c];
    return q1 - q2;
  }
}

template <typename T, int A, int PBC>
__global__ void bonded_terms_kernel(const T* __restrict__ pos, int64_t num_atoms,
                                    const int32_t* __restrict__ term_atoms,
                                    const double* __restrict__ reference, int num_terms,
                                    TermParams tp, const T* __restrict__ box, int box_mode,
                                    double* __restrict__ term_value, T* __restrict__ term_grad,
                                    int64_t total) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int64_t b = t / num_terms;
  const int k = (int)(t - b * num_terms);
  Box<double> bx;
  if (PBC != 0) {
    Box<T> s = load_box(box, box_mode, b);
    bx.ax = s.ax; bx.bx = s.bx; bx.by = s.by; bx.cx = s.cx; bx.cy = s.cy; bx.cz = s.cz;
    bx.inv_ax = 1.0 / bx.ax; bx.inv_by = 1.0 / bx.by; bx.inv_cz = 1.0 / bx.cz;
  }
  double r[A][3], g[A][3];
#pragma unroll
  for (int a = 0; a < A; ++a) {
    const T* p = pos + (b * num_atoms + term_atoms[k * A + a]) * 3;
    r[a][0] = p[0];
    r[a][1] = p[1];
    r[a][2] = p[2];
  }
  const double q = term_geometry<A, PBC>(r, bx, g);
  double f, df;
  if (tp.function == CVP_TERM_IDENTITY) {
    f = tp.coefficient * q;
    df = tp.coefficient;
  } else if (tp.function == CVP_TERM_HALF_COSINE) {
    // 0.5 (1 + cos(min(q, 2 pi - q))): cos is even and periodic, and both branches of the min have
    // the derivative -sin q, so the expression of torsion_similarity.py:89 is 0.5 (1 + cos q)
    f = tp.coefficient * 0.5 * (1.0 + cos(q));
    df = -tp.coefficient * 0.5 * sin(q);
  } else {
    const double x = (q - reference[k]) * tp.inv_tolerance;
    const double xp1 = powi(x, tp.exponent - 1);
    const double s = 1.0 / (1.0 + xp1 * x);
    f = tp.coefficient * s;
    df = -tp.coefficient * tp.exponent * xp1 * s * s * tp.inv_tolerance;
  }
  term_value[t] = f;
  T* out = term_grad + t * (A * 3);
#pragma unroll
  for (int a = 0; a < A; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c) out[a * 3 + c] = T(df * g[a][c]);
}

template <typename T>
__global__ void bonded_value_kernel(const double* __restrict__ term_value, int num_terms,
                                    T* __restrict__ value, int accumulate, int64_t num_frames) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (b >= num_frames) return;
  double v = 0.0;
  for (int k = lane; k < num_terms; k += 32) v += term_value[b * num_terms + k];
  v = warp_sum(v);
  if (lane == 0) value[b] = accumulate ? T((double)value[b] + v) : T(v);
}

template <typename T>
__global__ void bonded_gradient_kernel(const T* __restrict__ term_grad, int slots_per_frame,
                                       const int32_t* __restrict__ active_atom,
                                       const int32_t* __restrict__ atom_slot_offsets,
                                       const int32_t* __restrict__ atom_slots, int num_active,
                                       int64_t num_atoms, T* __restrict__ grad, int accumulate,
                                       int64_t total) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int64_t b = t / num_active;
  const int k = (int)(t - b * num_active);
  const T* frame = term_grad + b * (int64_t)slots_per_frame * 3;
  T gx = 0, gy = 0, gz = 0;
  for (int m = atom_slot_offsets[k]; m < atom_slot_offsets[k + 1]; ++m) {
    const T* s = frame + (int64_t)atom_slots[m] * 3;
    gx += s[0];
    gy += s[1];
    gz += s[2];
  }
  T* out = grad + (b * num_atoms + active_atom[k]) * 3;
  if (accumulate) {
    out[0] += gx;
    out[1] += gy;
    out[2] += gz;
  } else {
    out[0] = gx;
    out[1] = gy;
    out[2] = gz;
  }
}

struct BondedCV : cvp_cv {
  int num_terms = 0, atoms_per_term = 0, periodic = 0;
  TermParams tp{};
  std::vector<int32_t> term_atoms, atom_slot_offsets, atom_slots;
  std::vector<double> reference;
  DeviceBuffer d_term_atoms, d_reference, d_active, d_atom_slot_offsets, d_atom_slots;
  ~BondedCV() override {
    for (DeviceBuffer* buf :
         {&d_term_atoms, &d_reference, &d_active, &d_atom_slot_offsets, &d_atom_slots})
      buf->release();
  }
  int upload() override {
    int status;
    if ((status = cvp::upload(d_term_atoms, term_atoms)) != CVP_OK) return status;
    if ((status = cvp::upload(d_reference, reference)) != CVP_OK) return status;
    if ((status = cvp::upload(d_active, active_atoms)) != CVP_OK) return status;
    if ((status = cvp::upload(d_atom_slot_offsets, atom_slot_offsets)) != CVP_OK) return status;
    return cvp::upload(d_atom_slots, atom_slots);
  }
  size_t workspace_per_frame(int dtype, bool) const override {
    size_t real = dtype == CVP_F32 ? 4 : 8;
    return (size_t)num_terms * (sizeof(double) + atoms_per_term * 3 * real) + 64;
  }
  size_t workspace_fixed(int) const override { return 4 * 256; }

  template <typename T, int A, int PBC>
  int run_impl(const EvalArgs& a, void* ws) {
    const int64_t B = a.num_frames, N = a.num_atoms;
    const bool accumulate = (a.flags & CVP_FLAG_ACCUMULATE) != 0;
    const T* pos = static_cast<const T*>(a.positions);
    const T* box = static_cast<const T*>(a.box);
    T* value = static_cast<T*>(a.value);
    T* grad = static_cast<T*>(a.grad);
    cudaStream_t st = a.stream;
    Carver carve(ws);
    double* term_value = carve.take<double>((size_t)B * num_terms);
    T* term_grad = carve.take<T>((size_t)B * num_terms * A * 3);
    int64_t total = B * (int64_t)num_terms;
    bonded_terms_kernel<T, A, PBC><<<(unsigned)div_up(total, 128), 128, 0, st>>>(
        pos, N, static_cast<const int32_t*>(d_term_atoms.ptr),
        static_cast<const double*>(d_reference.ptr), num_terms, tp, box, a.box_mode, term_value,
        term_grad, total);
    CVP_LAUNCH_CHECK();
    bonded_value_kernel<T><<<(unsigned)div_up(B * 32, 128), 128, 0, st>>>(
        term_value, num_terms, value, accumulate ? 1 : 0, B);
    CVP_LAUNCH_CHECK();
    if (grad == nullptr) return CVP_OK;
    if (!accumulate) CVP_CUDA_CHECK(cudaMemsetAsync(grad, 0, (size_t)B * N * 3 * sizeof(T), st));
    const int num_active = (int)active_atoms.size();
    total = B * (int64_t)num_active;
    bonded_gradient_kernel<T><<<(unsigned)div_up(total, 128), 128, 0, st>>>(
        term_grad, num_terms * A, static_cast<const int32_t*>(d_active.ptr),
        static_cast<const int32_t*>(d_atom_slot_offsets.ptr),
        static_cast<const int32_t*>(d_atom_slots.ptr), num_active, N, grad, accumulate ? 1 : 0,
        total);
    CVP_LAUNCH_CHECK();
    return CVP_OK;
  }
  template <typename T, int A>
  int run_atoms(const EvalArgs& a, void* ws) {
    if (!periodic) return run_impl<T, A, 0>(a, ws);
    CVP_REQUIRE(a.box != nullptr && a.box_mode != CVP_BOX_NONE,
                "this collective variable uses periodic boundary conditions but no box was given");
    return (a.flags & CVP_FLAG_TRICLINIC) ? run_impl<T, A, 2>(a, ws) : run_impl<T, A, 1>(a, ws);
  }
  template <typename T>
  int run_type(const EvalArgs& a, void* ws) {
    if (atoms_per_term == 2) return run_atoms<T, 2>(a, ws);
    if (atoms_per_term == 3) return run_atoms<T, 3>(a, ws);
    if (atoms_per_term == 8) return run_atoms<T, 8>(a, ws);
    return run_atoms<T, 4>(a, ws);
  }
  int run(const EvalArgs& a, void* ws) override {
    return a.dtype == CVP_F32 ? run_type<float>(a, ws) : run_type<double>(a, ws);
  }
};

}  // namespace

int bonded_create(const cvp_bonded_desc* desc, cvp_cv** out) {
  CVP_REQUIRE(desc != nullptr && out != nullptr, "null argument");
  CVP_REQUIRE(desc->num_atoms > 0 && desc->num_terms > 0 && desc->atoms, "no bonded terms");
  CVP_REQUIRE((desc->atoms_per_term >= 2 && desc->atoms_per_term <= 4) || desc->atoms_per_term == 8,
              "bonded terms have 2, 3, 4 or (pairs of dihedrals) 8 atoms");
  CVP_REQUIRE(desc->function >= CVP_TERM_IDENTITY && desc->function <= CVP_TERM_HALF_COSINE,
              "unknown term function");
  CVP_REQUIRE((desc->atoms_per_term == 8) == (desc->function == CVP_TERM_HALF_COSINE),
              "the half-cosine function belongs to 8-atom terms (pairs of dihedrals)");
  if (desc->function == CVP_TERM_BOXCAR)
    CVP_REQUIRE(desc->reference && desc->tolerance > 0 && desc->half_exponent >= 1,
                "boxcar terms need references, a positive tolerance and an exponent >= 1");
  const int A = desc->atoms_per_term;
  auto cv = new BondedCV();
  cv->num_atoms = desc->num_atoms;
  cv->num_terms = desc->num_terms;
  cv->atoms_per_term = A;
  cv->periodic = desc->periodic;
  cv->tp.function = desc->function;
  cv->tp.inv_tolerance = desc->tolerance > 0 ? 1.0 / desc->tolerance : 1.0;
  cv->tp.exponent = 2 * desc->half_exponent;
  cv->tp.coefficient = desc->coefficient;
  cv->term_atoms.assign(desc->atoms, desc->atoms + (size_t)desc->num_terms * A);
  if (desc->reference)
    cv->reference.assign(desc->reference, desc->reference + desc->num_terms);
  else
    cv->reference.assign(desc->num_terms, 0.0);
  std::vector<std::vector<int32_t>> by_atom(desc->num_atoms);
  for (size_t s = 0; s < cv->term_atoms.size(); ++s) {
    int32_t atom = cv->term_atoms[s];
    if (atom < 0 || atom >= desc->num_atoms) {
      delete cv;
      set_error("atom index out of range in a bonded term");
      return CVP_ERR_INVALID;
    }
    by_atom[atom].push_back((int32_t)s);
  }
  cv->atom_slot_offsets.push_back(0);
  for (int a = 0; a < desc->num_atoms; ++a) {
    if (by_atom[a].empty()) continue;
    cv->active_atoms.push_back(a);
    for (int32_t s : by_atom[a]) cv->atom_slots.push_back(s);
    cv->atom_slot_offsets.push_back((int32_t)cv->atom_slots.size());
  }
  *out = cv;
  return CVP_OK;
}

}  // namespace cvp
