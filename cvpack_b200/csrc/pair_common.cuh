// Device-side building blocks shared by the pair-sum kernels: per-pair functions (energy and
// (dE/dr)/r), switching, minimum image, tags.
#pragma once

#include "common.cuh"

namespace cvp {
namespace {

template <typename T>
struct PairConsts {
  T cutoff2;    // squared cutoff (huge if none)
  T rs;         // switching distance
  T inv_range;  // 1/(rc - rs)
  T neg30_inv_range;  // -30/(rc - rs)
  int has_switch;
  T inv_r0;
  T inv_r0sq;   // 1/r0^2
  StepParams step;
  T inv_rc;     // attraction: x = r/rc
  T ke;         // attraction: Coulomb constant
  double neg_beta_over_rc;  // softmin
};

// ---- tags: atom index (+ "member of both groups" flag) carried in the w lane, bit-exact ---------
__device__ __forceinline__ float tag_encode(int tag, float) { return __int_as_float(tag); }
__device__ __forceinline__ double tag_encode(int tag, double) {
  return __longlong_as_double((long long)tag);
}
__device__ __forceinline__ int tag_decode(float w) { return __float_as_int(w); }
__device__ __forceinline__ int tag_decode(double w) { return (int)__double_as_longlong(w); }
constexpr int TAG_BOTH = 0x40000000;  // set when the atom belongs to both groups
constexpr int TAG_INDEX_MASK = 0x3fffffff;

template <typename T>
__device__ __forceinline__ T round_nearest(T x);
template <>
__device__ __forceinline__ float round_nearest<float>(float x) {
  // round-to-nearest-even through the 1.5*2^23 magic constant: two full-rate FADDs instead of
  // the quarter-rate FRND; valid for |x| < 2^22, far beyond any image count
  return __fadd_rn(__fadd_rn(x, 12582912.0f), -12582912.0f);
}
template <>
__device__ __forceinline__ double round_nearest<double>(double x) {
  return rint(x);
}

template <int PBC, typename T>
__device__ __forceinline__ void pair_min_image(T& dx, T& dy, T& dz, const Box<T>& box) {
  if (PBC == 1) {
    dx = fma(-box.ax, round_nearest(dx * box.inv_ax), dx);
    dy = fma(-box.by, round_nearest(dy * box.inv_by), dy);
    dz = fma(-box.cz, round_nearest(dz * box.inv_cz), dz);
  } else if (PBC == 2) {
    min_image<2>(dx, dy, dz, box);
  }
}

template <typename T>
__device__ __forceinline__ void switching(const PairConsts<T>& pc, T r, T& e, T& de) {
  if (pc.has_switch && r > pc.rs) {
    T t = (r - pc.rs) * pc.inv_range;
    T sw = T(1) + t * t * t * (T(-10) + t * (T(15) - T(6) * t));
    T dsw = t * t * (T(-30) + t * (T(60) - T(30) * t)) * pc.inv_range;
    de = de * sw + e * dsw;
    e = e * sw;
  }
}

// single-instruction reciprocal / reciprocal square root (MUFU); ~2 ulp, far inside the 1e-5 budget
__device__ __forceinline__ float fast_rsqrt(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ double fast_rsqrt(double x) { return 1.0 / sqrt(x); }
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ double fast_rcp(double x) { return 1.0 / x; }

// NumberOfContacts with the default step function 1/(1+x^6): everything but the switching needs
// only r^2 (x^6 = (r^2/r0^2)^3); branch-free switching.
// SWITCH_ALWAYS: evaluate the switching polynomial unconditionally -- without a switching function
// the host sets rs = 1/(rc - rs) = 0, so t = 0, sw = 1, dsw = 0 and the result is bitwise the same
template <typename T, bool SWITCH_ALWAYS = false>
__device__ __forceinline__ void contacts_rational6(const PairConsts<T>& pc, T r2, T& e, T& coef) {
  // r2 == 0 (coincident atoms): S(0) = 1 and no force, like the reference; avoids 0 * inf
  const T inv_r = r2 > T(0) ? fast_rsqrt(r2) : T(0);
  const T x2 = r2 * pc.inv_r0sq;
  const T x6 = x2 * x2 * x2;
  const T s = fast_rcp(T(1) + x6);
  const T inv_r2 = inv_r * inv_r;
  T c = T(-6) * x6 * s * s * inv_r2;  // (dS/dr)/r
  T en = s;
  if (SWITCH_ALWAYS || pc.has_switch) {
    // sw = 1 - t^3 (10 - 15 t + 6 t^2),  dsw/dr = -30 t^2 (1 - t)^2 / (rc - rs),  t clamped at 0
    const T r = r2 * inv_r;
    const T t = fmax((r - pc.rs) * pc.inv_range, T(0));
    const T tu = t * (T(1) - t);
    const T sw = fma(-(t * t * t), fma(t, fma(T(6), t, T(-15)), T(10)), T(1));
    const T dsw = pc.neg30_inv_range * tu * tu;
    c = fma(c, sw, s * dsw * inv_r);
    en = s * sw;
  }
  e = en;
  coef = c;
}

// ---- pair functions: energy e and coef = (de/dr)/r for a pair at squared distance r2 -----------
template <typename T, int KIND>
struct PairFn;

template <typename T>
struct PairFn<T, CVP_PAIR_CONTACTS> {
  using T4 = typename Vec4<T>::type;
  using Acc = T;  // type of the energy channels and of the per-frame coefficients
  static constexpr int NCH = 1;
  static constexpr bool HAS_PARAMS = false;
  __device__ static __forceinline__ void eval(const PairConsts<T>& pc, T r2, const T4&, const T4&,
                                              const T*, T (&e)[1], T& coef) {
    if (r2 == T(0)) {
      // coincident atoms (or a residue listed in both groups of ResidueCoordination): the reference
      // evaluates S(0) and a vanishing force; rsqrt(0) * 0 would be NaN
      T s0, ds0;
      step_function(pc.step, T(0), s0, ds0);
      e[0] = s0;
      coef = T(0);
      return;
    }
    T inv_r = fast_rsqrt(r2);
    T r = r2 * inv_r;
    T s, ds;
    step_function(pc.step, r * pc.inv_r0, s, ds);
    T de = ds * pc.inv_r0;
    switching(pc, r, s, de);
    e[0] = s;
    coef = de * inv_r;
  }
};

template <typename T>
struct PairFn<T, CVP_PAIR_ATTRACTION> {
  using T4 = typename Vec4<T>::type;
  using Acc = T;
  static constexpr int NCH = 1;
  static constexpr bool HAS_PARAMS = true;
  // params: x = charge, y = sigma, z = epsilon, w = sign
  __device__ static __forceinline__ void eval(const PairConsts<T>& pc, T r2, const T4& pi,
                                              const T4& pj, const T*, T (&e)[1], T& coef) {
    T inv_r = rsqrt_t(r2);
    T r = r2 * inv_r;
    T sigma = T(0.5) * (pi.y + pj.y);
    T epsilon = sqrt_t(pi.z * pj.z);
    T qq = fmin(T(0), pi.x * pj.x);
    T sgn = pi.w * pj.w;
    T u = r / sigma;
    T u2 = u * u;
    T u6 = u2 * u2 * u2;
    T a = u6 - T(2);
    T y = fabs(a) + T(2);
    T inv_y = T(1) / y;
    T lj = T(4) * epsilon * (inv_y * inv_y - inv_y);
    // d|a|/da = +1 at a == 0 (Lepton: 2*step(a)-1)
    T dy = (a >= T(0) ? T(6) : T(-6)) * u6 * inv_r;
    T dlj = T(4) * epsilon * (T(-2) * inv_y * inv_y * inv_y + inv_y * inv_y) * dy;
    T x = r * pc.inv_rc;
    T coul = pc.ke * qq * (T(1) / x + T(0.5) * (x * x - T(3)));
    T dcoul = pc.ke * qq * (T(-1) / (x * x) + x) * pc.inv_rc;
    T en = -sgn * (lj + coul);
    T de = -sgn * (dlj + dcoul);
    switching(pc, r, en, de);
    e[0] = en;
    coef = de * inv_r;
  }
};

template <typename T>
struct PairFn<T, CVP_PAIR_SOFTMIN> {
  using T4 = typename Vec4<T>::type;
  // The two channels sum r w and sum w, w = exp(-beta r/rc), and the per-frame quotient-rule
  // coefficients {1/D, -rmin/D} are carried in double whatever T is: with the default beta = 100
  // w spans e^-100 .. 1 and 1/D reaches e^100 -- beyond the range of FP32 -- while the products
  // w/D <= 1 that enter the gradient are harmless.
  using Acc = double;
  static constexpr int NCH = 2;
  static constexpr bool HAS_PARAMS = false;
  // With beta = 100 the weight changes by 1e-5 when r changes by 1e-7 r: the squared distance is
  // formed in double from the (exact) coordinate differences -- three FP32 roundings in r^2 alone
  // would use up the whole FP32-mode error budget.
  __device__ static __forceinline__ void eval_d(const PairConsts<T>& pc, double x, double y, double z,
                                                const double* c, double (&e)[2], T& coef) {
    eval_r(pc, sqrt(x * x + y * y + z * z), c, e, coef);
  }
  __device__ static __forceinline__ void eval(const PairConsts<T>& pc, T r2, const T4&, const T4&,
                                              const double* c, double (&e)[2], T& coef) {
    eval_r(pc, sqrt((double)r2), c, e, coef);
  }
  __device__ static __forceinline__ void eval_r(const PairConsts<T>& pc, double r, const double* c,
                                                double (&e)[2], T& coef) {
    double w = exp(pc.neg_beta_over_rc * r);
    e[0] = r * w;
    e[1] = w;
    if (c != nullptr && r > 0.0) {
      double dw = pc.neg_beta_over_rc * w;
      coef = T((c[0] * (w + r * dw) + c[1] * dw) / r);
    } else {
      coef = T(0);
    }
  }
};

}  // namespace
}  // namespace cvp
