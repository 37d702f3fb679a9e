// MetaCollectiveVariable: a function of other collective variables and named parameters.
//
// The reference hands the function string to an openmm.CustomCVForce (cvpack/
// meta_collective_variable.py:114-141): Lepton parses it, differentiates it symbolically with respect
// to every inner variable and evaluates value and partial derivatives once per frame.  Here the
// Python host compiles the same string into a short postfix program (cvpack_b200/expression.py) and
// this kernel interprets it for every frame of the batch with forward-mode derivatives:
//
//   expr_eval_kernel   one thread per frame: stack machine over (value, d/d input_k) pairs;
//                      inputs = inner CV values [ncv][B] followed by the named parameters
//
// value and d value / d cv_j then feed cvp_axpy_frames (path_cv.cu), which assembles the gradient.
// Derivative rules at kinks follow Lepton: step' = 0, abs'(x) = 2 step(x) - 1 (+1 at 0),
// min(a, b)' = b' when a >= b else a' (ties take the second argument), max(a, b)' = a' when a >= b,
// select differentiates the branch taken, floor' = ceil' = 0, a^b uses b a^(b-1) a' when b does not
// depend on the variable (valid for negative bases) and adds log(a) a^b b' when it does.
#include <math_constants.h>

#include "common.cuh"

namespace cvp {
namespace {

constexpr int EXPR_STACK = 24;

template <int ND>
struct ExprDual {
  double v;
  uint32_t dep;  // bit k: depends (structurally) on input k
  double d[ND];
};

template <int ND>
__device__ __forceinline__ void chain(ExprDual<ND>& x, double value, double slope) {
  x.v = value;
#pragma unroll
  for (int k = 0; k < ND; ++k) x.d[k] = (x.dep >> k) & 1u ? slope * x.d[k] : 0.0;
}

template <int ND>
__device__ __forceinline__ void combine(ExprDual<ND>& a, const ExprDual<ND>& b, double value, double fa,
                                        double fb) {
  a.v = value;
#pragma unroll
  for (int k = 0; k < ND; ++k) {
    // structural zeros stay exact zeros (inf * 0 must not poison a derivative that does not exist)
    const double ta = (a.dep >> k) & 1u ? fa * a.d[k] : 0.0;
    const double tb = (b.dep >> k) & 1u ? fb * b.d[k] : 0.0;
    a.d[k] = ta + tb;
  }
  a.dep |= b.dep;
}

template <typename T, int ND>
__global__ void expr_eval_kernel(cvp_expr_program prog, const T* __restrict__ cv_values /* [ncv][B] */,
                                 int64_t num_frames, T* __restrict__ out_value,
                                 T* __restrict__ out_partial /* [ninputs][B] or null */) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= num_frames) return;
  ExprDual<ND> stack[EXPR_STACK];
  ExprDual<ND> temp[CVP_EXPR_MAX_TEMPS];
  int sp = 0;
  const int ninputs = prog.num_variables + prog.num_parameters;
  for (int pc = 0; pc < prog.num_ops; ++pc) {
    const int op = prog.ops[pc].code;
    const int arg = prog.ops[pc].arg;
    if (op == CVP_EXPR_CONST) {
      ExprDual<ND>& x = stack[sp++];
      x.v = prog.ops[pc].value;
      x.dep = 0;
#pragma unroll
      for (int k = 0; k < ND; ++k) x.d[k] = 0.0;
    } else if (op == CVP_EXPR_INPUT) {
      ExprDual<ND>& x = stack[sp++];
      x.v = arg < prog.num_variables ? (double)cv_values[(int64_t)arg * num_frames + b]
                                     : prog.parameters[arg - prog.num_variables];
      x.dep = 1u << arg;
#pragma unroll
      for (int k = 0; k < ND; ++k) x.d[k] = k == arg ? 1.0 : 0.0;
    } else if (op == CVP_EXPR_LOAD) {
      stack[sp++] = temp[arg];
    } else if (op == CVP_EXPR_STORE) {
      temp[arg] = stack[--sp];
    } else if (op >= CVP_EXPR_ADD && op <= CVP_EXPR_ATAN2) {
      const ExprDual<ND>& y = stack[sp - 1];
      ExprDual<ND>& x = stack[sp - 2];
      --sp;
      switch (op) {
        case CVP_EXPR_ADD: combine(x, y, x.v + y.v, 1.0, 1.0); break;
        case CVP_EXPR_SUB: combine(x, y, x.v - y.v, 1.0, -1.0); break;
        case CVP_EXPR_MUL: combine(x, y, x.v * y.v, y.v, x.v); break;
        case CVP_EXPR_DIV: {
          const double inv = 1.0 / y.v;
          combine(x, y, x.v * inv, inv, -x.v * inv * inv);
          break;
        }
        case CVP_EXPR_POW: {
          const double value = pow(x.v, y.v);
          const double da = y.v == 0.0 ? 0.0 : y.v * pow(x.v, y.v - 1.0);
          const double db = y.dep ? value * log(x.v) : 0.0;
          combine(x, y, value, da, db);
          break;
        }
        case CVP_EXPR_MIN: {
          const double second = x.v - y.v >= 0.0 ? 1.0 : 0.0;
          combine(x, y, fmin(x.v, y.v), 1.0 - second, second);
          break;
        }
        case CVP_EXPR_MAX: {
          const double first = x.v - y.v >= 0.0 ? 1.0 : 0.0;
          combine(x, y, fmax(x.v, y.v), first, 1.0 - first);
          break;
        }
        default: {  // atan2(x, y) = atan(x / y)
          const double den = x.v * x.v + y.v * y.v;
          combine(x, y, atan2(x.v, y.v), y.v / den, -x.v / den);
          break;
        }
      }
    } else if (op == CVP_EXPR_SELECT) {
      const ExprDual<ND> no = stack[sp - 1];
      const ExprDual<ND> yes = stack[sp - 2];
      sp -= 2;
      const bool pick = stack[sp - 1].v != 0.0;
      stack[sp - 1] = pick ? yes : no;
    } else {
      ExprDual<ND>& x = stack[sp - 1];
      const double v = x.v;
      switch (op) {
        case CVP_EXPR_NEG: chain(x, -v, -1.0); break;
        case CVP_EXPR_SQRT: { const double r = sqrt(v); chain(x, r, 0.5 / r); break; }
        case CVP_EXPR_EXP: { const double r = exp(v); chain(x, r, r); break; }
        case CVP_EXPR_LOG: chain(x, log(v), 1.0 / v); break;
        case CVP_EXPR_SIN: chain(x, sin(v), cos(v)); break;
        case CVP_EXPR_COS: chain(x, cos(v), -sin(v)); break;
        case CVP_EXPR_TAN: { const double c = cos(v); chain(x, tan(v), 1.0 / (c * c)); break; }
        case CVP_EXPR_SEC: { const double c = cos(v); chain(x, 1.0 / c, tan(v) / c); break; }
        case CVP_EXPR_CSC: { const double s = sin(v); chain(x, 1.0 / s, -1.0 / (s * tan(v))); break; }
        case CVP_EXPR_COT: { const double s = sin(v); chain(x, 1.0 / tan(v), -1.0 / (s * s)); break; }
        case CVP_EXPR_ASIN: chain(x, asin(v), 1.0 / sqrt(1.0 - v * v)); break;
        case CVP_EXPR_ACOS: chain(x, acos(v), -1.0 / sqrt(1.0 - v * v)); break;
        case CVP_EXPR_ATAN: chain(x, atan(v), 1.0 / (1.0 + v * v)); break;
        case CVP_EXPR_SINH: chain(x, sinh(v), cosh(v)); break;
        case CVP_EXPR_COSH: chain(x, cosh(v), sinh(v)); break;
        case CVP_EXPR_TANH: { const double h = tanh(v); chain(x, h, 1.0 - h * h); break; }
        case CVP_EXPR_ERF: chain(x, erf(v), 1.1283791670955126 * exp(-v * v)); break;
        case CVP_EXPR_ERFC: chain(x, erfc(v), -1.1283791670955126 * exp(-v * v)); break;
        case CVP_EXPR_STEP: chain(x, v >= 0.0 ? 1.0 : 0.0, 0.0); x.dep = 0; break;
        case CVP_EXPR_DELTA: chain(x, v == 0.0 ? CUDART_INF : 0.0, 0.0); x.dep = 0; break;
        case CVP_EXPR_SQUARE: chain(x, v * v, 2.0 * v); break;
        case CVP_EXPR_CUBE: chain(x, v * v * v, 3.0 * v * v); break;
        case CVP_EXPR_RECIP: chain(x, 1.0 / v, -1.0 / (v * v)); break;
        case CVP_EXPR_ABS: chain(x, fabs(v), v >= 0.0 ? 1.0 : -1.0); break;
        case CVP_EXPR_FLOOR: chain(x, floor(v), 0.0); x.dep = 0; break;
        default: chain(x, ceil(v), 0.0); x.dep = 0; break;  // CVP_EXPR_CEIL
      }
    }
  }
  const ExprDual<ND>& result = stack[0];
  out_value[b] = T(result.v);
  if (out_partial != nullptr) {
    for (int k = 0; k < ninputs; ++k) out_partial[(int64_t)k * num_frames + b] = T(result.d[k]);
  }
}

// The program is checked on the host exactly as the kernel will walk it: stack depth, operand
// counts, slot and input indices -- a malformed program must fail here, not on the device.
int validate(const cvp_expr_program& p) {
  CVP_REQUIRE(p.num_ops >= 1 && p.num_ops <= CVP_EXPR_MAX_OPS, "program has no or too many operations");
  CVP_REQUIRE(p.num_variables >= 0 && p.num_parameters >= 0 &&
                  p.num_variables + p.num_parameters <= CVP_EXPR_MAX_INPUTS,
              "too many variables + parameters");
  int depth = 0;
  bool stored[CVP_EXPR_MAX_TEMPS] = {false};
  for (int pc = 0; pc < p.num_ops; ++pc) {
    const int op = p.ops[pc].code, arg = p.ops[pc].arg;
    int pops, pushes = 1;
    if (op == CVP_EXPR_CONST) pops = 0;
    else if (op == CVP_EXPR_INPUT) {
      pops = 0;
      CVP_REQUIRE(arg >= 0 && arg < p.num_variables + p.num_parameters, "input index out of range");
    } else if (op == CVP_EXPR_LOAD) {
      pops = 0;
      CVP_REQUIRE(arg >= 0 && arg < CVP_EXPR_MAX_TEMPS && stored[arg], "load of an unset slot");
    } else if (op == CVP_EXPR_STORE) {
      pops = 1;
      pushes = 0;
      CVP_REQUIRE(arg >= 0 && arg < CVP_EXPR_MAX_TEMPS, "slot out of range");
      stored[arg] = true;
    } else if (op >= CVP_EXPR_ADD && op <= CVP_EXPR_ATAN2) pops = 2;
    else if (op == CVP_EXPR_SELECT) pops = 3;
    else if (op >= CVP_EXPR_NEG && op <= CVP_EXPR_CEIL) pops = 1;
    else {
      set_error("unknown operation code " + std::to_string(op));
      return CVP_ERR_INVALID;
    }
    CVP_REQUIRE(depth >= pops, "stack underflow");
    depth += pushes - pops;
    CVP_REQUIRE(depth <= EXPR_STACK, "expression too deep");
  }
  CVP_REQUIRE(depth == 1, "program must leave exactly one value");
  return CVP_OK;
}

template <typename T>
int launch(const cvp_expr_program& p, const void* values, int64_t B, void* out_value, void* out_partial,
           cudaStream_t st) {
  const unsigned blocks = (unsigned)div_up(B, 128);
  const int ninputs = p.num_variables + p.num_parameters;
  const T* in = static_cast<const T*>(values);
  T* ov = static_cast<T*>(out_value);
  T* op = static_cast<T*>(out_partial);
  if (ninputs <= 4)
    expr_eval_kernel<T, 4><<<blocks, 128, 0, st>>>(p, in, B, ov, op);
  else if (ninputs <= 8)
    expr_eval_kernel<T, 8><<<blocks, 128, 0, st>>>(p, in, B, ov, op);
  else
    expr_eval_kernel<T, CVP_EXPR_MAX_INPUTS><<<blocks, 128, 0, st>>>(p, in, B, ov, op);
  CVP_LAUNCH_CHECK();
  return CVP_OK;
}

}  // namespace
}  // namespace cvp

using namespace cvp;

extern "C" int cvp_expr_eval(const cvp_expr_program* program, int dtype, const void* values,
                             int64_t num_frames, void* out_value, void* out_partial,
                             void* stream) {
  CVP_REQUIRE(program != nullptr && out_value != nullptr, "null argument");
  CVP_REQUIRE(dtype == CVP_F32 || dtype == CVP_F64, "dtype must be CVP_F32 or CVP_F64");
  CVP_REQUIRE(values != nullptr || program->num_variables == 0, "null values");
  int status = validate(*program);
  if (status != CVP_OK) return status;
  if (num_frames <= 0) return CVP_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return dtype == CVP_F32 ? launch<float>(*program, values, num_frames, out_value, out_partial, st)
                          : launch<double>(*program, values, num_frames, out_value, out_partial, st);
}
