"""
The expression compiler behind MetaCollectiveVariable (cvpack_b200/expression.py) and the kernel that
interprets its programs (csrc/expr.cu), against the oracle's independent Lepton evaluator
(oracle/refshim/lepton.py: recursive descent + forward-mode duals; the product is shunting-yard to
postfix + a stack machine).
"""

import math

import numpy as np
import pytest

from cvpack_b200 import _native, expression
from oracle.refshim import lepton

EXPRESSIONS = [
    ("a+b*c", "abc"),
    ("-a^2*b", "ab"),
    ("2^-a", "a"),
    ("a-b-c", "abc"),
    ("a/b/c", "abc"),
    ("a^b^c", "abc"),
    ("(a+b)*(a-b)/c", "abc"),
    ("-(-a)", "a"),
    ("+a - -b", "ab"),
    ("1.5e-1*a + .5*b + 2.*c", "abc"),
    ("sqrt(a)+exp(-b)+log(c)", "abc"),
    ("sin(a)*cos(b)+tan(c)/3", "abc"),
    ("sec(a)+csc(b)+cot(c)", "abc"),
    ("asin(a/4)+acos(b/4)+atan(c)", "abc"),
    ("sinh(a)+cosh(b)+tanh(c)", "abc"),
    ("erf(a)+erfc(b)", "ab"),
    ("step(a-b)*c + step(b-a)*a", "abc"),
    ("square(a)+cube(b)+recip(c)", "abc"),
    ("abs(a-b)+floor(c*3)+ceil(a*3)", "abc"),
    ("min(a,b)+max(b,c)+min(a,a)", "abc"),
    ("atan2(a-1,b-1)", "ab"),
    ("select(step(a-1.2), b^2, c^3)", "abc"),
    ("0.5*k*min(d, 6.283185307179586-d)^2; d=abs(a-b)", "ab", {"k": 3.5}),
    ("w/s; w=exp(-x/t); s=1+w; x=(a-b)^2+c; t=0.7", "abc"),
    ("a^p + p^a", "a", {"p": 1.7}),
    ("(a^2+b^2)^0.5", "ab"),
    ("u*v; u=a+b; v=u*c", "abc"),
]


def run_program(program, inputs):
    """Reference interpreter of the postfix programs in plain Python (values only)."""
    stack, temps = [], {}
    binary = {
        _native.CVP_EXPR_ADD: lambda x, y: x + y, _native.CVP_EXPR_SUB: lambda x, y: x - y,
        _native.CVP_EXPR_MUL: lambda x, y: x * y, _native.CVP_EXPR_DIV: lambda x, y: x / y,
        _native.CVP_EXPR_POW: lambda x, y: x**y, _native.CVP_EXPR_MIN: min, _native.CVP_EXPR_MAX: max,
        _native.CVP_EXPR_ATAN2: math.atan2,
    }
    for code, arg, value in program.ops:
        if code == _native.CVP_EXPR_CONST:
            stack.append(value)
        elif code == _native.CVP_EXPR_INPUT:
            stack.append(inputs[arg])
        elif code == _native.CVP_EXPR_LOAD:
            stack.append(temps[arg])
        elif code == _native.CVP_EXPR_STORE:
            temps[arg] = stack.pop()
        elif code in binary:
            y, x = stack.pop(), stack.pop()
            stack.append(binary[code](x, y))
        elif code == _native.CVP_EXPR_SELECT:
            no, yes, cond = stack.pop(), stack.pop(), stack.pop()
            stack.append(yes if cond != 0 else no)
        elif code == _native.CVP_EXPR_NEG:
            stack.append(-stack.pop())
        else:
            name = next(k for k, v in expression._UNARY.items() if v == code)  # pylint: disable=protected-access
            result = lepton.call(name, [lepton.Dual(stack.pop())]).v
            stack.append(float(result))
    (result,) = stack
    return result


def sample_inputs(rng, names):
    return {name: float(rng.uniform(0.3, 2.2)) for name in names}


@pytest.mark.parametrize("case", EXPRESSIONS, ids=[c[0][:40] for c in EXPRESSIONS])
def test_compiled_program_agrees_with_the_oracle_parser(case):
    text, names = case[0], list(case[1])
    params = case[2] if len(case) > 2 else {}
    program = expression.compile_expression(text, names, list(params))
    rng = np.random.default_rng(5)
    for _ in range(6):
        values = sample_inputs(rng, names)
        expected = float(lepton.Expression(text).evaluate({**values, **params}).v)
        got = run_program(program, [values[n] for n in names] + list(params.values()))
        assert got == pytest.approx(expected, rel=1e-13, abs=1e-13), text


@pytest.mark.parametrize("text,message", [
    ("a+", "incomplete"), ("a b", "missing operator"), ("foo(a)", "unknown function"),
    ("a+q", "unknown variable"), ("min(a)", "takes 2"), ("(a+b", "unbalanced"), ("a+b)", "unbalanced"),
    ("a; a=2", "defined more than once"), ("x; x=y; y=x", "cyclic"), ("a $ b", "cannot parse"),
    ("a; =b", "malformed"), ("", "empty"),
])
def test_malformed_expressions_are_rejected(text, message):
    with pytest.raises(expression.ExpressionError, match=message):
        expression.compile_expression(text, ["a", "b"])


def test_limits():
    with pytest.raises(expression.ExpressionError, match="at most"):
        expression.compile_expression("a", [f"v{k}" for k in range(17)])
    long_sum = "+".join(["a"] * 120)
    with pytest.raises(expression.ExpressionError, match="operations"):
        expression.compile_expression(long_sum, ["a"]).to_struct({})


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("case", EXPRESSIONS, ids=[c[0][:40] for c in EXPRESSIONS])
def test_kernel_value_and_partials(case, precision):
    import torch

    text, names = case[0], list(case[1])
    params = case[2] if len(case) > 2 else {}
    program = expression.compile_expression(text, names, list(params))
    rng = np.random.default_rng(6)
    frames = 37
    dtype = torch.float32 if precision == "single" else torch.float64
    host = rng.uniform(0.3, 2.2, size=(len(names), frames))
    values = torch.as_tensor(host, dtype=dtype, device="cuda").contiguous()
    seen = values.double().cpu().numpy()
    out = torch.empty(frames, dtype=dtype, device="cuda")
    partials = torch.empty((len(names) + len(params), frames), dtype=dtype, device="cuda")
    _native.expr_eval(
        program.to_struct(params), _native.CVP_F32 if precision == "single" else _native.CVP_F64,
        values.data_ptr(), frames, out.data_ptr(), partials.data_ptr(), 0,
    )
    torch.cuda.synchronize()
    tol = 2e-6 if precision == "single" else 1e-12
    for b in range(frames):
        variables = {n: lepton.Dual.variable(n, seen[k, b]) for k, n in enumerate(names)}
        variables.update({n: lepton.Dual.variable(n, v) for n, v in params.items()})
        ref = lepton.Expression(text).evaluate(variables)
        assert float(out[b]) == pytest.approx(float(ref.v), rel=tol, abs=tol)
        for k, n in enumerate(names + list(params)):
            assert float(partials[k, b]) == pytest.approx(float(ref.d.get(n, 0.0)), rel=tol, abs=tol), (text, n)


@pytest.mark.gpu
def test_malformed_program_is_rejected_on_the_host():
    import torch

    program = _native.ExprProgram()
    program.num_ops = 1
    program.num_variables = 1
    program.ops[0].code = _native.CVP_EXPR_ADD  # pops two values from an empty stack
    values = torch.zeros((1, 4), device="cuda")
    out = torch.empty(4, device="cuda")
    lib = _native.load()
    import ctypes

    status = lib.cvp_expr_eval(ctypes.byref(program), _native.CVP_F32, values.data_ptr(), 4, out.data_ptr(), None, None)
    assert status == -1 and b"underflow" in lib.cvp_last_error()
