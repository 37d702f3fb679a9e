"""
Compiler from Lepton-style expression strings to the postfix programs of ``cvp_expr_eval``.

The reference passes the ``function`` string of a :class:`MetaCollectiveVariable` to an
``openmm.CustomCVForce`` (``/root/reference/cvpack/meta_collective_variable.py:127-133``), where
OpenMM's Lepton parses and differentiates it.  The CUDA library interprets a short postfix program
per frame instead (``csrc/expr.cu``); this module produces that program with Lepton's grammar:

* numbers, names, ``+ - * / ^`` (``^`` right-associative and tighter than unary minus, unary minus
  tighter than ``* /``), parentheses, function calls;
* ``main; name = definition; ...``: a definition may be used by anything to its left and is
  evaluated once (a temporary slot of the program);
* functions: ``sqrt exp log sin cos sec csc tan cot asin acos atan atan2 sinh cosh tanh erf erfc
  step delta square cube recip min max abs floor ceil select``.

The parser is a shunting-yard pass straight to postfix (no tree is built).
"""

from __future__ import annotations

import re
import typing as t

from . import _native

_UNARY = {
    "sqrt": _native.CVP_EXPR_SQRT, "exp": _native.CVP_EXPR_EXP, "log": _native.CVP_EXPR_LOG,
    "sin": _native.CVP_EXPR_SIN, "cos": _native.CVP_EXPR_COS, "tan": _native.CVP_EXPR_TAN,
    "sec": _native.CVP_EXPR_SEC, "csc": _native.CVP_EXPR_CSC, "cot": _native.CVP_EXPR_COT,
    "asin": _native.CVP_EXPR_ASIN, "acos": _native.CVP_EXPR_ACOS, "atan": _native.CVP_EXPR_ATAN,
    "sinh": _native.CVP_EXPR_SINH, "cosh": _native.CVP_EXPR_COSH, "tanh": _native.CVP_EXPR_TANH,
    "erf": _native.CVP_EXPR_ERF, "erfc": _native.CVP_EXPR_ERFC, "step": _native.CVP_EXPR_STEP,
    "delta": _native.CVP_EXPR_DELTA, "square": _native.CVP_EXPR_SQUARE, "cube": _native.CVP_EXPR_CUBE,
    "recip": _native.CVP_EXPR_RECIP, "abs": _native.CVP_EXPR_ABS, "floor": _native.CVP_EXPR_FLOOR,
    "ceil": _native.CVP_EXPR_CEIL,
}
_BINARY_FUNCTIONS = {"min": _native.CVP_EXPR_MIN, "max": _native.CVP_EXPR_MAX, "atan2": _native.CVP_EXPR_ATAN2}
_OPERATORS = {
    "+": (0, False, _native.CVP_EXPR_ADD), "-": (0, False, _native.CVP_EXPR_SUB),
    "*": (1, False, _native.CVP_EXPR_MUL), "/": (1, False, _native.CVP_EXPR_DIV),
    "neg": (2, True, _native.CVP_EXPR_NEG), "^": (3, True, _native.CVP_EXPR_POW),
}
_TOKEN = re.compile(
    r"\s*(?:(?P<number>(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?)|(?P<name>[A-Za-z_]\w*)|(?P<symbol>[-+*/^(),]))"
)
_NAME = re.compile(r"[A-Za-z_]\w*$")

Op = t.Tuple[int, int, float]


class ExpressionError(ValueError):
    """The expression cannot be parsed or uses something the kernel does not implement."""


def _tokens(text: str) -> t.List[t.Tuple[str, str]]:
    out, pos = [], 0
    text = text.strip()
    while pos < len(text):
        match = _TOKEN.match(text, pos)
        if match is None:
            raise ExpressionError(f"cannot parse {text[pos:pos + 16]!r} in expression {text!r}")
        pos = match.end()
        kind = match.lastgroup
        out.append((kind, match.group(kind)))
    return out


def _to_postfix(text: str, leaf: t.Callable[[str], Op]) -> t.List[Op]:
    """Shunting-yard: one expression (no ``;``) to a list of (code, arg, value)."""
    output: t.List[Op] = []
    stack: t.List[t.Tuple[str, t.Any]] = []  # ("op", symbol) | ("(", None) | ("call", [name, argc])
    previous = "start"  # start | value | operator | open | comma

    def pop_operator() -> None:
        _, symbol = stack.pop()
        output.append((_OPERATORS[symbol][2], 0, 0.0))

    tokens = _tokens(text)
    if not tokens:
        raise ExpressionError("empty expression")
    for index, (kind, value) in enumerate(tokens):
        if kind == "number":
            if previous == "value":
                raise ExpressionError(f"missing operator before {value!r} in {text!r}")
            output.append((_native.CVP_EXPR_CONST, 0, float(value)))
            previous = "value"
        elif kind == "name":
            if previous == "value":
                raise ExpressionError(f"missing operator before {value!r} in {text!r}")
            is_call = index + 1 < len(tokens) and tokens[index + 1] == ("symbol", "(")
            if is_call:
                if value not in _UNARY and value not in _BINARY_FUNCTIONS and value != "select":
                    raise ExpressionError(f"unknown function {value!r} in {text!r}")
                stack.append(("call", [value, 1]))
                previous = "operator"
            else:
                output.append(leaf(value))
                previous = "value"
        elif value == "(":
            if previous == "value":
                raise ExpressionError(f"missing operator before '(' in {text!r}")
            stack.append(("(", None))
            previous = "open"
        elif value == ",":
            while stack and stack[-1][0] == "op":
                pop_operator()
            if len(stack) < 2 or stack[-1][0] != "(" or stack[-2][0] != "call" or previous != "value":
                raise ExpressionError(f"misplaced ',' in {text!r}")
            stack[-2][1][1] += 1
            previous = "comma"
        elif value == ")":
            if previous != "value":
                raise ExpressionError(f"misplaced ')' in {text!r}")
            while stack and stack[-1][0] == "op":
                pop_operator()
            if not stack or stack[-1][0] != "(":
                raise ExpressionError(f"unbalanced ')' in {text!r}")
            stack.pop()
            if stack and stack[-1][0] == "call":
                name, argc = stack.pop()[1]
                expected = 1 if name in _UNARY else 3 if name == "select" else 2
                if argc != expected:
                    raise ExpressionError(f"{name} takes {expected} argument(s), got {argc}, in {text!r}")
                code = _UNARY.get(name) or _BINARY_FUNCTIONS.get(name) or _native.CVP_EXPR_SELECT
                output.append((code, 0, 0.0))
            previous = "value"
        else:  # + - * / ^
            if previous != "value":
                if value == "+":
                    continue
                if value != "-":
                    raise ExpressionError(f"misplaced {value!r} in {text!r}")
                symbol = "neg"
            else:
                symbol = value
            precedence, right = _OPERATORS[symbol][:2]
            if symbol != "neg":  # a prefix operator never closes what is on its left
                while stack and stack[-1][0] == "op":
                    top = _OPERATORS[stack[-1][1]][0]
                    if top > precedence or (top == precedence and not right):
                        pop_operator()
                    else:
                        break
            stack.append(("op", symbol))
            previous = "operator"
    if previous != "value":
        raise ExpressionError(f"incomplete expression {text!r}")
    while stack:
        if stack[-1][0] != "op":
            raise ExpressionError(f"unbalanced '(' in {text!r}")
        pop_operator()
    return output


class Program:
    """A compiled expression: ``ops`` for ``cvp_expr_eval`` plus the input layout."""

    def __init__(self, ops: t.List[Op], variables: t.Sequence[str], parameters: t.Sequence[str]) -> None:
        self.ops = ops
        self.variables = list(variables)
        self.parameters = list(parameters)

    def to_struct(self, parameter_values: t.Mapping[str, float]) -> "_native.ExprProgram":
        """The ``cvp_expr_program`` with the given parameter values filled in."""
        if len(self.ops) > _native.CVP_EXPR_MAX_OPS:
            raise ExpressionError(
                f"the expression compiles to {len(self.ops)} operations (limit {_native.CVP_EXPR_MAX_OPS})"
            )
        program = _native.ExprProgram()
        program.num_ops = len(self.ops)
        program.num_variables = len(self.variables)
        program.num_parameters = len(self.parameters)
        for k, name in enumerate(self.parameters):
            program.parameters[k] = float(parameter_values[name])
        for k, (code, arg, value) in enumerate(self.ops):
            program.ops[k].code, program.ops[k].arg, program.ops[k].value = code, arg, value
        return program


def compile_expression(text: str, variables: t.Sequence[str], parameters: t.Sequence[str] = ()) -> Program:
    """
    Compile ``"main; a = ...; b = ..."`` over the given variable and parameter names.

    Raises :class:`ExpressionError` for unknown names or functions, malformed syntax, cyclic
    definitions, or more inputs / temporaries than the kernel supports.
    """
    variables, parameters = list(variables), list(parameters)
    inputs = variables + parameters
    if len(set(inputs)) != len(inputs):
        raise ExpressionError("variable and parameter names must be distinct")
    if len(inputs) > _native.CVP_EXPR_MAX_INPUTS:
        raise ExpressionError(f"at most {_native.CVP_EXPR_MAX_INPUTS} variables + parameters are supported")
    parts = text.split(";")
    definitions: t.Dict[str, str] = {}
    for part in parts[1:]:
        if not part.strip():
            continue
        name, equals, body = part.partition("=")
        name = name.strip()
        if not equals or not _NAME.match(name) or not body.strip():
            raise ExpressionError(f"malformed definition {part.strip()!r}")
        if name in definitions or name in inputs:
            raise ExpressionError(f"{name!r} is defined more than once")
        definitions[name] = body
    slots: t.Dict[str, int] = {}
    in_progress: t.Set[str] = set()
    ops: t.List[Op] = []

    def leaf(name: str) -> Op:
        if name in definitions:
            if name not in slots:
                define(name)
            return (_native.CVP_EXPR_LOAD, slots[name], 0.0)
        if name in inputs:
            return (_native.CVP_EXPR_INPUT, inputs.index(name), 0.0)
        raise ExpressionError(f"unknown variable {name!r} in expression {text!r}")

    def define(name: str) -> None:
        if name in in_progress:
            raise ExpressionError(f"cyclic definition of {name!r}")
        in_progress.add(name)
        body = _to_postfix(definitions[name], leaf)  # (defines what it needs first)
        in_progress.discard(name)
        if len(slots) >= _native.CVP_EXPR_MAX_TEMPS:
            raise ExpressionError(f"more than {_native.CVP_EXPR_MAX_TEMPS} definitions are in use")
        slots[name] = len(slots)
        ops.extend(body)
        ops.append((_native.CVP_EXPR_STORE, slots[name], 0.0))

    main = _to_postfix(parts[0], leaf)
    ops.extend(main)
    return Program(ops, variables, parameters)


__all__ = ["ExpressionError", "Program", "compile_expression"]
