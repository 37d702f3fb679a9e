"""
A small evaluator for the Lepton expression language used by OpenMM's ``Custom*Force`` classes.

TEST INFRASTRUCTURE ONLY (part of ``oracle/``): it exists so that the *reference's own constructors*
(``/root/reference/cvpack/*.py``) can run without OpenMM and the expression strings they build
(e.g. ``number_of_contacts.py:143``, ``attraction_strength.py:189-199``, ``base_path_cv.py:43-59``,
``base_rmsd_content.py:48-59``) can be evaluated, together with first derivatives, in FP64.

OpenMM/Lepton is an absent third-party dependency (unpinned); what is restated
here from its published behaviour:

* grammar: numbers, names, ``+ - * / ^``, unary minus (binds tighter than ``* /``, looser than
  ``^``), calls ``f(a, b, ...)``, and ``expr; name = expr; ...`` definitions that may be used by
  anything to their left;
* functions: sqrt exp log sin cos sec csc tan cot asin acos atan atan2 sinh cosh tanh erf erfc
  step delta square cube recip min max abs floor ceil select;
* derivative rules that matter at kinks: ``step' = 0`` (``delta`` away from 0),
  ``abs'(x) = 2*step(x) - 1`` (so ``abs'(0) = +1``), ``min(a, b)' = b'`` when ``a >= b`` else ``a'``
  (ties pick the second argument), ``max(a, b)' = a'`` when ``a >= b`` else ``b'``,
  ``select(c, a, b)`` differentiates the chosen branch, ``floor' = ceil' = 0``.

Evaluation is forward-mode automatic differentiation over NumPy arrays: every value is a
:class:`Dual` ``(v, {variable: dv/dvariable})``.
"""

from __future__ import annotations

import math
import re
import typing as t

import numpy as np
from scipy import special

_TOKEN = re.compile(
    r"\s*(?:(?P<num>(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?)|(?P<name>[A-Za-z_][A-Za-z_0-9]*)|(?P<op>[-+*/^(),;=]))"
)


def tokenize(text: str) -> t.List[t.Tuple[str, str]]:
    """Split an expression into (kind, text) tokens."""
    tokens, pos = [], 0
    text = text.rstrip()
    while pos < len(text):
        match = _TOKEN.match(text, pos)
        if match is None:
            raise ValueError(f"Lepton: cannot tokenize {text[pos:pos + 20]!r}")
        pos = match.end()
        for kind in ("num", "name", "op"):
            if match.group(kind) is not None:
                tokens.append((kind, match.group(kind)))
    return tokens


class Node:
    """Expression tree node: ('num', value) | ('var', name) | ('call', name, args) | (op, a, b)."""

    __slots__ = ("kind", "value", "args")

    def __init__(self, kind: str, value: t.Any = None, args: t.Sequence["Node"] = ()) -> None:
        self.kind, self.value, self.args = kind, value, tuple(args)

    def __repr__(self) -> str:
        return f"Node({self.kind}, {self.value}, {self.args})"


_PRECEDENCE = {"+": 0, "-": 0, "*": 1, "/": 1, "^": 3}


class _Parser:
    def __init__(self, tokens: t.List[t.Tuple[str, str]]) -> None:
        self.tokens, self.pos = tokens, 0

    def peek(self) -> t.Optional[t.Tuple[str, str]]:
        return self.tokens[self.pos] if self.pos < len(self.tokens) else None

    def take(self) -> t.Tuple[str, str]:
        token = self.tokens[self.pos]
        self.pos += 1
        return token

    def parse(self, precedence: int = 0) -> Node:
        token = self.take()
        if token[0] == "num":
            left = Node("num", float(token[1]))
        elif token[0] == "name":
            nxt = self.peek()
            if nxt == ("op", "("):
                self.take()
                args = []
                if self.peek() != ("op", ")"):
                    while True:
                        args.append(self.parse(0))
                        sep = self.take()
                        if sep == ("op", ")"):
                            break
                        if sep != ("op", ","):
                            raise ValueError("Lepton: expected , or ) in call")
                else:
                    self.take()
                left = Node("call", token[1], args)
            else:
                left = Node("var", token[1])
        elif token == ("op", "("):
            left = self.parse(0)
            if self.take() != ("op", ")"):
                raise ValueError("Lepton: expected )")
        elif token == ("op", "-"):
            left = Node("neg", None, [self.parse(2)])
        elif token == ("op", "+"):
            left = self.parse(2)
        else:
            raise ValueError(f"Lepton: unexpected token {token}")
        while True:
            nxt = self.peek()
            if nxt is None or nxt[0] != "op" or nxt[1] not in _PRECEDENCE:
                break
            op = nxt[1]
            prec = _PRECEDENCE[op]
            if prec < precedence:
                break
            self.take()
            # ^ is right-associative, the others left-associative
            right = self.parse(prec if op == "^" else prec + 1)
            left = Node(op, None, [left, right])
        return left


class Expression:
    """A parsed ``main; name1 = def1; ...`` expression."""

    def __init__(self, text: str) -> None:
        self.text = text
        parts = [part for part in text.split(";")]
        if not parts[0].strip():
            raise ValueError("Lepton: empty expression")
        self.main = self._parse_one(parts[0])
        self.definitions: t.Dict[str, Node] = {}
        for part in parts[1:]:
            if not part.strip():
                continue
            name, _, body = part.partition("=")
            name = name.strip()
            if not re.fullmatch(r"[A-Za-z_][A-Za-z_0-9]*", name) or not body.strip():
                raise ValueError(f"Lepton: bad definition {part!r}")
            self.definitions[name] = self._parse_one(body)

    @staticmethod
    def _parse_one(text: str) -> Node:
        parser = _Parser(tokenize(text))
        node = parser.parse(0)
        if parser.pos != len(parser.tokens):
            raise ValueError(f"Lepton: trailing tokens in {text!r}")
        return node

    def free_variables(self) -> t.Set[str]:
        """Names that are not defined inside the expression (call names excluded)."""
        found: t.Set[str] = set()

        def walk(node: Node) -> None:
            if node.kind == "var":
                if node.value in self.definitions:
                    return
                found.add(node.value)
            for arg in node.args:
                walk(arg)

        walk(self.main)
        for body in self.definitions.values():
            walk(body)
        return found

    def calls(self, names: t.Collection[str]) -> t.List[Node]:
        """All call nodes whose function name is in ``names`` (in first-seen order, unique)."""
        seen: t.Dict[t.Tuple, Node] = {}

        def key(node: Node) -> t.Tuple:
            return (node.kind, node.value, tuple(key(a) for a in node.args))

        def walk(node: Node) -> None:
            if node.kind == "call" and node.value in names:
                seen.setdefault(key(node), node)
            for arg in node.args:
                walk(arg)

        walk(self.main)
        for body in self.definitions.values():
            walk(body)
        return list(seen.values())

    def evaluate(
        self,
        variables: t.Mapping[str, t.Any],
        special_calls: t.Optional[t.Callable[[Node], t.Optional["Dual"]]] = None,
    ) -> "Dual":
        """Value and derivatives; ``variables`` maps names to numbers / arrays / :class:`Dual`."""
        cache: t.Dict[str, Dual] = {}

        def lookup(name: str) -> Dual:
            if name in cache:
                return cache[name]
            if name in self.definitions:
                cache[name] = ev(self.definitions[name])
            elif name in variables:
                cache[name] = as_dual(variables[name])
            else:
                raise KeyError(f"Lepton: unknown variable {name!r} in {self.text[:80]!r}")
            return cache[name]

        def ev(node: Node) -> Dual:
            kind = node.kind
            if kind == "num":
                return Dual(node.value)
            if kind == "var":
                return lookup(node.value)
            if kind == "neg":
                return -ev(node.args[0])
            if kind == "call":
                if special_calls is not None:
                    result = special_calls(node)
                    if result is not None:
                        return result
                return call(node.value, [ev(arg) for arg in node.args])
            a, b = ev(node.args[0]), ev(node.args[1])
            if kind == "+":
                return a + b
            if kind == "-":
                return a - b
            if kind == "*":
                return a * b
            if kind == "/":
                return a / b
            if kind == "^":
                return power(a, b)
            raise ValueError(kind)

        return ev(self.main)


class Dual:
    """Value plus partial derivatives with respect to named variables."""

    __slots__ = ("v", "d")

    def __init__(self, v: t.Any, d: t.Optional[t.Dict[str, t.Any]] = None) -> None:
        self.v = v
        self.d = d or {}

    @staticmethod
    def variable(name: str, value: t.Any) -> "Dual":
        """An independent variable: d(name)/d(name) = 1."""
        return Dual(value, {name: np.ones_like(np.asarray(value, dtype=np.float64))})

    def _combine(self, other: "Dual", fa: t.Any, fb: t.Any) -> t.Dict[str, t.Any]:
        out: t.Dict[str, t.Any] = {}
        for key, val in self.d.items():
            out[key] = fa * val
        for key, val in other.d.items():
            out[key] = out[key] + fb * val if key in out else fb * val
        return out

    def __neg__(self) -> "Dual":
        return Dual(-self.v, {k: -v for k, v in self.d.items()})

    def __add__(self, other: "Dual") -> "Dual":
        return Dual(self.v + other.v, self._combine(other, 1.0, 1.0))

    def __sub__(self, other: "Dual") -> "Dual":
        return Dual(self.v - other.v, self._combine(other, 1.0, -1.0))

    def __mul__(self, other: "Dual") -> "Dual":
        return Dual(self.v * other.v, self._combine(other, other.v, self.v))

    def __truediv__(self, other: "Dual") -> "Dual":
        inv = 1.0 / other.v
        return Dual(self.v * inv, self._combine(other, inv, -self.v * inv * inv))

    def chain(self, value: t.Any, slope: t.Any) -> "Dual":
        """f(self) given f's value and f'(self.v)."""
        return Dual(value, {k: slope * v for k, v in self.d.items()})


def as_dual(value: t.Any) -> Dual:
    """Wrap a plain number / array."""
    return value if isinstance(value, Dual) else Dual(value)


def _step(x: t.Any) -> t.Any:
    return np.where(np.asarray(x) >= 0, 1.0, 0.0)


def power(a: Dual, b: Dual) -> Dual:
    """``a ^ b``; a constant exponent uses ``b a^(b-1)`` (valid for negative bases)."""
    with np.errstate(all="ignore"):
        value = np.power(np.asarray(a.v, dtype=np.float64), b.v)
        if not b.d:
            slope = b.v * np.power(np.asarray(a.v, dtype=np.float64), b.v - 1.0)
            slope = np.where(np.asarray(b.v) == 0, 0.0, slope)
            return a.chain(value, slope)
        da = b.v * np.power(np.asarray(a.v, dtype=np.float64), b.v - 1.0)
        db = value * np.log(a.v)
    return Dual(value, a._combine(b, da, db))  # pylint: disable=protected-access


def call(name: str, args: t.List[Dual]) -> Dual:
    """Built-in Lepton functions with Lepton's derivative rules."""
    x = args[0]
    v = np.asarray(x.v, dtype=np.float64)
    with np.errstate(all="ignore"):
        if name == "sqrt":
            r = np.sqrt(v)
            return x.chain(r, 0.5 / r)
        if name == "exp":
            r = np.exp(v)
            return x.chain(r, r)
        if name == "log":
            return x.chain(np.log(v), 1.0 / v)
        if name == "sin":
            return x.chain(np.sin(v), np.cos(v))
        if name == "cos":
            return x.chain(np.cos(v), -np.sin(v))
        if name == "tan":
            return x.chain(np.tan(v), 1.0 / np.cos(v) ** 2)
        if name == "sec":
            return x.chain(1.0 / np.cos(v), np.tan(v) / np.cos(v))
        if name == "csc":
            return x.chain(1.0 / np.sin(v), -1.0 / (np.sin(v) * np.tan(v)))
        if name == "cot":
            return x.chain(1.0 / np.tan(v), -1.0 / np.sin(v) ** 2)
        if name == "asin":
            return x.chain(np.arcsin(v), 1.0 / np.sqrt(1.0 - v * v))
        if name == "acos":
            return x.chain(np.arccos(v), -1.0 / np.sqrt(1.0 - v * v))
        if name == "atan":
            return x.chain(np.arctan(v), 1.0 / (1.0 + v * v))
        if name == "sinh":
            return x.chain(np.sinh(v), np.cosh(v))
        if name == "cosh":
            return x.chain(np.cosh(v), np.sinh(v))
        if name == "tanh":
            return x.chain(np.tanh(v), 1.0 - np.tanh(v) ** 2)
        if name == "erf":
            return x.chain(special.erf(v), 2.0 / math.sqrt(math.pi) * np.exp(-v * v))
        if name == "erfc":
            return x.chain(special.erfc(v), -2.0 / math.sqrt(math.pi) * np.exp(-v * v))
        if name == "step":
            return Dual(_step(v))
        if name == "delta":
            return Dual(np.where(v == 0, np.inf, 0.0))
        if name == "square":
            return x.chain(v * v, 2.0 * v)
        if name == "cube":
            return x.chain(v * v * v, 3.0 * v * v)
        if name == "recip":
            return x.chain(1.0 / v, -1.0 / (v * v))
        if name == "abs":
            return x.chain(np.abs(v), 2.0 * _step(v) - 1.0)
        if name == "floor":
            return Dual(np.floor(v))
        if name == "ceil":
            return Dual(np.ceil(v))
        if name in ("min", "max"):
            y = args[1]
            step = _step(v - np.asarray(y.v, dtype=np.float64))
            if name == "min":
                # ties (step == 1) take the derivative of the second argument
                return Dual(np.minimum(v, y.v), x._combine(y, 1.0 - step, step))  # pylint: disable=protected-access
            return Dual(np.maximum(v, y.v), x._combine(y, step, 1.0 - step))  # pylint: disable=protected-access
        if name == "atan2":
            y = args[1]  # atan2(x=first, y=second) = atan(first/second)
            den = v * v + np.asarray(y.v) ** 2
            return Dual(np.arctan2(v, y.v), x._combine(y, y.v / den, -v / den))  # pylint: disable=protected-access
        if name == "select":
            a, b = args[1], args[2]
            pick = np.where(v != 0, 1.0, 0.0)
            return Dual(np.where(v != 0, a.v, b.v), a._combine(b, pick, 1.0 - pick))  # pylint: disable=protected-access
    raise ValueError(f"Lepton: unknown function {name!r}")
