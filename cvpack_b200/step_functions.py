"""
Whitelist of step-function expressions.

The reference accepts any Lepton expression in ``x`` for ``stepFunction``
(``/root/reference/cvpack/number_of_contacts.py:132``, ``residue_coordination.py:113``,
``helix_rmsd_content.py:135``) because OpenMM compiles it at run time.  Hand-written kernels cannot:
the forms below (which cover every expression used in the reference's code, docs and tests) map to
the ``CVP_STEP_*`` kernels; anything else raises ``NotImplementedError``.  The string itself is kept
verbatim for serialization.
"""

from __future__ import annotations

import re
import typing as t

from . import _native

_INT = r"(\d+)"
_PATTERNS: t.List[t.Tuple[t.Pattern, t.Callable[..., t.Optional[t.Tuple[int, float, float]]]]] = [
    (
        re.compile(rf"^1/\(1\+x\^{_INT}\)$"),
        lambda n: (_native.CVP_STEP_RATIONAL, float(n), 0.0),
    ),
    (re.compile(r"^step\(1-x\)$"), lambda: (_native.CVP_STEP_HEAVISIDE, 0.0, 0.0)),
    (
        re.compile(rf"^\(1\+x\^{_INT}\)/\(1\+x\^{_INT}\+x\^{_INT}\)$"),
        lambda n, n2, m: (
            (_native.CVP_STEP_RATIONAL_PAIR, float(n), 0.0)
            if int(n) == int(n2) and int(m) == 2 * int(n)
            else None
        ),
    ),
    (
        re.compile(rf"^\(1-x\^{_INT}\)/\(1-x\^{_INT}\)$"),
        lambda n, m: (_native.CVP_STEP_PLUMED, float(n), float(m)),
    ),
]

_NAMES = {
    _native.CVP_STEP_RATIONAL: "rational",
    _native.CVP_STEP_HEAVISIDE: "heaviside",
    _native.CVP_STEP_RATIONAL_PAIR: "rational_pair",
    _native.CVP_STEP_PLUMED: "plumed",
}


def parse_step_function(expression: str) -> t.Tuple[int, float, float]:
    """
    Map a step-function expression to ``(CVP_STEP_* kind, p0, p1)``.

    >>> parse_step_function("1/(1+x^6)")
    (0, 6.0, 0.0)
    >>> parse_step_function("step(1-x)")
    (1, 0.0, 0.0)
    >>> parse_step_function("(1+x^4)/(1+x^4+x^8)")
    (2, 4.0, 0.0)
    >>> parse_step_function("(1 - x**6)/(1 - x**12)")
    (3, 6.0, 12.0)
    """
    text = re.sub(r"\s+", "", expression).replace("**", "^")
    for pattern, build in _PATTERNS:
        match = pattern.match(text)
        if match:
            result = build(*match.groups())
            if result is not None:
                if result[0] != _native.CVP_STEP_HEAVISIDE and result[1] < 1:
                    break
                return result
    raise NotImplementedError(
        f"step function {expression!r} is not one of the forms the CUDA kernels implement: "
        "'1/(1+x^n)', 'step(1-x)', '(1+x^n)/(1+x^n+x^(2n))', '(1-x^n)/(1-x^m)'"
    )


def step_function_name(kind: int) -> str:
    """Name of a ``CVP_STEP_*`` kind as used by the CPU oracle."""
    return _NAMES[kind]
