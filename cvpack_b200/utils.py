"""
Host-side helpers shared by the collective variables.

Counterparts of ``/root/reference/cvpack/utils.py``: ``preprocess_args`` (``:187-247``),
``convert_to_matrix`` (``:72-93``), ``evaluate_in_context`` (``:32-69``) and
``compute_effective_mass`` (``:151-184``) -- the last two evaluate through a
:class:`~cvpack_b200.batch.BatchContext` (CUDA) instead of an ``openmm.Context``.
"""

from __future__ import annotations

import functools
import inspect
import typing as t

import numpy as np

from . import app as mmapp
from . import mm, mmunit
from .serialization import (
    Serializable,
    SerializableAtom,
    SerializableForce,
    SerializableResidue,
)
from .units import Quantity, Unit


def _is_atom(data: t.Any) -> bool:
    return isinstance(data, mmapp.Atom) or (
        type(data).__name__ == "Atom" and hasattr(data, "residue") and hasattr(data, "element")
    )


def _is_residue(data: t.Any) -> bool:
    return isinstance(data, mmapp.Residue) or (
        type(data).__name__ == "Residue" and hasattr(data, "chain") and hasattr(data, "atoms")
    )


def _is_nonbonded_force(data: t.Any) -> bool:
    return not isinstance(data, Serializable) and all(
        hasattr(data, attr)
        for attr in ("getNumParticles", "getNumExceptions", "getParticleParameters")
    )


def _is_foreign_quantity(data: t.Any) -> bool:
    return (
        not isinstance(data, mmunit.Quantity)
        and hasattr(data, "value_in_unit")
        and hasattr(data, "unit")
    )


def convert_argument(data: t.Any) -> t.Any:  # pylint: disable=too-many-return-statements
    """Replace unserializable values by their serializable twins (recursively)."""
    if isinstance(data, np.integer):
        return int(data)
    if isinstance(data, np.floating):
        return float(data)
    if isinstance(data, mmunit.Quantity):
        return data if isinstance(data, Quantity) else Quantity(data)
    if _is_foreign_quantity(data):
        unit = mmunit.parse_unit(str(data.unit))
        return Quantity(mmunit.Quantity(data.value_in_unit(data.unit), unit))
    if isinstance(data, mmunit.Unit):
        return data if isinstance(data, Unit) else Unit(data)
    if _is_atom(data):
        return SerializableAtom(data)
    if _is_residue(data):
        return SerializableResidue(data)
    if _is_nonbonded_force(data):
        return SerializableForce(data)
    if isinstance(data, np.ndarray):
        return data
    if isinstance(data, t.Sequence) and not isinstance(data, str):
        if isinstance(data, tuple) and hasattr(data, "_fields"):  # namedtuple (Vec3)
            return type(data)(*map(convert_argument, data))
        if isinstance(data, range):
            return list(data)
        return type(data)(map(convert_argument, data))
    if isinstance(data, t.Dict):
        return type(data)((key, convert_argument(value)) for key, value in data.items())
    return data


def preprocess_args(func: t.Callable) -> t.Callable:
    """
    Decorator converting numpy scalars, quantities, units, atoms, residues and nonbonded forces in
    the arguments of ``func`` to serializable counterparts (``utils.py:187-247`` in the reference).

    Example
    -------
    >>> from cvpack_b200 import mmunit, units, utils
    >>> @utils.preprocess_args
    ... def function(data):
    ...     return data
    >>> assert isinstance(function(mmunit.angstrom), units.Unit)
    >>> assert isinstance(function(5 * mmunit.angstrom), units.Quantity)
    >>> seq = [mmunit.angstrom, mmunit.nanometer]
    >>> assert isinstance(function(seq), list)
    >>> assert all(isinstance(item, units.Unit) for item in function(seq))
    >>> dct = {"length": 3 * mmunit.angstrom, "time": 2 * mmunit.picosecond}
    >>> assert isinstance(function(dct), dict)
    >>> assert all(isinstance(item, units.Quantity) for item in function(dct).values())
    """
    signature = inspect.signature(func)

    @functools.wraps(func)
    def wrapper(*args: t.Any, **kwargs: t.Any) -> t.Any:
        bound = signature.bind(*args, **kwargs)
        for name, data in bound.arguments.items():
            bound.arguments[name] = convert_argument(data)
        return func(*bound.args, **bound.kwargs)

    return wrapper


def convert_to_matrix(array: t.Any) -> t.Tuple[np.ndarray, int, int]:
    """Convert a 1D or 2D array-like object to a 2D numpy array plus its shape."""
    array = np.atleast_2d(array)
    numrows, numcols, *other_dimensions = array.shape
    if other_dimensions:
        raise ValueError("Array-like object cannot have more than two dimensions.")
    return array, numrows, numcols


def as_index_list(atoms: t.Iterable[t.Any]) -> t.List[int]:
    """Atom indices as a list of Python ints (bit-exact integer handling; bools rejected)."""
    result = []
    for atom in atoms:
        if isinstance(atom, (bool, np.bool_)) or not isinstance(atom, (int, np.integer)):
            if hasattr(atom, "index") and isinstance(atom.index, (int, np.integer)):
                atom = atom.index
            else:
                raise TypeError(f"Atom indices must be integers; got {atom!r}")
        result.append(int(atom))
    return result


def find_cv_in_context(force: t.Any, context: t.Any) -> None:
    """
    Raise like ``get_single_force_state`` (``utils.py:132-142``) when the CV cannot be evaluated
    on its own in the given context.
    """
    forces_and_groups = [(f, f.getForceGroup()) for f in context.getSystem().getForces()]
    if not any(f is force for f, _ in forces_and_groups):
        raise RuntimeError("This force is not present in the given context.")


def evaluate_in_context(
    forces: t.Union[t.Any, t.Iterable[t.Any]], context: t.Any
) -> t.Union[float, t.Tuple[float, ...]]:
    """
    Evaluate collective variables at the first frame of a batch context without adding them to
    its system (the role of ``utils.evaluate_in_context``, ``utils.py:32-69``: a scratch system of
    unit-mass particles sharing the positions and box of ``context``).
    """
    is_single = not isinstance(forces, (list, tuple))
    force_list = [forces] if is_single else list(forces)
    values = [float(context._evaluateDetached(force)) for force in force_list]
    return values[0] if is_single else tuple(values)


def is_batch_context(obj: t.Any) -> bool:
    """Whether ``obj`` is a batch context (the analogue of ``isinstance(x, openmm.Context)``)."""
    return hasattr(obj, "_evaluateDetached") and hasattr(obj, "getSystem")


__all__ = [
    "preprocess_args",
    "convert_to_matrix",
    "evaluate_in_context",
    "as_index_list",
    "find_cv_in_context",
    "is_batch_context",
]

_ = mm  # re-exported for type references
