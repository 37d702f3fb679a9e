"""
.. class:: Distance -- the distance between two atoms (``/root/reference/cvpack/distance.py``).
"""

from __future__ import annotations

import typing as t

from . import _native, mm, mmunit
from .collective_variable import CollectiveVariable


def bonded_spec(  # pylint: disable=too-many-arguments
    num_atoms: int,
    terms: t.Sequence[t.Sequence[int]],
    pbc: bool,
    function: int = _native.CVP_TERM_IDENTITY,
    reference: t.Optional[t.Sequence[float]] = None,
    tolerance: float = 1.0,
    half_exponent: int = 1,
    coefficient: float = 1.0,
) -> _native.NativeCV:
    """Kernel spec for a sum of bond / angle / torsion terms (``cvp_bonded_create``)."""
    atoms = _native.as_i32(terms)
    refs = None if reference is None else _native.as_f64(reference)
    desc = _native.BondedDesc(
        num_atoms=num_atoms,
        num_terms=atoms.shape[0],
        atoms_per_term=atoms.shape[1],
        atoms=_native.ptr_i32(atoms),
        function=function,
        reference=_native.ptr_f64(refs),
        tolerance=float(tolerance),
        half_exponent=int(half_exponent),
        coefficient=float(coefficient),
        periodic=int(bool(pbc)),
    )
    return _native.NativeCV.create("cvp_bonded_create", desc, (atoms, refs))


class Distance(CollectiveVariable):
    r"""
    The distance between two atoms, :math:`d = \|{\bf r}_2 - {\bf r}_1\|` (nm).

    Parameters
    ----------
    atom1, atom2
        Indices of the two atoms
    pbc
        Whether to use the minimum-image convention
    name
        The name of the collective variable

    Example (the known-answer case of ``distance.py:37-50``)
    -------
    >>> import cvpack_b200 as cvpack
    >>> cvpack.Distance(0, 1).getUnit()
    nm
    """

    def __init__(self, atom1: int, atom2: int, pbc: bool = False, name: str = "distance") -> None:
        self._atoms = (int(atom1), int(atom2))
        self._pbc = bool(pbc)
        self._registerCV(name, mmunit.nanometers, atom1=atom1, atom2=atom2, pbc=pbc)

    def usesPeriodicBoundaryConditions(self) -> bool:
        return self._pbc

    def _createNative(self, system: mm.System) -> _native.NativeCV:
        return bonded_spec(system.getNumParticles(), [self._atoms], self._pbc)


Distance.registerTag("!cvpack.Distance")
