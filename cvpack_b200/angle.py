"""
.. class:: Angle -- the angle formed by three atoms (``/root/reference/cvpack/angle.py``).
"""

from __future__ import annotations

import numpy as np

from . import _native, mm, mmunit
from .collective_variable import CollectiveVariable
from .distance import bonded_spec


class Angle(CollectiveVariable):
    r"""
    The angle at ``atom2`` formed with ``atom1`` and ``atom3``,
    :math:`\theta = \arccos(\hat{\bf r}_{21} \cdot \hat{\bf r}_{23})` (rad), with periodic bounds
    :math:`(-\pi, \pi)` registered as in ``angle.py:78`` of the reference.

    Parameters
    ----------
    atom1, atom2, atom3
        Indices of the three atoms
    pbc
        Whether to use the minimum-image convention
    name
        The name of the collective variable
    """

    def __init__(
        self, atom1: int, atom2: int, atom3: int, pbc: bool = False, name: str = "angle"
    ) -> None:
        self._atoms = (int(atom1), int(atom2), int(atom3))
        self._pbc = bool(pbc)
        self._registerCV(name, mmunit.radians, atom1=atom1, atom2=atom2, atom3=atom3, pbc=pbc)
        self._registerPeriodicBounds(-np.pi, np.pi)

    def usesPeriodicBoundaryConditions(self) -> bool:
        return self._pbc

    def _createNative(self, system: mm.System) -> _native.NativeCV:
        return bonded_spec(system.getNumParticles(), [self._atoms], self._pbc)


Angle.registerTag("!cvpack.Angle")
