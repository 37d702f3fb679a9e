"""
.. class:: Torsion -- the dihedral formed by four atoms (``/root/reference/cvpack/torsion.py``).
"""

from __future__ import annotations

import numpy as np

from . import _native, mm, mmunit
from .collective_variable import CollectiveVariable
from .distance import bonded_spec


class Torsion(CollectiveVariable):
    r"""
    The torsion angle of four atoms (rad, periodic in :math:`(-\pi, \pi)`),

    .. math::
        \varphi = \mathrm{atan2}\big(({\bf r}_{21} \times {\bf r}_{34}) \cdot {\bf u}_{23},\;
        {\bf r}_{21} \cdot {\bf r}_{34} - ({\bf r}_{21} \cdot {\bf u}_{23})
        ({\bf r}_{34} \cdot {\bf u}_{23})\big),

    with :math:`{\bf r}_{ij} = {\bf r}_j - {\bf r}_i` and :math:`{\bf u}_{23}` the unit vector
    along the central bond -- the sign convention of ``torsion.py:22-33`` in the reference
    (``[[0,0,0],[1,0,0],[1,1,0],[1,1,1]]`` gives :math:`+\pi/2`).

    Parameters
    ----------
    atom1, atom2, atom3, atom4
        Indices of the four atoms
    pbc
        Whether to use the minimum-image convention
    name
        The name of the collective variable
    """

    def __init__(  # pylint: disable=too-many-arguments
        self,
        atom1: int,
        atom2: int,
        atom3: int,
        atom4: int,
        pbc: bool = False,
        name: str = "torsion",
    ) -> None:
        self._atoms = (int(atom1), int(atom2), int(atom3), int(atom4))
        self._pbc = bool(pbc)
        self._registerCV(
            name, mmunit.radians, atom1=atom1, atom2=atom2, atom3=atom3, atom4=atom4, pbc=pbc
        )
        self._registerPeriodicBounds(-np.pi, np.pi)

    def usesPeriodicBoundaryConditions(self) -> bool:
        return self._pbc

    def _createNative(self, system: mm.System) -> _native.NativeCV:
        return bonded_spec(system.getNumParticles(), [self._atoms], self._pbc)


Torsion.registerTag("!cvpack.Torsion")
