"""
.. class:: ShortestDistance (``/root/reference/cvpack/shortest_distance.py``).
"""

from __future__ import annotations

import typing as t

from . import _native, mm, mmunit
from .collective_variable import CollectiveVariable
from .units import Quantity, value_in_md_units
from .utils import as_index_list


class ShortestDistance(CollectiveVariable):
    r"""
    A smooth approximation of the shortest distance between two atom groups (nm),

    .. math::
        r_{\rm min}({\bf r}) = \frac{r_c e^{-\beta} + \sum_{i \in {\bf g}_1} \sum_{j \in {\bf g}_2}
        r_{ij}\, e^{-\beta r_{ij}/r_c}}{e^{-\beta} + \sum_{i \in {\bf g}_1} \sum_{j \in {\bf g}_2}
        e^{-\beta r_{ij}/r_c}}, \qquad r_{ij} < r_c,

    which is the form the reference's code and test compute (``shortest_distance.py:122-127``,
    ``tests/test_cvpack.py:919-927``); the larger :math:`\beta`, the closer to the true minimum.
    Atoms of the two groups are never excluded from each other, so the groups must not contain
    pairs that should not count (e.g. bonded atoms).

    Parameters
    ----------
    group1, group2
        Indices of the atoms in each group
    numAtoms
        The total number of atoms in the system
    beta
        The sharpness parameter :math:`\beta`
    cutoffDistance
        The cutoff distance :math:`r_c`
    pbc
        Whether to use the minimum-image convention
    name
        The name of the collective variable
    """

    def __init__(  # pylint: disable=too-many-arguments
        self,
        group1: t.Sequence[int],
        group2: t.Sequence[int],
        numAtoms: int,
        beta: float = 100,
        cutoffDistance: t.Any = Quantity(1 * mmunit.nanometers),
        pbc: bool = True,
        name: str = "shortest_distance",
    ) -> None:
        group1 = as_index_list(group1)
        group2 = as_index_list(group2)
        rc = float(value_in_md_units(cutoffDistance))
        self._groups = (group1, group2)
        self._num_atoms = int(numAtoms)
        self._beta = float(beta)
        self._rc = rc
        self._pbc = bool(pbc)
        self._registerCV(
            name,
            mmunit.nanometers,
            group1=group1,
            group2=group2,
            numAtoms=numAtoms,
            beta=beta,
            # (the reference registers the bare number it got from the quantity: `1`, not `1.0`)
            cutoffDistance=value_in_md_units(cutoffDistance),
            pbc=pbc,
        )

    def usesPeriodicBoundaryConditions(self) -> bool:
        return self._pbc

    def _createNative(self, system: mm.System) -> _native.NativeCV:
        group1 = _native.as_i32(self._groups[0])
        group2 = _native.as_i32(self._groups[1])
        desc = _native.PairSumDesc(
            kind=_native.CVP_PAIR_SOFTMIN,
            num_atoms=self._num_atoms,
            n1=len(group1),
            group1=_native.ptr_i32(group1),
            n2=len(group2),
            group2=_native.ptr_i32(group2),
            periodic=int(self._pbc),
            cutoff=self._rc,
            switch_distance=-1.0,
            scale=1.0,
            beta=self._beta,
        )
        return _native.NativeCV.create("cvp_pair_sum_create", desc, (group1, group2))


ShortestDistance.registerTag("!cvpack.ShortestDistance")
