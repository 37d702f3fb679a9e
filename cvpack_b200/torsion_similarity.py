"""
.. class:: TorsionSimilarity -- the degree of similarity between pairs of torsion angles
   (``/root/reference/cvpack/torsion_similarity.py``).
"""

from __future__ import annotations

import typing as t

from . import _native, mm, mmunit
from .collective_variable import CollectiveVariable
from .distance import bonded_spec


class TorsionSimilarity(CollectiveVariable):
    r"""
    The degree of similarity between :math:`n` pairs of torsion angles,

    .. math::
        s({\bf r}) = \frac{n}{2} + \frac{1}{2} \sum_{i=1}^n
        \cos\big(\phi^{\rm 1st}_i({\bf r}) - \phi^{\rm 2nd}_i({\bf r})\big),

    PLUMED's ``DIHCOR``.  The reference evaluates ``0.5*(1 + cos(min(delta, 2*pi - delta)))`` with
    ``delta = dihedral(p1..p4) - dihedral(p5..p8)`` in a ``CustomCompoundBondForce``
    (``torsion_similarity.py:89-93``); the cosine makes the ``min`` immaterial for the value and for
    the gradient alike.  Dimensionless.

    As in the reference, ``pbc`` is not part of the registered (serialized) arguments
    (``torsion_similarity.py:94-99``): a deserialized copy does not use periodic boundary
    conditions.

    Parameters
    ----------
    firstList
        :math:`n` tuples of four atom indices: the first torsion of each pair
    secondList
        :math:`n` tuples of four atom indices: the second torsion of each pair
    pbc
        Whether to use the minimum-image convention
    name
        The name of the collective variable
    """

    def __init__(
        self,
        firstList: t.Iterable[t.Sequence[int]],
        secondList: t.Iterable[t.Sequence[int]],
        pbc: bool = False,
        name: str = "torsion_similarity",
    ) -> None:
        firstList = [list(map(int, first)) for first in firstList]
        secondList = [list(map(int, second)) for second in secondList]
        for torsion in firstList + secondList:
            if len(torsion) != 4:
                raise ValueError("Every torsion is defined by four atom indices")
        # zip() like the reference: surplus entries of the longer list are ignored
        self._terms = [first + second for first, second in zip(firstList, secondList)]
        self._pbc = bool(pbc)
        self._registerCV(name, mmunit.dimensionless, firstList=firstList, secondList=secondList)

    def usesPeriodicBoundaryConditions(self) -> bool:
        return self._pbc

    def _createNative(self, system: mm.System) -> _native.NativeCV:
        if not self._terms:
            raise ValueError("TorsionSimilarity needs at least one pair of torsions")
        return bonded_spec(
            system.getNumParticles(), self._terms, self._pbc, function=_native.CVP_TERM_HALF_COSINE
        )


TorsionSimilarity.registerTag("!cvpack.TorsionSimilarity")
