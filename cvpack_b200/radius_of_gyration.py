"""
.. class:: RadiusOfGyration (``/root/reference/cvpack/radius_of_gyration.py``).
"""

from __future__ import annotations

import typing as t

from . import mmunit
from .base_radius_of_gyration import BaseRadiusOfGyration


class RadiusOfGyration(BaseRadiusOfGyration):
    r"""
    The radius of gyration of a group of :math:`n` atoms (nm),

    .. math::
        r_g = \sqrt{\frac{1}{n} \sum_{i=1}^n \|{\bf r}_i - {\bf r}_c\|^2},

    where :math:`{\bf r}_c` is the geometric centre of the group, or its centre of mass if
    ``weighByMass`` is on (the sum itself stays unweighted, as in the reference).  The reference
    warns that its implementation "lacks parallelization" for large groups
    (``radius_of_gyration.py:42-45``); here both this CV and :class:`RadiusOfGyrationSq` run the
    same streaming reduction kernel, so there is no reason to prefer one over the other.

    Parameters
    ----------
    group
        Indices of the atoms in the group
    pbc
        Whether to apply the minimum-image convention to each :math:`{\bf r}_i - {\bf r}_c`
    weighByMass
        Whether :math:`{\bf r}_c` is the centre of mass rather than the geometric centre
    name
        The name of the collective variable
    """

    _squared = False

    def __init__(
        self,
        group: t.Iterable[int],
        pbc: bool = False,
        weighByMass: bool = False,
        name: str = "radius_of_gyration",
    ) -> None:
        group = self._setGroup(group, pbc, weighByMass)
        self._registerCV(name, mmunit.nanometers, group=group, pbc=pbc, weighByMass=weighByMass)


RadiusOfGyration.registerTag("!cvpack.RadiusOfGyration")
