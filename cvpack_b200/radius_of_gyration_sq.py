"""
.. class:: RadiusOfGyrationSq (``/root/reference/cvpack/radius_of_gyration_sq.py``).
"""

from __future__ import annotations

import typing as t

from . import mmunit
from .base_radius_of_gyration import BaseRadiusOfGyration


class RadiusOfGyrationSq(BaseRadiusOfGyration):
    r"""
    The square of the radius of gyration of a group of :math:`n` atoms (nm\ :sup:`2`),

    .. math::
        r_g^2 = \frac{1}{n} \sum_{i=1}^n \|{\bf r}_i - {\bf r}_c\|^2 .

    The reference builds one two-group bond per atom and passes the *atom index* where a group index
    is expected (``radius_of_gyration_sq.py:88-89``), so it is only correct when the group is
    ``0..n-1``; this class implements the formula for any group and agrees with the reference
    wherever the reference is valid.

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

    _squared = True

    def __init__(
        self,
        group: t.Iterable[int],
        pbc: bool = False,
        weighByMass: bool = False,
        name: str = "radius_of_gyration_sq",
    ) -> None:
        group = self._setGroup(group, pbc, weighByMass)
        self._registerCV(
            name, mmunit.nanometers**2, group=group, pbc=pbc, weighByMass=weighByMass
        )


RadiusOfGyrationSq.registerTag("!cvpack.RadiusOfGyrationSq")
