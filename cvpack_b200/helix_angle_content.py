"""
.. class:: HelixAngleContent (``/root/reference/cvpack/helix_angle_content.py``).
"""

from __future__ import annotations

import typing as t

from . import _native, mm, mmunit
from .collective_variable import CollectiveVariable
from .distance import bonded_spec
from .units import Quantity, ScalarQuantity, value_in_md_units


class HelixAngleContent(CollectiveVariable):
    r"""
    The alpha-helix angle content of a sequence of :math:`n` residues (dimensionless),

    .. math::
        \alpha_{\theta}({\bf r}) = \sum_{i=2}^{n-1}
        B_m\!\left(\frac{\theta(\mathrm{C}^\alpha_{i-1}, \mathrm{C}^\alpha_i,
        \mathrm{C}^\alpha_{i+1}) - \theta_{\rm ref}}{\theta_{\rm tol}}\right),
        \qquad B_m(x) = \frac{1}{1 + x^{2m}},

    divided by :math:`n-2` if ``normalize`` is on.

    Parameters
    ----------
    residues
        The residues of the sequence, contiguous and ordered from N- to C-terminus
    pbc
        Whether to use the minimum-image convention
    thetaReference
        :math:`\theta_{\rm ref}`
    tolerance
        :math:`\theta_{\rm tol}`
    halfExponent
        :math:`m`
    normalize
        Whether to divide by the number of angles
    name
        The name of the collective variable

    Raises
    ------
    ValueError
        If a residue has no CA atom
    """

    def __init__(  # pylint: disable=too-many-arguments
        self,
        residues: t.Sequence[t.Any],
        pbc: bool = False,
        thetaReference: ScalarQuantity = Quantity(88 * mmunit.degrees),
        tolerance: ScalarQuantity = Quantity(15 * mmunit.degrees),
        halfExponent: int = 3,
        normalize: bool = False,
        name: str = "helix_angle_content",
    ) -> None:
        def find_alpha_carbon(residue: t.Any) -> int:
            for atom in residue.atoms():
                if atom.name == "CA":
                    return atom.index
            raise ValueError(f"Could not find atom CA in residue {residue.name}{residue.id}")

        num_angles = len(residues) - 2
        self._coefficient = 1 / num_angles if normalize else 1
        self._theta0 = float(value_in_md_units(thetaReference))
        self._tolerance = float(value_in_md_units(tolerance))
        self._half_exponent = int(halfExponent)
        self._pbc = bool(pbc)
        self._terms = [
            [
                find_alpha_carbon(residues[i - 1]),
                find_alpha_carbon(residues[i]),
                find_alpha_carbon(residues[i + 1]),
            ]
            for i in range(1, len(residues) - 1)
        ]
        self._registerCV(
            name,
            mmunit.dimensionless,
            residues=residues,
            pbc=pbc,
            thetaReference=thetaReference,
            tolerance=tolerance,
            halfExponent=halfExponent,
            normalize=normalize,
        )

    def usesPeriodicBoundaryConditions(self) -> bool:
        return self._pbc

    def _createNative(self, system: mm.System) -> _native.NativeCV:
        return bonded_spec(
            system.getNumParticles(),
            self._terms,
            self._pbc,
            function=_native.CVP_TERM_BOXCAR,
            reference=[self._theta0] * len(self._terms),
            tolerance=self._tolerance,
            half_exponent=self._half_exponent,
            coefficient=self._coefficient,
        )


HelixAngleContent.registerTag("!cvpack.HelixAngleContent")
