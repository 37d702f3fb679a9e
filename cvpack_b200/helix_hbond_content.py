"""
.. class:: HelixHBondContent (``/root/reference/cvpack/helix_hbond_content.py``).
"""

from __future__ import annotations

import re as regex
import typing as t

from . import _native, mm, mmunit
from .collective_variable import CollectiveVariable
from .distance import bonded_spec
from .units import Quantity, ScalarQuantity, value_in_md_units


class HelixHBondContent(CollectiveVariable):
    r"""
    The alpha-helix hydrogen-bond content of a sequence of :math:`n` residues (dimensionless),

    .. math::
        \alpha_{\rm hb}({\bf r}) = \sum_{i=5}^{n}
        B_m\!\left(\frac{\|{\bf r}_{\mathrm{H}_i} - {\bf r}_{\mathrm{O}_{i-4}}\|}{d_0}\right),
        \qquad B_m(x) = \frac{1}{1 + x^{2m}},

    between the backbone oxygen of residue :math:`i-4` and the amide hydrogen of residue
    :math:`i`, divided by :math:`n-4` if ``normalize`` is on.  Atoms are found by name with the
    regular expressions of ``helix_hbond_content.py:105-106`` in the reference.

    Parameters
    ----------
    residues
        The residues of the sequence, contiguous and ordered from N- to C-terminus
    pbc
        Whether to use the minimum-image convention
    thresholdDistance
        :math:`d_0`
    halfExponent
        :math:`m`
    normalize
        Whether to divide by the number of hydrogen bonds
    name
        The name of the collective variable

    Raises
    ------
    ValueError
        If a residue lacks the backbone oxygen / amide hydrogen that is needed
    """

    def __init__(  # pylint: disable=too-many-arguments
        self,
        residues: t.Sequence[t.Any],
        pbc: bool = False,
        thresholdDistance: ScalarQuantity = Quantity(0.33 * mmunit.nanometers),
        halfExponent: int = 3,
        normalize: bool = False,
        name: str = "helix_hbond_content",
    ) -> None:
        def find_atom(residue: t.Any, pattern: t.Pattern) -> int:
            for atom in residue.atoms():
                if regex.match(pattern, atom.name):
                    return atom.index
            raise ValueError(
                f"Could not find atom matching regex "
                f"'{pattern.pattern}'"
                f" in residue {residue.name}{residue.id}"
            )

        self._coefficient = 1 / (len(residues) - 4) if normalize else 1
        self._threshold = float(value_in_md_units(thresholdDistance))
        self._half_exponent = int(halfExponent)
        self._pbc = bool(pbc)
        hydrogen_pattern = regex.compile(r"\b(H|1H|HN1|HT1|H1|HN)\b")
        oxygen_pattern = regex.compile(r"\b(O|OCT1|OC1|OT1|O1)\b")
        self._terms = [
            [find_atom(residues[i - 4], oxygen_pattern), find_atom(residues[i], hydrogen_pattern)]
            for i in range(4, len(residues))
        ]
        self._registerCV(
            name,
            mmunit.dimensionless,
            residues=residues,
            pbc=pbc,
            thresholdDistance=thresholdDistance,
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
            reference=[0.0] * len(self._terms),
            tolerance=self._threshold,
            half_exponent=self._half_exponent,
            coefficient=self._coefficient,
        )


HelixHBondContent.registerTag("!cvpack.HelixHBondContent")
