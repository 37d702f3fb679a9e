"""
.. class:: HelixTorsionContent (``/root/reference/cvpack/helix_torsion_content.py``).
"""

from __future__ import annotations

import typing as t

from . import _native, mm, mmunit
from .collective_variable import CollectiveVariable
from .distance import bonded_spec
from .units import Quantity, ScalarQuantity, value_in_md_units


class HelixTorsionContent(CollectiveVariable):
    r"""
    The alpha-helix Ramachandran content of a sequence of :math:`n` residues (dimensionless),

    .. math::
        \alpha_{\phi,\psi}({\bf r}) = \frac{1}{2} \sum_{i=2}^{n-1} \left[
        B_m\!\left(\frac{\phi_i - \phi_{\rm ref}}{\theta_{\rm tol}}\right) +
        B_m\!\left(\frac{\psi_i - \psi_{\rm ref}}{\theta_{\rm tol}}\right) \right],
        \qquad B_m(x) = \frac{1}{1 + x^{2m}},

    with :math:`\phi_i` = (C\ :sub:`i-1`, N\ :sub:`i`, CA\ :sub:`i`, C\ :sub:`i`) and
    :math:`\psi_i` = (N\ :sub:`i`, CA\ :sub:`i`, C\ :sub:`i`, N\ :sub:`i+1`).  As in the reference
    the difference to the reference angle is **not** wrapped periodically
    (``helix_torsion_content.py:128``), and ``normalize`` -- which divides by :math:`n-2` -- is not
    part of the serialized state (``:147-156``).

    Parameters
    ----------
    residues
        The residues of the sequence, contiguous and ordered from N- to C-terminus
    pbc
        Whether to use the minimum-image convention
    phiReference, psiReference
        Reference values of the two backbone dihedrals
    tolerance
        :math:`\theta_{\rm tol}`
    halfExponent
        :math:`m`
    normalize
        Whether to divide by the number of (phi, psi) pairs
    name
        The name of the collective variable

    Raises
    ------
    ValueError
        If a residue lacks one of the atoms N, CA, C
    """

    def __init__(  # pylint: disable=too-many-arguments
        self,
        residues: t.Sequence[t.Any],
        pbc: bool = False,
        phiReference: ScalarQuantity = Quantity(-63.8 * mmunit.degrees),
        psiReference: ScalarQuantity = Quantity(-41.1 * mmunit.degrees),
        tolerance: ScalarQuantity = Quantity(25 * mmunit.degrees),
        halfExponent: int = 3,
        normalize: bool = False,
        name: str = "helix_torsion_content",
    ) -> None:
        def find_atom(residue: t.Any, atom_name: str) -> int:
            for atom in residue.atoms():
                if atom.name == atom_name:
                    return atom.index
            raise ValueError(f"Could not find atom {atom_name} in residue {residue.name}{residue.id}")

        num_residues = len(residues)
        self._coefficient = 1 / (2 * (num_residues - 2)) if normalize else 1 / 2
        self._tolerance = float(value_in_md_units(tolerance))
        phi, psi = float(value_in_md_units(phiReference)), float(value_in_md_units(psiReference))
        self._half_exponent = int(halfExponent)
        self._pbc = bool(pbc)
        self._terms, self._references = [], []
        for i in range(1, num_residues - 1):
            self._terms.append(
                [
                    find_atom(residues[i - 1], "C"),
                    find_atom(residues[i], "N"),
                    find_atom(residues[i], "CA"),
                    find_atom(residues[i], "C"),
                ]
            )
            self._references.append(phi)
            self._terms.append(
                [
                    find_atom(residues[i], "N"),
                    find_atom(residues[i], "CA"),
                    find_atom(residues[i], "C"),
                    find_atom(residues[i + 1], "N"),
                ]
            )
            self._references.append(psi)
        self._registerCV(
            name,
            mmunit.dimensionless,
            residues=residues,
            pbc=pbc,
            phiReference=phiReference,
            psiReference=psiReference,
            tolerance=tolerance,
            halfExponent=halfExponent,
        )

    def usesPeriodicBoundaryConditions(self) -> bool:
        return self._pbc

    def _createNative(self, system: mm.System) -> _native.NativeCV:
        return bonded_spec(
            system.getNumParticles(),
            self._terms,
            self._pbc,
            function=_native.CVP_TERM_BOXCAR,
            reference=self._references,
            tolerance=self._tolerance,
            half_exponent=self._half_exponent,
            coefficient=self._coefficient,
        )


HelixTorsionContent.registerTag("!cvpack.HelixTorsionContent")
