"""
.. class:: ResidueCoordination (``/root/reference/cvpack/residue_coordination.py``).
"""

from __future__ import annotations

import typing as t

import numpy as np

from . import _native, mm, mmunit
from . import app as mmapp
from .collective_variable import CollectiveVariable
from .step_functions import parse_step_function
from .units import Quantity, ScalarQuantity, value_in_md_units


class ResidueCoordination(CollectiveVariable):
    r"""
    The number of contacts between two disjoint groups of residues (dimensionless),

    .. math::
        N({\bf r}) = \frac{1}{N_{\rm ref}} \sum_{I \in {\bf G}_1} \sum_{J \in {\bf G}_2}
        S\!\left(\frac{\|{\bf R}_J - {\bf R}_I\|}{r_0}\right),

    where :math:`{\bf R}` is the centroid of a residue -- centre of mass if ``weighByMass``, else
    geometric centre, hydrogens optionally left out -- and :math:`N_{\rm ref}` is the number of
    residue pairs if ``normalize`` is on, else 1.  There is no cutoff.  Centroids are formed from
    the raw coordinates; periodic boundary conditions, if on, apply to the centroid--centroid
    vector only (``residue_coordination.py:129-144`` in the reference).

    Parameters
    ----------
    residueGroup1, residueGroup2
        The residues of each group (objects with ``atoms()`` yielding atoms with ``index`` and
        ``element``)
    pbc
        Whether to use the minimum-image convention
    stepFunction
        One of the whitelisted forms of :mod:`cvpack_b200.step_functions`
    thresholdDistance
        :math:`r_0`
    normalize
        Whether to divide by the number of residue pairs
    weighByMass
        Whether residue centroids are centres of mass
    includeHydrogens
        Whether hydrogen atoms take part in the centroids
    name
        The name of the collective variable
    """

    def __init__(  # pylint: disable=too-many-arguments
        self,
        residueGroup1: t.Iterable[t.Any],
        residueGroup2: t.Iterable[t.Any],
        pbc: bool = True,
        stepFunction: str = "1/(1+x^6)",
        thresholdDistance: ScalarQuantity = Quantity(1.0 * mmunit.nanometers),
        normalize: bool = False,
        weighByMass: bool = True,
        includeHydrogens: bool = True,
        name: str = "residue_coordination",
    ) -> None:
        residueGroup1 = list(residueGroup1)
        residueGroup2 = list(residueGroup2)
        nr1, nr2 = len(residueGroup1), len(residueGroup2)
        self._counts = (nr1, nr2)
        self._ref_val = nr1 * nr2 if normalize else 1.0
        self._r0 = float(value_in_md_units(thresholdDistance))
        self._step = parse_step_function(stepFunction)
        self._pbc = bool(pbc)
        self._weigh_by_mass = bool(weighByMass)
        self._groups = []
        for res in residueGroup1 + residueGroup2:
            atoms = [
                int(atom.index)
                for atom in res.atoms()
                if includeHydrogens or not _is_hydrogen(atom)
            ]
            self._groups.append(atoms)
        self._registerCV(
            name,
            mmunit.dimensionless,
            residueGroup1=residueGroup1,
            residueGroup2=residueGroup2,
            pbc=pbc,
            stepFunction=stepFunction,
            thresholdDistance=thresholdDistance,
            normalize=normalize,
            weighByMass=weighByMass,
            includeHydrogens=includeHydrogens,
        )

    def usesPeriodicBoundaryConditions(self) -> bool:
        return self._pbc

    def getReferenceValue(self) -> mmunit.Quantity:
        """Get the reference value used for normalizing the residue coordination."""
        return self._ref_val * self.getUnit()

    def setReferenceValue(self, value: ScalarQuantity) -> None:
        """
        Set the reference value used for normalizing the residue coordination.  Takes effect in
        contexts (re)initialized afterwards, like the reference's ``setEnergyFunction`` edit
        (``residue_coordination.py:169-183``).
        """
        self._ref_val = float(value_in_md_units(value))

    def _createNative(self, system: mm.System) -> _native.NativeCV:
        masses = system.getMasses()
        offsets, atoms, weights = [0], [], []
        for group in self._groups:
            if not group:
                raise ValueError("A residue has no atoms to define its centroid.")
            if self._weigh_by_mass:
                w = masses[group]
                w = w / w.sum()
            else:
                w = np.full(len(group), 1.0 / len(group))
            atoms.extend(group)
            weights.extend(w.tolist())
            offsets.append(len(atoms))
        offsets = _native.as_i32(offsets)
        atoms = _native.as_i32(atoms)
        weights = _native.as_f64(weights)
        kind, p0, p1 = self._step
        desc = _native.CentroidPairSumDesc(
            num_atoms=system.getNumParticles(),
            n1=self._counts[0],
            n2=self._counts[1],
            group_offsets=_native.ptr_i32(offsets),
            group_atoms=_native.ptr_i32(atoms),
            group_weights=_native.ptr_f64(weights),
            periodic=int(self._pbc),
            step_kind=kind,
            step_p0=p0,
            step_p1=p1,
            r0=self._r0,
            scale=1.0 / self._ref_val,
        )
        return _native.NativeCV.create(
            "cvp_centroid_pair_sum_create", desc, (offsets, atoms, weights)
        )


def _is_hydrogen(atom: t.Any) -> bool:
    element = atom.element
    if element is None:
        return False
    return element is mmapp.hydrogen or getattr(element, "symbol", None) == "H"


ResidueCoordination.registerTag("!cvpack.ResidueCoordination")
