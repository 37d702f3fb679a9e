"""
.. class:: AttractionStrength (``/root/reference/cvpack/attraction_strength.py``).
"""

from __future__ import annotations

import typing as t

import numpy as np

from . import _native, mm, mmunit
from .collective_variable import CollectiveVariable
from .number_of_contacts import nonbonded_exclusions
from .units import Quantity, ScalarQuantity, value_in_md_units
from .utils import as_index_list, evaluate_in_context, is_batch_context

ONE_4PI_EPS0 = 138.93545764438198  # kJ nm/(mol e^2), the constant of attraction_strength.py:20


class AttractionStrength(CollectiveVariable):
    r"""
    The strength of the attraction between two atom groups (dimensionless),

    .. math::
        S_{\rm attr}({\bf r}) = -\frac{1}{E_{\rm ref}} \Big[
        \sum_{i \in {\bf g}_1} \sum_{j \in {\bf g}_2} U_{ij}
        - \sum_{i \in {\bf g}_1} \sum_{k \in {\bf g}_3} U^{(\alpha)}_{ik} \Big], \qquad
        U_{ij} = 4 \epsilon_{ij} \left(\frac{1}{y^2} - \frac{1}{y}\right)
        + \frac{\min(0, q_i q_j)}{4 \pi \varepsilon_0 r_c}
        \left(\frac{1}{x} + \frac{x^2 - 3}{2}\right),

    with :math:`y = |(r/\sigma_{ij})^6 - 2| + 2`, :math:`x = r/r_c`, Lorentz--Berthelot mixing, and
    charges / well depths of the contrast group :math:`{\bf g}_3` scaled by :math:`\alpha` /
    :math:`\alpha^2`.  Cutoff, switching and periodicity are those of ``nonbondedForce`` and all of
    its exceptions are excluded (``attraction_strength.py:188-224``).  The reference value, when a
    context is given, is computed without the contrast group (``:227-233``).

    Parameters
    ----------
    group1, group2
        Indices of the atoms in each group
    nonbondedForce
        Source of charges, sigmas, epsilons, exceptions, cutoff, switching and periodicity
    contrastGroup
        Indices of the atoms of the contrast group, or None
    reference
        The energy :math:`E_{\rm ref}` (number in kJ/mol or quantity), or a batch context at whose
        first frame it is evaluated
    contrastScaling
        The factor :math:`\alpha`
    name
        The name of the collective variable
    """

    def __init__(  # pylint: disable=too-many-arguments
        self,
        group1: t.Iterable[int],
        group2: t.Iterable[int],
        nonbondedForce: t.Any,
        contrastGroup: t.Optional[t.Iterable[int]] = None,
        reference: t.Union[ScalarQuantity, t.Any] = Quantity(1.0 * mmunit.kilojoule_per_mole),
        contrastScaling: float = 1.0,
        name: str = "attraction_strength",
    ) -> None:
        group1 = as_index_list(group1)
        group2 = as_index_list(group2)
        contrasting = contrastGroup is not None
        if contrasting:
            contrastGroup = as_index_list(contrastGroup)
        num_atoms = nonbondedForce.getNumParticles()
        self._num_atoms = num_atoms
        self._pbc = bool(nonbondedForce.usesPeriodicBoundaryConditions())
        self._cutoff = float(value_in_md_units(nonbondedForce.getCutoffDistance()))
        use_switch = bool(nonbondedForce.getUseSwitchingFunction())
        switch = float(value_in_md_units(nonbondedForce.getSwitchingDistance()))
        self._switch = switch if use_switch else None
        self._exclusions = nonbonded_exclusions(nonbondedForce)
        params = np.array(
            [
                [float(value_in_md_units(p)) for p in nonbondedForce.getParticleParameters(atom)]
                for atom in range(num_atoms)
            ],
            dtype=np.float64,
        ).reshape(num_atoms, 3)
        self._charge, self._sigma, self._epsilon = (params[:, k].copy() for k in range(3))
        self._sign = np.ones(num_atoms)
        if contrasting:
            in_group = np.zeros(num_atoms, dtype=bool)
            in_group[contrastGroup] = True
            self._charge[in_group] *= contrastScaling
            self._epsilon[in_group] *= contrastScaling**2
            self._sign[in_group] = -1.0
        self._group1 = group1
        self._group2 = group2
        # the normalisation comes from (group1, group2) alone: the contrast group joins after it
        self._active_group2 = group2
        self._reference = 1.0
        if is_batch_context(reference):
            reference = evaluate_in_context(self, reference)
        self._reference = float(value_in_md_units(reference))
        if contrasting:
            # two interaction groups (g1,g2) and (g1,g3) of one force = one group (g1, g2 U g3)
            self._active_group2 = sorted(set(group2) | set(contrastGroup))
        self._registerCV(
            name,
            mmunit.dimensionless,
            group1=group1,
            group2=group2,
            nonbondedForce=nonbondedForce,
            contrastGroup=contrastGroup,
            reference=self._reference,
            contrastScaling=contrastScaling,
        )

    def usesPeriodicBoundaryConditions(self) -> bool:
        return self._pbc

    def _createNative(self, system: mm.System) -> _native.NativeCV:
        group1 = _native.as_i32(self._group1)
        group2 = _native.as_i32(self._active_group2)
        charge, sigma = _native.as_f64(self._charge), _native.as_f64(self._sigma)
        epsilon, sign = _native.as_f64(self._epsilon), _native.as_f64(self._sign)
        desc = _native.PairSumDesc(
            kind=_native.CVP_PAIR_ATTRACTION,
            num_atoms=self._num_atoms,
            n1=len(group1),
            group1=_native.ptr_i32(group1),
            n2=len(group2),
            group2=_native.ptr_i32(group2),
            num_exclusions=len(self._exclusions),
            exclusions=_native.ptr_i32(self._exclusions),
            periodic=int(self._pbc),
            cutoff=self._cutoff,
            switch_distance=-1.0 if self._switch is None else self._switch,
            scale=1.0 / self._reference,
            charge=_native.ptr_f64(charge),
            sigma=_native.ptr_f64(sigma),
            epsilon=_native.ptr_f64(epsilon),
            sign=_native.ptr_f64(sign),
            coulomb_constant=ONE_4PI_EPS0,
        )
        keep = (group1, group2, self._exclusions, charge, sigma, epsilon, sign)
        return _native.NativeCV.create("cvp_pair_sum_create", desc, keep)


AttractionStrength.registerTag("!cvpack.AttractionStrength")
