"""
.. class:: NumberOfContacts (``/root/reference/cvpack/number_of_contacts.py``).
"""

from __future__ import annotations

import typing as t
from numbers import Real

import numpy as np

from . import _native, mm, mmunit
from .collective_variable import CollectiveVariable
from .step_functions import parse_step_function
from .units import Quantity, ScalarQuantity, value_in_md_units
from .utils import as_index_list, evaluate_in_context, is_batch_context


def nonbonded_exclusions(nonbondedForce: t.Any) -> np.ndarray:
    """
    Every exception of a nonbonded force as an ``[E, 2]`` int32 array -- all of them become
    exclusions, scaled 1-4 pairs included (``number_of_contacts.py:53-56,149-151``).
    """
    pairs = []
    for index in range(nonbondedForce.getNumExceptions()):
        i, j, *_ = nonbondedForce.getExceptionParameters(index)
        pairs.append((int(i), int(j)))
    return _native.as_i32(pairs).reshape(-1, 2)


class NumberOfContacts(CollectiveVariable):
    r"""
    The number of contacts between two atom groups (dimensionless),

    .. math::
        N({\bf r}) = \frac{1}{N_{\rm ref}} \sum_{i \in {\bf g}_1} \sum_{j \in {\bf g}_2}
        S\!\left(\frac{\|{\bf r}_j - {\bf r}_i\|}{r_0}\right),

    where :math:`S` is a step function, by default :math:`S(x) = 1/(1+x^6)`, truncated at
    :math:`r_c = f_c r_0` and, unless ``switchFactor`` is None, brought to zero smoothly between
    :math:`f_s r_0` and :math:`r_c` by the quintic :math:`1 - 10t^3 + 15t^4 - 6t^5`.

    The sum follows the interaction-group rules of the OpenMM force the reference configures
    (``number_of_contacts.py:47-56``): self pairs are skipped, a pair of atoms that are both in
    both groups is counted once, and every pair that is an *exception* of ``nonbondedForce`` is
    left out.  Periodic boundary conditions are used iff ``nonbondedForce`` uses them.

    Parameters
    ----------
    group1, group2
        Indices of the atoms in each group
    nonbondedForce
        The nonbonded force that supplies the number of atoms, the exceptions and whether to use
        periodic boundary conditions (:class:`cvpack_b200.mm.NonbondedForce` or the OpenMM original)
    reference
        A number to divide the sum by, or a :class:`~cvpack_b200.batch.BatchContext` at whose first
        frame the un-normalised sum is evaluated to get that number
    stepFunction
        One of the whitelisted forms of :mod:`cvpack_b200.step_functions`
    thresholdDistance
        :math:`r_0`
    cutoffFactor
        :math:`f_c`
    switchFactor
        :math:`f_s`, or None for a sharp cutoff
    name
        The name of the collective variable
    """

    def __init__(  # pylint: disable=too-many-arguments
        self,
        group1: t.Sequence[int],
        group2: t.Sequence[int],
        nonbondedForce: t.Any,
        reference: t.Union[Real, t.Any] = 1.0,
        stepFunction: str = "1/(1+x^6)",
        thresholdDistance: ScalarQuantity = Quantity(0.3 * mmunit.nanometers),
        cutoffFactor: float = 2.0,
        switchFactor: t.Optional[float] = 1.5,
        name: str = "number_of_contacts",
    ) -> None:
        group1 = as_index_list(group1)
        group2 = as_index_list(group2)
        self._num_atoms = nonbondedForce.getNumParticles()
        self._pbc = bool(nonbondedForce.usesPeriodicBoundaryConditions())
        self._r0 = float(value_in_md_units(thresholdDistance))
        self._step = parse_step_function(stepFunction)
        self._groups = (group1, group2)
        self._exclusions = nonbonded_exclusions(nonbondedForce)
        self._cutoff = float(cutoffFactor) * self._r0
        self._switch = None if switchFactor is None else float(switchFactor) * self._r0
        self._reference = 1.0
        if is_batch_context(reference):
            reference = evaluate_in_context(self, reference)
        self._reference = float(value_in_md_units(reference))
        self._registerCV(
            name,
            mmunit.dimensionless,
            group1=group1,
            group2=group2,
            nonbondedForce=nonbondedForce,
            reference=self._reference,
            stepFunction=stepFunction,
            thresholdDistance=thresholdDistance,
            cutoffFactor=cutoffFactor,
            switchFactor=switchFactor,
        )

    def usesPeriodicBoundaryConditions(self) -> bool:
        return self._pbc

    def _createNative(self, system: mm.System) -> _native.NativeCV:
        group1 = _native.as_i32(self._groups[0])
        group2 = _native.as_i32(self._groups[1])
        kind, p0, p1 = self._step
        desc = _native.PairSumDesc(
            kind=_native.CVP_PAIR_CONTACTS,
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
            step_kind=kind,
            step_p0=p0,
            step_p1=p1,
            r0=self._r0,
        )
        return _native.NativeCV.create(
            "cvp_pair_sum_create", desc, (group1, group2, self._exclusions)
        )


NumberOfContacts.registerTag("!cvpack.NumberOfContacts")
