"""
Abstract base of all collective variables.

Mirrors ``/root/reference/cvpack/collective_variable.py``: the same public methods
(``getName``, ``getUnit``, ``getMassUnit``, ``getPeriodicBounds``, ``addToSystem``, ``getValue``,
``getEffectiveMass``; ``:122-309``), the same protected registration helpers (``_registerCV``,
``_registerPeriodicBounds``, ``_setUnusedForceGroup``; ``:44-120``) and the same YAML state
(constructor keyword arguments; ``:32-42``).  The difference is what sits underneath: where the
reference is an ``openmm.Force`` evaluated one frame at a time by ``Context.getState``, a CV here
compiles into a *kernel spec* of the CUDA library (``include/cvpack_b200.h``) and is evaluated for a
whole batch of frames by :class:`cvpack_b200.batch.BatchContext`.
"""

from __future__ import annotations

import typing as t

import yaml

from . import mm, mmunit
from .serialization import Serializable
from .units import Quantity, ScalarQuantity, Unit
from .utils import find_cv_in_context, preprocess_args


class CollectiveVariable(mm.Force, Serializable):
    """An abstract class with common attributes and methods for all CVs."""

    _unit: Unit = Unit("dimensionless")
    _mass_unit: Unit = Unit("dalton")
    _args: t.Dict[str, t.Any] = {}
    _periodic_bounds: t.Optional[Quantity] = None

    def __getstate__(self) -> t.Dict[str, t.Any]:
        return self._args

    def __setstate__(self, keywords: t.Dict[str, t.Any]) -> None:
        self.__init__(**keywords)

    def __copy__(self) -> "CollectiveVariable":
        return yaml.safe_load(yaml.safe_dump(self))

    def __deepcopy__(self, _: t.Any) -> "CollectiveVariable":
        return yaml.safe_load(yaml.safe_dump(self))

    @preprocess_args
    def _registerCV(self, name: str, cvUnit: Unit, **kwargs: t.Any) -> None:
        """
        Register the newly created instance: name, unit, effective-mass unit
        (Da (nm/unit)^2, ``collective_variable.py:69`` in the reference) and the keyword
        arguments needed to rebuild it.  Must be called from ``Subclass.__init__``.
        """
        mm.Force.__init__(self)
        self.setName(name)
        self._unit = cvUnit
        self._mass_unit = Unit(mmunit.dalton * (mmunit.nanometers / self._unit) ** 2)
        self._args = {"name": name}
        self._args.update(kwargs)

    def _registerPeriodicBounds(self, lower: ScalarQuantity, upper: ScalarQuantity) -> None:
        """Register the periodic bounds of this collective variable (if it is periodic)."""
        if mmunit.is_quantity(lower):
            lower = lower.value_in_unit(self.getUnit())
        if mmunit.is_quantity(upper):
            upper = upper.value_in_unit(self.getUnit())
        self._periodic_bounds = Quantity((lower, upper), self.getUnit())

    def _setUnusedForceGroup(self, system: mm.System) -> None:
        """
        Move this CV to the lowest force group (0..31) not used by any force of ``system``.

        Raises
        ------
        RuntimeError
            If all force groups are already in use
        """
        used_groups = {force.getForceGroup() for force in system.getForces()}
        new_group = next(filter(lambda i: i not in used_groups, range(32)), None)
        if new_group is None:
            raise RuntimeError("All force groups are already in use.")
        self.setForceGroup(new_group)

    # ---- what a subclass provides to the CUDA library ----------------------------------------
    def _createNative(self, native: t.Any, system: mm.System) -> t.Any:
        """Build the kernel spec of this CV for ``system`` (masses, atom count)."""
        raise NotImplementedError

    # ---- public API ---------------------------------------------------------------------------
    def getUnit(self) -> Unit:
        """Get the unit of measurement of this collective variable."""
        return self._unit

    def getMassUnit(self) -> Unit:
        """Get the unit of measurement of the effective mass of this collective variable."""
        return self._mass_unit

    def getPeriodicBounds(self) -> t.Union[Quantity, None]:
        """Get the periodic bounds of this collective variable (None if not periodic)."""
        return self._periodic_bounds

    def addToSystem(self, system: mm.System, setUnusedForceGroup: bool = True) -> None:
        """
        Add this collective variable to a :class:`~cvpack_b200.mm.System`.

        Parameters
        ----------
        system
            The system to which this collective variable should be added
        setUnusedForceGroup
            If True, the force group of this collective variable will be set to the
            first available force group in the system
        """
        if setUnusedForceGroup:
            self._setUnusedForceGroup(system)
        system.addForce(self)

    def getValue(self, context: t.Any, allowReinitialization: bool = False) -> Quantity:
        """
        Evaluate this collective variable at every frame of a batch context.

        Returns a :class:`Quantity` whose value is a tensor of shape ``[B]`` (on the device of the
        context) in the unit of this CV.  ``allowReinitialization`` is accepted for signature
        parity with the reference and has no effect: a batch context always evaluates a CV on
        its own.

        Raises
        ------
        RuntimeError
            ``"This force is not present in the given context."`` if the CV was not added to the
            system of ``context``.
        """
        del allowReinitialization
        find_cv_in_context(self, context)
        value, _ = context._evaluate(self, need_gradient=False)
        return Quantity(value, self.getUnit())

    def getGradient(self, context: t.Any) -> t.Any:
        """
        The gradient of this collective variable with respect to all atom positions at every
        frame: a tensor ``[B, N, 3]`` in (CV unit)/nm.  OpenMM would return its negative as
        "forces" (``utils.py:172-174`` in the reference).
        """
        find_cv_in_context(self, context)
        _, gradient = context._evaluate(self, need_gradient=True)
        return gradient

    def getValueAndGradient(self, context: t.Any) -> t.Tuple[Quantity, t.Any]:
        """Value (``[B]`` quantity) and gradient (``[B, N, 3]`` tensor) from one evaluation."""
        find_cv_in_context(self, context)
        value, gradient = context._evaluate(self, need_gradient=True)
        return Quantity(value, self.getUnit()), gradient

    def getEffectiveMass(self, context: t.Any, allowReinitialization: bool = False) -> Quantity:
        r"""
        Compute the effective mass of this collective variable at every frame of a batch context,

        .. math::

            m_\mathrm{eff}({\bf r}) = \left(
                \sum_{i=1}^N \frac{1}{m_i} \left\| \frac{dq}{d{\bf r}_i} \right\|^2
            \right)^{-1}

        with the sum restricted to atoms whose gradient is non-zero and ``inf`` when there is
        none (``utils.py:172-184`` in the reference).  Returns a ``[B]`` quantity in
        :meth:`getMassUnit`.
        """
        del allowReinitialization
        find_cv_in_context(self, context)
        return Quantity(context._effectiveMass(self), self._mass_unit)
