"""
.. class:: MetaCollectiveVariable -- a function of other collective variables
   (``/root/reference/cvpack/meta_collective_variable.py``).
"""

from __future__ import annotations

import typing as t
from copy import copy

import torch

from . import _native, mmunit
from .collective_variable import CollectiveVariable
from .composite import evaluate_composite
from .expression import compile_expression
from .units import Quantity, ScalarQuantity, VectorQuantity, in_md_units
from .utils import find_cv_in_context


class MetaCollectiveVariable(CollectiveVariable):
    r"""
    A collective variable that is a function of other collective variables:

    .. math::

        f_{\rm meta}({\bf r}) = F(c_1({\bf r}), c_2({\bf r}), \ldots, c_n({\bf r}))

    where :math:`F` is a user-defined function and :math:`c_i` is the value of the
    :math:`i`-th inner variable.  The function is an algebraic expression of the inner variables'
    *names* and of named parameters passed as keyword arguments, in OpenMM's expression syntax
    (``meta_collective_variable.py:30-66``).

    The reference builds an ``openmm.CustomCVForce`` (``:127-133``); here the expression is compiled
    once into a postfix program (``cvpack_b200/expression.py``), the inner variables run their own
    batched kernels, ``cvp_expr_eval`` produces the value and :math:`\partial F/\partial c_j` for
    every frame, and ``cvp_axpy_frames`` assembles the gradient
    :math:`\sum_j \partial F/\partial c_j \nabla c_j`.

    Parameters
    ----------
    function
        The function to be evaluated: an expression of the names of the inner variables and of
        the parameters; ``"main; a = ...; b = ..."`` definitions are allowed
    variables
        The inner collective variables (copied)
    unit
        The unit of measurement of the result (keeping it consistent with the function is the
        caller's responsibility, as in the reference)
    periodicBounds
        The periodic bounds of the result, or None
    name
        The name of the collective variable
    **parameters
        Named parameters of the function (quantities are converted to MD units)

    Example
    -------
    >>> import cvpack_b200 as cvpack
    >>> from cvpack_b200 import mmunit as unit
    >>> phi = cvpack.Torsion(6, 8, 14, 16, name="phi")
    >>> driving_force = cvpack.MetaCollectiveVariable(
    ...     "0.5*kappa*min(delta,2*pi-delta)^2; delta=abs(phi-phi0)",
    ...     [phi], unit.kilojoules_per_mole,
    ...     kappa=1e3 * unit.kilojoules_per_mole / unit.radian**2,
    ...     phi0=120 * unit.degrees, pi=3.141592653589793,
    ... )
    >>> driving_force.getParameterDefaultValues()["phi0"]
    2.0943951023931953 rad
    """

    def __init__(  # pylint: disable=too-many-arguments
        self,
        function: str,
        variables: t.Iterable[CollectiveVariable],
        unit: mmunit.Unit,
        periodicBounds: t.Optional[VectorQuantity] = None,
        name: str = "meta_collective_variable",
        **parameters: ScalarQuantity,
    ) -> None:
        variables = list(variables)
        self._function = function
        self._cvs = tuple(map(copy, variables))
        self._parameters = {k: in_md_units(v) for k, v in parameters.items()}
        names = [cv.getName() for cv in self._cvs]
        if len(set(names)) != len(names):
            raise ValueError("The inner collective variables must have distinct names.")
        # compiled here so that a malformed function fails at construction, as in the reference
        self._program = compile_expression(function, names, list(self._parameters))
        self._registerCV(
            name,
            unit,
            function=function,
            variables=variables,
            unit=unit,
            periodicBounds=periodicBounds,
            **self._parameters,
        )
        if periodicBounds is not None:
            self._registerPeriodicBounds(*periodicBounds)

    def usesPeriodicBoundaryConditions(self) -> bool:
        return any(cv.usesPeriodicBoundaryConditions() for cv in self._cvs)

    # ---- evaluation ---------------------------------------------------------------------------
    def _parameterValues(self, context: t.Any) -> t.Dict[str, float]:
        values = {}
        for key, default in self._parameters.items():
            values[key] = float(context.getParameter(key, default / default.unit))
        return values

    def _evaluateComposite(
        self,
        context: t.Any,
        need_gradient: bool,
        positions: t.Optional[torch.Tensor],
        out: t.Optional[t.Tuple[torch.Tensor, t.Optional[torch.Tensor]]],
    ) -> t.Tuple[torch.Tensor, t.Optional[torch.Tensor]]:
        value, grad, _, _ = self._evaluateAll(context, need_gradient, positions, out)
        return value, grad

    def _evaluateAll(self, context, need_gradient, positions=None, out=None):
        program = self._program.to_struct(self._parameterValues(context))
        cvp_dtype = context._cvp_dtype  # pylint: disable=protected-access

        def combine(values, out_value, out_partials, stream):
            _native.expr_eval(
                program, cvp_dtype, values.data_ptr() if values.numel() else None, int(values.shape[1]),
                out_value.data_ptr(), None if out_partials is None else out_partials.data_ptr(), stream,
            )

        return evaluate_composite(
            context, self._cvs, combine, len(self._cvs) + len(self._parameters), need_gradient,
            positions, out,
        )

    # ---- the reference's extra public methods (meta_collective_variable.py:143-261) ------------
    def getInnerVariables(self) -> t.Tuple[CollectiveVariable, ...]:
        """The (copied) collective variables on which this one depends."""
        return self._cvs

    def getInnerValues(self, context: t.Any) -> t.Dict[str, Quantity]:
        """Values ``[B]`` of the inner variables, keyed by name."""
        find_cv_in_context(self, context)
        _, _, values, _ = self._evaluateAll(context, False)
        return {cv.getName(): Quantity(values[j], cv.getUnit()) for j, cv in enumerate(self._cvs)}

    def getInnerEffectiveMasses(self, context: t.Any) -> t.Dict[str, Quantity]:
        """Effective masses ``[B]`` of the inner variables, keyed by name."""
        find_cv_in_context(self, context)
        result = {}
        for cv in self._cvs:
            _, grad = context._run(cv, True)  # pylint: disable=protected-access
            mass = context._massFromGradient(grad)  # pylint: disable=protected-access
            result[cv.getName()] = Quantity(mass, cv.getMassUnit())
        return result

    def getParameterDefaultValues(self) -> t.Dict[str, Quantity]:
        """Default values of the named parameters."""
        return self._parameters.copy()

    def getParameterValues(self, context: t.Any) -> t.Dict[str, Quantity]:
        """Values of the named parameters in a context (``context.setParameter`` overrides defaults)."""
        values = self._parameterValues(context)
        return {key: Quantity(values[key], default.unit) for key, default in self._parameters.items()}

    def getParameterDerivatives(
        self, context: t.Any, allowReinitialization: bool = False
    ) -> t.Dict[str, Quantity]:
        """Derivatives ``[B]`` of this variable with respect to the named parameters."""
        del allowReinitialization
        find_cv_in_context(self, context)
        _, _, _, partials = self._evaluateAll(context, True)
        offset = len(self._cvs)
        return {
            key: Quantity(partials[offset + k], self._unit / default.unit)
            for k, (key, default) in enumerate(self._parameters.items())
        }


MetaCollectiveVariable.registerTag("!cvpack.MetaCollectiveVariable")
