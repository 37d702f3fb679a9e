"""
.. class:: PathInCVSpace -- progress or deviation with respect to a path in CV space
   (``/root/reference/cvpack/path_in_cv_space.py``).
"""

from __future__ import annotations

import typing as t
from copy import deepcopy

import numpy as np
import torch

from . import _native, mmunit
from .base_path_cv import BasePathCV
from .composite import evaluate_composite
from .collective_variable import CollectiveVariable
from .path import Metric, progress
from .units import value_in_md_units
from .utils import convert_to_matrix


class PathInCVSpace(BasePathCV):
    r"""
    A metric of the system's progress (:math:`s`) or deviation (:math:`z`) with respect to a path
    defined by :math:`n` milestones positioned in a collective-variable space:

    .. math::
        s = \frac{\sum_{i=0}^{n-1} i\, w_i}{(n-1)\sum_i w_i}
        \quad \text{or} \quad
        z = -2\sigma^2 \ln \sum_i w_i, \qquad
        w_i = \exp\Big(-\frac{\|{\bf c}({\bf r}) - \hat{\bf c}_i\|^2}{2\sigma^2}\Big),

    with :math:`\|{\bf x}\|^2 = {\bf x}^T {\bf D}^{-2} {\bf x}`, :math:`{\bf D}` the diagonal matrix
    of characteristic scales; for a periodic variable the difference is
    :math:`\min(|\delta|, P - |\delta|)` (``path_in_cv_space.py:141-152``).  Dimensionless.

    The reference nests the variables in an ``openmm.CustomCVForce``; here the inner variables are
    evaluated by their own CUDA kernels for the whole batch, ``cvp_path_cv_combine`` applies the
    chain rule per frame and ``cvp_axpy_frames`` accumulates the gradient.

    Parameters
    ----------
    metric
        ``cvpack.path.progress`` or ``cvpack.path.deviation``
    variables
        The collective variables that define the space
    milestones
        One row per milestone, one column per variable
    sigma
        The width of the Gaussian kernels
    scales
        The characteristic scales of the variables (default: 1 in MD units)
    name
        Default ``path_progress_in_cv_space`` / ``path_deviation_in_cv_space``

    Raises
    ------
    ValueError
        If the milestones matrix has fewer than 2 rows or a number of columns different from the
        number of variables, or if the metric is invalid
    """

    def __init__(  # pylint: disable=too-many-arguments
        self,
        metric: Metric,
        variables: t.Iterable[CollectiveVariable],
        milestones: t.Any,
        sigma: float,
        scales: t.Optional[t.Iterable[t.Any]] = None,
        name: t.Optional[str] = None,
    ) -> None:
        name = self._generateName(metric, name, "cv")
        variables = list(variables)
        cv_scales = [1.0] * len(variables) if scales is None else list(scales)
        milestones, n, numvars = convert_to_matrix(milestones)
        if numvars != len(variables):
            raise ValueError("Wrong number of columns in the milestones matrix.")
        if n < 2:
            raise ValueError("At least two rows are required in the milestones matrix.")
        if len(cv_scales) != len(variables):
            raise ValueError("One scale per collective variable is required.")
        periods = []
        for variable in variables:
            bounds = variable.getPeriodicBounds()
            if bounds is None:
                periods.append(0.0)
            else:
                lower, upper = bounds.value
                periods.append(float(value_in_md_units((upper - lower) * bounds.unit)))
        self._metric = metric
        self._sigma = float(value_in_md_units(sigma))
        self._variables = [deepcopy(variable) for variable in variables]
        self._milestones = np.array(
            [[float(value_in_md_units(v)) for v in row] for row in milestones], dtype=np.float64
        )
        self._periods = periods
        self._scales = [float(value_in_md_units(scale)) for scale in cv_scales]
        self._registerCV(
            name,
            mmunit.dimensionless,
            metric=metric,
            variables=variables,
            milestones=self._milestones.tolist(),
            sigma=sigma,
            scales=scales,
        )

    def usesPeriodicBoundaryConditions(self) -> bool:
        return any(variable.usesPeriodicBoundaryConditions() for variable in self._variables)

    def getInnerVariables(self) -> t.List[CollectiveVariable]:
        """The (copied) collective variables that span the space."""
        return list(self._variables)

    # Called by BatchContext._run instead of a single kernel spec.
    def _evaluateComposite(  # pylint: disable=too-many-locals
        self,
        context: t.Any,
        need_gradient: bool,
        positions: t.Optional[torch.Tensor],
        out: t.Optional[t.Tuple[torch.Tensor, t.Optional[torch.Tensor]]],
    ) -> t.Tuple[torch.Tensor, t.Optional[torch.Tensor]]:
        cvp_dtype = context._cvp_dtype  # pylint: disable=protected-access
        metric = (
            _native.CVP_COMBINE_PATH_PROGRESS
            if self._metric == progress
            else _native.CVP_COMBINE_PATH_DEVIATION
        )

        def combine(values, out_value, out_partials, stream):
            _native.path_cv_combine(
                self._milestones, self._periods, self._scales, self._sigma, metric, cvp_dtype,
                values.data_ptr(), int(values.shape[1]), out_value.data_ptr(),
                None if out_partials is None else out_partials.data_ptr(), stream,
            )

        # every inner variable is evaluated once, value and gradient together (composite.py)
        value, grad, _, _ = evaluate_composite(
            context, self._variables, combine, len(self._variables), need_gradient, positions, out
        )
        return value, grad

PathInCVSpace.registerTag("!cvpack.PathInCVSpace")
