"""
.. class:: BasePathCV -- shared part of the path collective variables
   (``/root/reference/cvpack/base_path_cv.py``).
"""

from __future__ import annotations

import typing as t

from .collective_variable import CollectiveVariable
from .path import Metric, deviation, progress


class BasePathCV(CollectiveVariable):
    r"""
    A base class for path-related collective variables.  With squared distances :math:`x_i` to
    :math:`n` milestones and :math:`\lambda = 1/(2\sigma^2)`, the weights are
    :math:`w_i = e^{\lambda(x_{\rm min} - x_i)}` and

    .. math::
        s = \frac{\sum_{i=0}^{n-1} i\, w_i}{(n-1) \sum_i w_i}, \qquad
        z = x_{\rm min} - \frac{1}{\lambda} \ln \sum_i w_i

    (``base_path_cv.py:44-59`` in the reference; the shift by :math:`x_{\rm min}` only stabilises
    the exponentials).
    """

    def _generateName(self, metric: Metric, name: t.Optional[str], kind: str) -> str:
        if metric not in (progress, deviation):
            raise ValueError(
                "Invalid metric. Use 'cvpack.path.progress' or 'cvpack.path.deviation'."
            )
        if name is None:
            return f"path_{metric.name}_in_{kind}_space"
        return name
