"""
.. class:: PathInRMSDSpace (``/root/reference/cvpack/path_in_rmsd_space.py``).
"""

from __future__ import annotations

import typing as t

import numpy as np

from . import _native, mm, mmunit
from .base_path_cv import BasePathCV
from .base_rmsd import rmsd_spec
from .path import Metric, progress
from .units import VectorQuantity, value_in_md_units


class PathInRMSDSpace(BasePathCV):
    r"""
    A metric of progress along, or deviation from, a path of :math:`n` milestone structures in
    RMSD space (Branduardi et al., 2007): the :class:`BasePathCV` formulas with
    :math:`x_i = d_{\rm rms}^2({\bf r}, {\bf r}^{\rm ref}_i)`.  Progress is dimensionless;
    deviation is in nm\ :sup:`2`.  Each milestone is a mapping ``{atom index: coordinates}`` and
    carries its own atom set and order (``path_in_rmsd_space.py:137-140`` in the reference).

    Parameters
    ----------
    metric
        ``cvpack_b200.path.progress`` or ``cvpack_b200.path.deviation``
    milestones
        The reference structures, each ``{atom index: [x, y, z]}``
    numAtoms
        The total number of atoms in the system
    sigma
        The width :math:`\sigma` of the Gaussian kernels (nm)
    name
        The name of the collective variable; defaults to ``path_progress_in_rmsd_space`` or
        ``path_deviation_in_rmsd_space``

    Raises
    ------
    ValueError
        If there are fewer than two milestones or the metric is invalid
    """

    def __init__(  # pylint: disable=too-many-arguments
        self,
        metric: Metric,
        milestones: t.Sequence[t.Dict[int, VectorQuantity]],
        numAtoms: int,
        sigma: t.Any,
        name: t.Optional[str] = None,
    ) -> None:
        name = self._generateName(metric, name, "rmsd")
        sigma = float(value_in_md_units(sigma))
        n = len(milestones)
        if n < 2:
            raise ValueError("At least two reference structures are required.")
        self._metric = metric
        self._sigma = sigma
        self._num_atoms = int(numAtoms)
        self._groups = [[int(atom) for atom in milestone.keys()] for milestone in milestones]
        self._references = [
            np.array(
                [[float(x) for x in value_in_md_units(coords)] for coords in milestone.values()],
                dtype=np.float64,
            ).reshape(-1, 3)
            for milestone in milestones
        ]
        self._registerCV(
            name,
            mmunit.dimensionless if metric == progress else mmunit.nanometers**2,
            metric=metric,
            milestones=[
                {int(k): [float(x) for x in value_in_md_units(v)] for k, v in m.items()}
                for m in milestones
            ],
            numAtoms=numAtoms,
            sigma=sigma,
        )

    def _createNative(self, system: mm.System) -> _native.NativeCV:
        combine = (
            _native.CVP_COMBINE_PATH_PROGRESS
            if self._metric == progress
            else _native.CVP_COMBINE_PATH_DEVIATION
        )
        return rmsd_spec(
            self._num_atoms, self._groups, self._references, combine=combine, sigma=self._sigma
        )


PathInRMSDSpace.registerTag("!cvpack.PathInRMSDSpace")
