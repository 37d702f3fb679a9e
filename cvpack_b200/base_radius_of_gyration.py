"""
.. class:: BaseRadiusOfGyration -- shared part of the two radius-of-gyration CVs
   (``/root/reference/cvpack/base_radius_of_gyration.py``).
"""

from __future__ import annotations

import typing as t

import numpy as np

from . import _native, mm
from .collective_variable import CollectiveVariable
from .utils import as_index_list


class BaseRadiusOfGyration(CollectiveVariable):
    """Abstract class for the radius of gyration of a group of `n` atoms."""

    _squared = False

    def _setGroup(self, group: t.Iterable[int], pbc: bool, weighByMass: bool) -> t.List[int]:
        self._group = as_index_list(group)
        if len(set(self._group)) != len(self._group):
            raise ValueError("The group contains repeated atoms.")
        self._pbc = bool(pbc)
        self._weigh_by_mass = bool(weighByMass)
        return self._group

    def usesPeriodicBoundaryConditions(self) -> bool:
        return self._pbc

    def _createNative(self, system: mm.System) -> _native.NativeCV:
        group = _native.as_i32(self._group)
        n = len(group)
        if self._weigh_by_mass:
            # centre of mass; the sum of squared distances itself stays unweighted
            # (tests/test_cvpack.py:185-187 in the reference)
            masses = system.getMasses()[group]
            weights = masses / masses.sum()
        else:
            weights = np.full(n, 1.0 / n)
        weights = _native.as_f64(weights)
        desc = _native.GyrationDesc(
            num_atoms=system.getNumParticles(),
            n=n,
            group=_native.ptr_i32(group),
            weights=_native.ptr_f64(weights),
            periodic=int(self._pbc),
            squared=int(self._squared),
        )
        return _native.NativeCV.create("cvp_gyration_create", desc, (group, weights))
