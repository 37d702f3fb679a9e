"""
.. class:: CompositeRMSD (``/root/reference/cvpack/composite_rmsd.py``).
"""

from __future__ import annotations

import typing as t

import numpy as np

from . import _native, mm, mmunit
from .base_rmsd import BaseRMSD, rmsd_spec
from .units import MatrixQuantity, VectorQuantity


class CompositeRMSD(BaseRMSD):
    r"""
    The composite RMSD of :math:`m` disjoint atom groups (bodies) with :math:`n = \sum_j n_j`
    atoms in total (nm),

    .. math::
        d_{\rm crms}({\bf r}) = \sqrt{\frac{1}{n} \min_{\bf q} \sum_{j=1}^m \sum_{i \in {\bf g}_j}
        \left\| {\bf A}({\bf q}) \big({\bf r}_i - {\bf c}_j\big) -
        \big({\bf r}_i^{\rm ref} - {\bf c}_j^{\rm ref}\big) \right\|^2},

    each body being centred on its own centroid :math:`{\bf c}_j` in both structures while one
    rotation is shared by all -- so the measure is blind to the relative translation of the bodies
    but not to their relative orientation (``composite_rmsd.py:36-59`` in the reference).  The
    reference needs the ``openmm-cpp-forces`` plugin for this CV and raises ``ImportError``
    without it; here it is one more variant of the RMSD kernels and always available.

    Parameters
    ----------
    referencePositions
        A coordinate matrix or a mapping ``{atom index: coordinates}`` containing all atoms of
        ``groups``
    groups
        A sequence of disjoint atom groups
    numAtoms
        Total number of atoms of the system; needed only when ``referencePositions`` does not
        list all of them
    name
        The name of the collective variable

    Raises
    ------
    ValueError
        If ``groups`` is not a sequence of disjoint atom groups
    """

    def __init__(
        self,
        referencePositions: t.Union[MatrixQuantity, t.Dict[int, VectorQuantity]],
        groups: t.Iterable[t.Iterable[int]],
        numAtoms: t.Optional[int] = None,
        name: str = "composite_rmsd",
    ) -> None:
        num_atoms = numAtoms or len(referencePositions)
        groups = [[int(atom) for atom in group] for group in groups]
        all_atoms = sum(groups, [])
        if len(set(all_atoms)) != len(all_atoms) or any(len(group) == 0 for group in groups):
            raise ValueError("The groups must be a sequence of disjoint, non-empty atom groups.")
        defined_coords = self._getDefinedCoords(referencePositions, all_atoms)
        self._groups = groups
        self._num_atoms = int(num_atoms)
        self._reference = np.array([defined_coords[atom] for atom in all_atoms], dtype=np.float64)
        self._registerCV(
            name,
            mmunit.nanometers,
            referencePositions=defined_coords,
            groups=groups,
            numAtoms=num_atoms,
        )

    def _createNative(self, system: mm.System) -> _native.NativeCV:
        atoms = sum(self._groups, [])
        bodies = sum(([j] * len(group) for j, group in enumerate(self._groups)), [])
        return rmsd_spec(self._num_atoms, [atoms], [self._reference], bodies=[bodies])


CompositeRMSD.registerTag("!cvpack.CompositeRMSD")
