"""
.. class:: RMSD (``/root/reference/cvpack/rmsd.py``).
"""

from __future__ import annotations

import typing as t

import numpy as np

from . import _native, mm, mmunit
from .base_rmsd import BaseRMSD, rmsd_spec
from .units import MatrixQuantity, VectorQuantity


class RMSD(BaseRMSD):
    r"""
    The root-mean-square deviation of a group of :math:`n` atoms from a reference structure after
    optimal superposition (nm),

    .. math::
        d_{\rm rms}({\bf r}) = \sqrt{\frac{1}{n} \min_{\bf q} \sum_{i \in {\bf g}}
        \left\| {\bf A}({\bf q}) \hat{\bf r}_i - \hat{\bf r}_i^{\rm ref} \right\|^2},

    where hats denote coordinates relative to the (unweighted) centroid of the group and
    :math:`{\bf A}({\bf q})` is the rotation of unit quaternion :math:`{\bf q}`, found per frame by
    a 4x4 eigen-solve (Horn / Coutsias).  Periodic boundary conditions are not applied: make sure
    the group is whole (``rmsd.py:40-46`` in the reference).

    Parameters
    ----------
    referencePositions
        A coordinate matrix covering all atoms up to the highest index of ``group``, or a mapping
        ``{atom index: coordinates}`` containing at least the atoms of ``group``
    group
        Indices of the atoms in the group
    numAtoms
        Total number of atoms of the system; needed only when ``referencePositions`` does not
        list all of them
    name
        The name of the collective variable
    """

    def __init__(
        self,
        referencePositions: t.Union[MatrixQuantity, t.Dict[int, VectorQuantity]],
        group: t.Iterable[int],
        numAtoms: t.Optional[int] = None,
        name: str = "rmsd",
    ) -> None:
        group = list(map(int, group))
        num_atoms = numAtoms or len(referencePositions)
        defined_coords = self._getDefinedCoords(referencePositions, group)
        self._group = group
        self._num_atoms = int(num_atoms)
        self._reference = np.array([defined_coords[atom] for atom in group], dtype=np.float64)
        self._registerCV(
            name,
            mmunit.nanometers,
            referencePositions=defined_coords,
            group=group,
            numAtoms=num_atoms,
        )

    def getNullBondForce(self) -> t.List[t.Tuple[int, int]]:
        """
        The chain of zero-strength bonds ``(group[k], group[k+1])`` that the reference returns as a
        ``HarmonicBondForce`` so that OpenMM keeps the group in one periodic image
        (``rmsd.py:104-147``).  Without OpenMM the pairs themselves are returned.
        """
        return list(zip(self._group[:-1], self._group[1:]))

    def _createNative(self, system: mm.System) -> _native.NativeCV:
        return rmsd_spec(self._num_atoms, [self._group], [self._reference])


RMSD.registerTag("!cvpack.RMSD")
