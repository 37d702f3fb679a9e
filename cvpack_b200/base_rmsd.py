"""
.. class:: BaseRMSD -- reference-coordinate handling shared by the RMSD CVs
   (``/root/reference/cvpack/base_rmsd.py``).
"""

from __future__ import annotations

import typing as t

import numpy as np

from . import _native
from .collective_variable import CollectiveVariable
from .units import MatrixQuantity, VectorQuantity, value_in_md_units


class BaseRMSD(CollectiveVariable):
    """A base class for root-mean-square deviation (RMSD) collective variables."""

    def _getDefinedCoords(
        self,
        referencePositions: t.Union[MatrixQuantity, t.Dict[int, VectorQuantity]],
        group: t.List[int],
    ) -> t.Dict[int, t.List[float]]:
        """``{atom: [x, y, z]}`` (nm) for the atoms of ``group`` only."""
        if isinstance(referencePositions, t.Dict):
            positions = {
                atom: value_in_md_units(coords) for atom, coords in referencePositions.items()
            }
        else:
            positions = value_in_md_units(referencePositions)
        return {atom: [float(x) for x in positions[atom]] for atom in group}

    def _getAllCoords(
        self, definedCoords: t.Dict[int, t.List[float]], numAtoms: int
    ) -> np.ndarray:
        """``[numAtoms, 3]`` matrix with the defined rows filled in and zeros elsewhere."""
        all_coords = np.zeros((numAtoms, 3))
        for atom, coords in definedCoords.items():
            all_coords[atom] = coords
        return all_coords


def rmsd_spec(  # pylint: disable=too-many-arguments,too-many-locals
    num_atoms: int,
    groups: t.Sequence[t.Sequence[int]],
    references: t.Sequence[np.ndarray],
    bodies: t.Optional[t.Sequence[t.Sequence[int]]] = None,
    combine: int = _native.CVP_COMBINE_IDENTITY,
    step: t.Tuple[int, float, float] = (_native.CVP_STEP_RATIONAL, 6.0, 0.0),
    r0: float = 1.0,
    scale: float = 1.0,
    sigma: float = 0.0,
) -> _native.NativeCV:
    """
    Kernel spec of the RMSD family (``cvp_rmsd_create``): one superposition problem per group,
    ``references[g]`` rows pairing with ``groups[g]``; ``bodies[g]`` gives the body of each entry.
    """
    offsets = np.zeros(len(groups) + 1, dtype=np.int64)
    offsets[1:] = np.cumsum([len(group) for group in groups])
    atoms = _native.as_i32(np.concatenate([np.asarray(group, dtype=np.int64) for group in groups]))
    reference = _native.as_f64(np.concatenate([np.asarray(ref).reshape(-1, 3) for ref in references]))
    if reference.shape[0] != atoms.shape[0]:
        raise ValueError("reference coordinates do not match the atom groups")
    body_ids = None
    if bodies is not None:
        body_ids = _native.as_i32(np.concatenate([np.asarray(b, dtype=np.int64) for b in bodies]))
    offsets32 = _native.as_i32(offsets)
    desc = _native.RmsdDesc(
        num_atoms=int(num_atoms),
        num_groups=len(groups),
        group_offsets=_native.ptr_i32(offsets32),
        atoms=_native.ptr_i32(atoms),
        body_ids=_native.ptr_i32(body_ids),
        reference=_native.ptr_f64(reference),
        combine=combine,
        step_kind=step[0],
        step_p0=step[1],
        step_p1=step[2],
        r0=float(r0),
        scale=float(scale),
        sigma=float(sigma),
    )
    return _native.NativeCV.create(
        "cvp_rmsd_create", desc, (offsets32, atoms, body_ids, reference)
    )
