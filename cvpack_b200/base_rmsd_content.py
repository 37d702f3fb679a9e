"""
.. class:: BaseRMSDContent -- secondary-structure RMSD content
   (``/root/reference/cvpack/base_rmsd_content.py``).
"""

from __future__ import annotations

import json
import os
import typing as t

import numpy as np

from . import _native, mm
from .collective_variable import CollectiveVariable
from .step_functions import parse_step_function
from .units import ScalarQuantity, value_in_md_units

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "ideal_structures.json")


class BaseRMSDContent(CollectiveVariable):
    """Abstract class for secondary-structure RMSD content of a sequence of `n` residues."""

    def __init__(  # pylint: disable=too-many-arguments,super-init-not-called
        self,
        residue_blocks: t.List[t.List[int]],
        ideal_positions: np.ndarray,
        residues: t.Sequence[t.Any],
        numAtoms: int,
        thresholdRMSD: ScalarQuantity,
        stepFunction: str = "(1+x^4)/(1+x^4+x^8)",
        normalize: bool = False,
    ) -> None:
        num_residue_blocks = self._num_residue_blocks = len(residue_blocks)
        if not 1 <= num_residue_blocks <= 1024:
            raise ValueError(
                f"{len(residues)} residues yield {num_residue_blocks} blocks, "
                "which is not between 1 and 1024"
            )
        residue_atoms = list(map(self._getAtomList, residues))
        self._block_atoms = [
            sum([residue_atoms[index] for index in block], []) for block in residue_blocks
        ]
        self._ideal = np.asarray(ideal_positions, dtype=np.float64)
        self._num_atoms = int(numAtoms)
        self._r0 = float(value_in_md_units(thresholdRMSD))
        self._step = parse_step_function(stepFunction)
        self._normalize = bool(normalize)

    @classmethod
    def _loadPositions(cls, key: str) -> np.ndarray:
        """Ideal coordinates (30 x 3, nm) of one of the tabulated six-residue fragments."""
        with open(_DATA, encoding="utf-8") as file:
            table = json.load(file)
        return 0.1 * np.asarray(table[key], dtype=np.float64)

    @staticmethod
    def _getAtomList(residue: t.Any) -> t.List[int]:
        """Indices of N, CA, CB, C, O of a residue (HA2 stands in for CB in glycine)."""
        residue_atoms = {atom.name: atom.index for atom in residue.atoms()}
        if residue.name == "GLY":
            residue_atoms["CB"] = residue_atoms["HA2"]
        atom_list = []
        for atom in ("N", "CA", "CB", "C", "O"):
            try:
                atom_list.append(residue_atoms[atom])
            except KeyError as error:
                raise ValueError(
                    f"Atom {atom} not found in residue {residue.name}{residue.id}"
                ) from error
        return atom_list

    def getNumResidueBlocks(self) -> int:
        """Get the number of residue blocks."""
        return self._num_residue_blocks

    def _createNative(self, system: mm.System) -> _native.NativeCV:
        nblocks = self._num_residue_blocks
        from .base_rmsd import rmsd_spec  # pylint: disable=import-outside-toplevel

        return rmsd_spec(
            self._num_atoms,
            self._block_atoms,
            [self._ideal] * nblocks,
            combine=_native.CVP_COMBINE_STEP_SUM,
            step=self._step,
            r0=self._r0,
            scale=1.0 / nblocks if self._normalize else 1.0,
        )
