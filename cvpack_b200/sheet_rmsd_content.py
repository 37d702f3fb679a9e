"""
.. class:: SheetRMSDContent (``/root/reference/cvpack/sheet_rmsd_content.py``).
"""

from __future__ import annotations

import typing as t

import numpy as np

from . import mmunit
from .base_rmsd_content import BaseRMSDContent
from .units import Quantity, ScalarQuantity

# pylint: disable=protected-access
PARABETA_POSITIONS = BaseRMSDContent._loadPositions("parallel_beta_sheet")
ANTIBETA_POSITIONS = BaseRMSDContent._loadPositions("antiparallel_beta_sheet")
# pylint: enable=protected-access


class SheetRMSDContent(BaseRMSDContent):
    r"""
    The beta-sheet RMSD content of :math:`n` residues (Pietrucci & Laio, 2009),

    .. math::
        \beta_{\rm rmsd}({\bf r}) = \sum_{i} \sum_{j \ge i + d_{\rm min}}
        S\!\left(\frac{d_{\rm rms}({\bf g}_{i,j}, {\bf g}_{\rm ref})}{r_0}\right),

    where :math:`{\bf g}_{i,j}` holds the N, CA, CB, C, O atoms of residues
    :math:`i, i+1, i+2, j, j+1, j+2` (30 atoms), :math:`d_{\rm min}` is 5 for antiparallel and 6
    for parallel sheets, and :math:`{\bf g}_{\rm ref}` is the matching ideal fragment.  With
    ``blockSizes`` the residues are split in consecutive blocks and :math:`i`, :math:`j` run over
    neighbouring blocks only (``sheet_rmsd_content.py:190-204`` in the reference).

    Parameters
    ----------
    residues
        The residues to be considered
    numAtoms
        The total number of atoms in the system
    parallel
        Whether to look for parallel rather than antiparallel sheets
    blockSizes
        Sizes of consecutive residue blocks, or None
    thresholdRMSD
        :math:`r_0`
    stepFunction
        One of the whitelisted forms of :mod:`cvpack_b200.step_functions`
    normalize
        Whether to divide by the number of residue sextets
    name
        The name of the collective variable

    Raises
    ------
    ValueError
        If the block sizes do not add up to the number of residues, or the number of sextets is
        not between 1 and 1024
    """

    def __init__(  # pylint: disable=too-many-arguments
        self,
        residues: t.Sequence[t.Any],
        numAtoms: int,
        parallel: bool = False,
        blockSizes: t.Optional[t.Sequence[int]] = None,
        thresholdRMSD: ScalarQuantity = Quantity(0.08 * mmunit.nanometers),
        stepFunction: str = "(1+x^4)/(1+x^4+x^8)",
        normalize: bool = False,
        name: str = "sheet_rmsd_content",
    ) -> None:
        num_residues = len(residues)
        if blockSizes is None:
            min_distance = 6 if parallel else 5
            residue_groups = [
                [i, i + 1, i + 2, j, j + 1, j + 2]
                for i in range(num_residues - 2 - min_distance)
                for j in range(i + min_distance, num_residues - 2)
            ]
        elif sum(blockSizes) == num_residues:
            bounds = [0] + [int(b) for b in np.cumsum(blockSizes)]
            residue_groups = [
                [i, i + 1, i + 2, j, j + 1, j + 2]
                for k in range(len(blockSizes) - 1)
                for i in range(bounds[k], bounds[k + 1] - 2)
                for j in range(bounds[k + 1], bounds[k + 2] - 2)
            ]
        else:
            raise ValueError(
                f"The sum of block sizes ({sum(blockSizes)}) and the "
                f"number of residues ({num_residues}) must be equal."
            )
        super().__init__(
            residue_groups,
            PARABETA_POSITIONS if parallel else ANTIBETA_POSITIONS,
            residues,
            numAtoms,
            thresholdRMSD,
            stepFunction,
            normalize,
        )
        self._registerCV(
            name,
            mmunit.dimensionless,
            residues=residues,
            numAtoms=numAtoms,
            parallel=parallel,
            blockSizes=blockSizes,
            thresholdRMSD=thresholdRMSD,
            stepFunction=stepFunction,
            normalize=normalize,
        )


SheetRMSDContent.registerTag("!cvpack.SheetRMSDContent")
