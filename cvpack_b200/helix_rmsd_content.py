"""
.. class:: HelixRMSDContent (``/root/reference/cvpack/helix_rmsd_content.py``).
"""

from __future__ import annotations

import typing as t

from . import mmunit
from .base_rmsd_content import BaseRMSDContent
from .units import Quantity, ScalarQuantity

ALPHA_POSITIONS = BaseRMSDContent._loadPositions("alpha_helix")  # pylint: disable=protected-access


class HelixRMSDContent(BaseRMSDContent):
    r"""
    The alpha-helix RMSD content of a sequence of :math:`n` residues (Pietrucci & Laio, 2009),

    .. math::
        \alpha_{\rm rmsd}({\bf r}) = \sum_{i=1}^{n-5}
        S\!\left(\frac{d_{\rm rms}({\bf g}_i, {\bf g}_{\rm ref})}{r_0}\right),
        \qquad S(x) = \frac{1 + x^4}{1 + x^4 + x^8} \ \text{by default},

    where :math:`{\bf g}_i` holds the N, CA, CB, C and O atoms of residues :math:`i \dots i+5`
    (30 atoms; HA2 stands in for CB in glycine) and :math:`{\bf g}_{\rm ref}` is an ideal alpha
    helix.  With ``normalize`` the sum is divided by the number of windows.  The residues must be
    a contiguous stretch of one chain in N- to C-terminus order.  Where the reference nests up to
    32 ``RMSDForce`` objects per inner ``CustomCVForce`` (``base_rmsd_content.py:52-79``), one
    kernel launch here solves every window of every frame.

    Parameters
    ----------
    residues
        The residues of the sequence (objects with ``name``, ``id`` and ``atoms()``)
    numAtoms
        The total number of atoms in the system
    thresholdRMSD
        :math:`r_0`
    stepFunction
        One of the whitelisted forms of :mod:`cvpack_b200.step_functions`
    normalize
        Whether to divide by the number of windows
    name
        The name of the collective variable

    Raises
    ------
    ValueError
        If the number of windows is not between 1 and 1024, or a backbone atom is missing
    """

    def __init__(  # pylint: disable=too-many-arguments
        self,
        residues: t.Sequence[t.Any],
        numAtoms: int,
        thresholdRMSD: ScalarQuantity = Quantity(0.08 * mmunit.nanometers),
        stepFunction: str = "(1+x^4)/(1+x^4+x^8)",
        normalize: bool = False,
        name: str = "helix_rmsd_content",
    ) -> None:
        residue_blocks = [list(range(index, index + 6)) for index in range(len(residues) - 5)]
        super().__init__(
            residue_blocks,
            ALPHA_POSITIONS,
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
            thresholdRMSD=thresholdRMSD,
            stepFunction=stepFunction,
            normalize=normalize,
        )


HelixRMSDContent.registerTag("!cvpack.HelixRMSDContent")
