"""
cvpack_b200 -- a B200-native batched evaluator for the collective variables of cvpack.

Same class names, constructor arguments, units and YAML serialization as
``/root/reference/cvpack/__init__.py`` for the data-parallel core (value + per-atom gradient +
effective mass); evaluation runs hand-written sm_100a CUDA kernels over a whole batch of frames
(:class:`BatchContext`) instead of one ``openmm.Context`` call per frame.
"""

from . import app, bias, mm, mmunit, path, serialization, trajectory, units  # noqa: F401
from ._version import __version__  # noqa: F401
from .angle import Angle  # noqa: F401
from .attraction_strength import AttractionStrength  # noqa: F401
from .batch import BatchContext, shard_frames  # noqa: F401
from .collective_variable import CollectiveVariable  # noqa: F401
from .composite_rmsd import CompositeRMSD  # noqa: F401
from .distance import Distance  # noqa: F401
from .helix_angle_content import HelixAngleContent  # noqa: F401
from .helix_hbond_content import HelixHBondContent  # noqa: F401
from .helix_rmsd_content import HelixRMSDContent  # noqa: F401
from .helix_torsion_content import HelixTorsionContent  # noqa: F401
from .meta_collective_variable import MetaCollectiveVariable  # noqa: F401
from .mm import NonbondedForce, System, Vec3  # noqa: F401
from .number_of_contacts import NumberOfContacts  # noqa: F401
from .path_in_cv_space import PathInCVSpace  # noqa: F401
from .path_in_rmsd_space import PathInRMSDSpace  # noqa: F401
from .radius_of_gyration import RadiusOfGyration  # noqa: F401
from .radius_of_gyration_sq import RadiusOfGyrationSq  # noqa: F401
from .residue_coordination import ResidueCoordination  # noqa: F401
from .rmsd import RMSD  # noqa: F401
from .sheet_rmsd_content import SheetRMSDContent  # noqa: F401
from .shortest_distance import ShortestDistance  # noqa: F401
from .torsion import Torsion  # noqa: F401
from .torsion_similarity import TorsionSimilarity  # noqa: F401
from .trajectory import DCDFile, NumpyTrajectory, TrajectoryEvaluator, XTCFile  # noqa: F401

__all__ = [
    "Angle",
    "AttractionStrength",
    "BatchContext",
    "CollectiveVariable",
    "CompositeRMSD",
    "Distance",
    "HelixAngleContent",
    "HelixHBondContent",
    "HelixRMSDContent",
    "HelixTorsionContent",
    "MetaCollectiveVariable",
    "NonbondedForce",
    "NumberOfContacts",
    "PathInCVSpace",
    "PathInRMSDSpace",
    "RadiusOfGyration",
    "RadiusOfGyrationSq",
    "ResidueCoordination",
    "RMSD",
    "SheetRMSDContent",
    "ShortestDistance",
    "System",
    "Torsion",
    "TorsionSimilarity",
    "Vec3",
    "shard_frames",
]
