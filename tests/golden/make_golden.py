"""
Regenerates tests/golden/known_answers.json.

The reference cannot be imported here (``import cvpack`` needs OpenMM, which is not installed and
cannot be: no network), so the golden vectors are the known-answer values the reference's own
docstrings and tests state without needing openmmtools coordinates; each entry cites its source.
Run: ``python tests/golden/make_golden.py``.
"""

import json
import os

import numpy as np

GOLDEN = {
    "distance": {
        "source": "cvpack/distance.py:37-50 (doctest: 1.7320... nm)",
        "positions": [[0, 0, 0], [1, 1, 1]],
        "atoms": [0, 1],
        "value": float(np.sqrt(3.0)),
    },
    "angle": {
        "source": "cvpack/angle.py:46-60 (doctest: 1.570796... rad)",
        "positions": [[0, 0, 0], [0, 0, 1], [0, 1, 1]],
        "atoms": [0, 1, 2],
        "value": float(np.pi / 2),
    },
    "torsion": {
        "source": "cvpack/torsion.py:53-67 (doctest: 1.5707... rad)",
        "positions": [[0, 0, 0], [1, 0, 0], [1, 1, 0], [1, 1, 1]],
        "atoms": [0, 1, 2, 3],
        "value": float(np.pi / 2),
    },
    "composite_rmsd": {
        "source": "cvpack/composite_rmsd.py:116-121 (doctest: 0.0 nm, and 0.0 nm after moving one "
        "body by 1 nm)",
        "value": 0.0,
    },
    "serialization_radius_of_gyration": {
        "source": "cvpack/serialization/serialization.py:138-154 (doctest: exact YAML text)",
        "text": "!cvpack.RadiusOfGyration\ngroup:\n- 0\n- 1\n- 2\nname: radius_of_gyration\n"
        "pbc: false\nweighByMass: false\n",
    },
    "mass_units": {
        "source": "cvpack/collective_variable.py:299-304 (doctest effective-mass units)",
        "rad": "nm**2 Da/(rad**2)",
        "nm": "Da",
    },
    "error_not_in_context": {
        "source": "cvpack/tests/test_cvpack.py:104",
        "text": "This force is not present in the given context.",
    },
}

if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "known_answers.json")
    with open(path, "w", encoding="utf-8") as file:
        json.dump(GOLDEN, file, indent=1)
    print("wrote", path)
