"""
Generates ``tests/golden/reference_vectors.npz`` + ``reference_vectors.json`` by running the
**reference's own code** (``/root/reference/cvpack``: constructors, ``addToSystem``, ``getValue``,
``getEffectiveMass``, YAML round trip) over the OpenMM stand-in of ``oracle/refshim`` on the seeded
synthetic systems of ``tests/golden/cases.py``.

Run here (the build container; ``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_reference_golden.py

For every case the file holds, per frame: the value (``cv.getValue(context)``), the gradient
(``-context.getState(getForces=True, groups=1<<g).getForces()``) and the effective mass
(``cv.getEffectiveMass(context)``), all FP64; plus the positions / box the case was evaluated on
and, in the JSON, what the reference constructor produced (energy expression, unit, serialized
``reference`` value) for traceability.  What is reference-produced and what is restated is spelled
out in ``oracle/refshim/fake_openmm.py``.
"""

from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402  pylint: disable=wrong-import-position
from tests.golden import cases  # noqa: E402  pylint: disable=wrong-import-position


def normalised_yaml(cv) -> str:
    """YAML text of a CV with the embedded force XML blanked (the stand-in's XML is a token; the
    product writes real OpenMM XML -- both sides are compared on everything else)."""
    import re  # pylint: disable=import-outside-toplevel

    import yaml  # pylint: disable=import-outside-toplevel

    return re.sub(r"xml_code:.*?(?=\n[A-Za-z]|\Z)", "xml_code: X", yaml.safe_dump(cv), flags=re.S)


def evaluate_reference(case: cases.Case, ref_ns) -> dict:
    """Run one case through the reference package; returns arrays and metadata."""
    import openmm  # the stand-in  pylint: disable=import-outside-toplevel
    import yaml  # pylint: disable=import-outside-toplevel

    context0 = None
    if case.needs_context:
        # a context of frame 0 *without* the CV, as a user would pass for ``reference=``
        probe = case.build(ref_ns, context=1.0)
        context0 = openmm.Context(probe.system, openmm.VerletIntegrator(1.0))
        context0.setPositions(probe.positions[0])
        if probe.box is not None:
            context0.setPeriodicBoxVectors(*probe.box)
    built = case.build(ref_ns, context=context0)
    cv, system = built.cv, built.system
    # the reference's perform_common_tests (tests/test_cvpack.py:75-86): a YAML round trip must
    # reproduce the variable; evaluate the round-tripped copy as well
    yaml_text = normalised_yaml(cv)
    clone = yaml.safe_load(yaml.safe_dump(cv))
    cv.addToSystem(system)
    clone.addToSystem(system)
    context = openmm.Context(system, openmm.VerletIntegrator(1.0))
    values, grads, emasses = [], [], []
    roundtrip_changed = False
    for frame in built.positions:
        context.setPositions(frame)
        if built.box is not None:
            context.setPeriodicBoxVectors(*built.box)
        value = cv.getValue(context)
        values.append(float(value / value.unit))
        state = context.getState(getForces=True, groups=1 << cv.getForceGroup())
        grads.append(-np.asarray(state.getForces(asNumpy=True)._value))  # pylint: disable=protected-access
        emass = cv.getEffectiveMass(context)
        emasses.append(float(emass / emass.unit))
        again = clone.getValue(context)
        roundtrip_changed = roundtrip_changed or float(again / again.unit) != values[-1]
    meta = {
        "row": case.row,
        "class": type(cv).__name__,
        "unit": str(cv.getUnit()),
        "mass_unit": cv.getMassUnit().get_symbol(),
        "name": cv.getName(),
        "num_atoms": int(built.positions.shape[1]),
        "num_frames": int(built.positions.shape[0]),
        # True only where the reference loses a constructor argument on serialization
        # (helix_torsion_content.py:147-156 does not record ``normalize``)
        "yaml_roundtrip_changes_value": bool(roundtrip_changed),
        # the serialized form must be the reference's byte for byte (same tags, keys, values)
        "yaml_sha256": hashlib.sha256(yaml_text.encode()).hexdigest(),
        "yaml_bytes": len(yaml_text),
    }
    if hasattr(cv, "getEnergyFunction"):
        text = cv.getEnergyFunction()
        meta["energy_function"] = text if len(text) <= 400 else text[:400] + f"... [{len(text)} chars]"
    if "reference" in cv._args:  # pylint: disable=protected-access
        meta["reference"] = float(cv._args["reference"])  # pylint: disable=protected-access
    return {
        "positions": built.positions,
        "box": np.zeros((0, 3)) if built.box is None else np.asarray(built.box, dtype=np.float64),
        "value": np.array(values),
        "gradient": np.stack(grads),
        "emass": np.array(emasses),
        "meta": meta,
    }


def main() -> None:
    ref_ns, _ = refshim.namespaces()
    arrays, index = {}, {}
    for case in cases.CASES:
        result = evaluate_reference(case, ref_ns)
        index[case.name] = result.pop("meta")
        for key, value in result.items():
            arrays[f"{case.name}/{key}"] = value
        print(f"{case.name:42s} value[0] = {arrays[case.name + '/value'][0]: .12g}")
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **arrays)
    header = {
        "generator": "tests/golden/make_reference_golden.py",
        "reference_version": ref_ns.cvpack.__version__,
        "produced_by": "reference constructors + getValue/getEffectiveMass over oracle/refshim "
        "(OpenMM Force semantics restated; see oracle/refshim/fake_openmm.py)",
        "cases": index,
    }
    with open(os.path.join(HERE, "reference_vectors.json"), "w", encoding="utf-8") as file:
        json.dump(header, file, indent=1, sort_keys=True)
    print(f"wrote {len(index)} cases")


if __name__ == "__main__":
    main()
