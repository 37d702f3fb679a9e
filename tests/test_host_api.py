"""
Host-side behaviour shared with the reference: units, serialization, constructor validation,
exception strings, force-group bookkeeping (cvpack/tests/test_cvpack.py:52-104 and doctests).
No GPU needed.
"""

import copy
import inspect
import io
import json
import os

import numpy as np
import pytest

import cvpack_b200 as cvpack
from cvpack_b200 import mmunit, serialization, units, utils
from cvpack_b200.step_functions import parse_step_function
from tests import helpers

with open(os.path.join(os.path.dirname(__file__), "golden", "known_answers.json"), encoding="utf-8") as _f:
    GOLDEN = json.load(_f)


def all_cvs():
    rng = np.random.default_rng(0)
    topology = helpers.make_peptide(12, glycines=(2,))
    residues = list(topology.residues())
    n = topology.getNumAtoms()
    nb = helpers.make_nonbonded(n, rng, periodic=True, switch=0.8, num_exceptions=5)
    x = rng.normal(size=(n, 3))
    milestones = [{a: x[a] + 0.1 * k for a in range(10)} for k in range(3)]
    return n, [
        cvpack.Distance(0, 1),
        cvpack.Angle(0, 1, 2, pbc=True),
        cvpack.Torsion(0, 1, 2, 3),
        cvpack.RadiusOfGyration(range(10), weighByMass=True),
        cvpack.RadiusOfGyrationSq(np.arange(10), pbc=True),
        cvpack.NumberOfContacts(range(5), range(5, 12), nb, switchFactor=None),
        cvpack.NumberOfContacts(range(5), range(5, 12), nb, stepFunction="step(1-x)", reference=2.5),
        cvpack.ResidueCoordination(residues[:4], residues[6:], includeHydrogens=False, normalize=True),
        cvpack.AttractionStrength(range(5), range(5, 12), nb, contrastGroup=range(20, 30), contrastScaling=0.5),
        cvpack.ShortestDistance(range(5), range(5, 12), n),
        cvpack.RMSD(x, range(8), n),
        cvpack.RMSD({a: x[a] for a in range(3, 9)}, range(3, 9), n),
        cvpack.CompositeRMSD(x, [range(4), range(6, 11)], n),
        cvpack.HelixTorsionContent(residues),
        cvpack.HelixAngleContent(residues, normalize=True),
        cvpack.HelixHBondContent(residues),
        cvpack.HelixRMSDContent(residues, n),
        cvpack.SheetRMSDContent(residues, n, parallel=True),
        cvpack.SheetRMSDContent(residues, n, blockSizes=[6, 6]),
        cvpack.PathInRMSDSpace(cvpack.path.progress, milestones, n, 0.5 * mmunit.angstroms),
        cvpack.PathInRMSDSpace(cvpack.path.deviation, milestones, n, 0.05),
        cvpack.TorsionSimilarity([[0, 1, 2, 3], (4, 5, 6, 7)], np.array([[4, 5, 6, 7], [8, 9, 10, 11]])),
        cvpack.PathInCVSpace(
            cvpack.path.progress, [cvpack.Torsion(0, 1, 2, 3), cvpack.Distance(0, 5)],
            [[1.3, 0.2], [1.2, 0.4], [-2.7, 0.5]], np.pi / 6, scales=[1.0, 0.1],
        ),
        cvpack.PathInCVSpace(
            cvpack.path.deviation, [cvpack.Torsion(0, 1, 2, 3), cvpack.Torsion(1, 2, 3, 4)],
            [[1.3, -0.2], [1.2, 3.1]], 0.5,
        ),
    ]


def test_units_format_like_reference():
    assert cvpack.Torsion(0, 1, 2, 3).getMassUnit().get_symbol() == GOLDEN["mass_units"]["rad"]
    assert cvpack.Distance(0, 1).getMassUnit().get_symbol() == GOLDEN["mass_units"]["nm"]
    assert cvpack.RadiusOfGyrationSq([0, 1]).getUnit().get_symbol() == "nm**2"
    assert repr(units.Quantity(0.3, mmunit.nanometers)) == "0.3 nm"
    assert str(units.Unit("kilojoule/(mole*radian**2)").get_symbol()) == "kJ/(mol rad**2)"
    assert (5 * mmunit.angstroms).value_in_unit(mmunit.nanometers) == pytest.approx(0.5)
    assert units.value_in_md_units(90 * mmunit.degrees) == pytest.approx(np.pi / 2)
    assert units.value_in_md_units(1 * mmunit.kilocalories_per_mole) == pytest.approx(4.184)
    assert units.value_in_md_units(3.5) == 3.5
    assert units.in_md_units(2 * mmunit.angstroms).value == pytest.approx(0.2)
    assert units.in_md_units(1.5).unit.is_dimensionless()
    with pytest.raises(TypeError):
        (1 * mmunit.nanometers).value_in_unit(mmunit.picoseconds)


def test_preprocess_args_conversions():
    """The doctest of cvpack/utils.py:203-215."""

    @utils.preprocess_args
    def function(data):
        return data

    assert isinstance(function(mmunit.angstrom), units.Unit)
    assert isinstance(function(5 * mmunit.angstrom), units.Quantity)
    seq = [mmunit.angstrom, mmunit.nanometer]
    assert isinstance(function(seq), list)
    assert all(isinstance(item, units.Unit) for item in function(seq))
    dct = {"length": 3 * mmunit.angstrom, "time": 2 * mmunit.picosecond}
    assert isinstance(function(dct), dict)
    assert all(isinstance(item, units.Quantity) for item in function(dct).values())
    assert type(function(np.int64(3))) is int and type(function(np.float32(3))) is float


def test_golden_yaml_text():
    stream = io.StringIO()
    serialization.serialize(cvpack.RadiusOfGyration([0, 1, 2]), stream)
    assert stream.getvalue() == GOLDEN["serialization_radius_of_gyration"]["text"]
    stream.seek(0)
    assert type(serialization.deserialize(stream)) is cvpack.RadiusOfGyration


def test_every_cv_round_trips_and_is_annotated():
    """perform_common_tests of the reference, minus the evaluations (covered on the GPU)."""
    _, cvs = all_cvs()
    for cv in cvs:
        unity = 1 * cv.getUnit()
        assert unity.value_in_unit_system(mmunit.md_unit_system) == 1
        for _, parameter in inspect.signature(cv.__init__).parameters.items():
            assert parameter.annotation is not inspect.Parameter.empty
        assert cv.yaml_tag.startswith("!cvpack.")
        pipe = io.StringIO()
        serialization.serialize(cv, pipe)
        text = pipe.getvalue()
        assert text.startswith(cv.yaml_tag + "\n")
        pipe.seek(0)
        new_cv = serialization.deserialize(pipe)
        assert type(new_cv) is type(cv)
        assert new_cv.getName() == cv.getName()
        assert new_cv.getPeriodicBounds() == cv.getPeriodicBounds()
        assert new_cv.getUnit() == cv.getUnit() and new_cv.getMassUnit() == cv.getMassUnit()
        again = io.StringIO()
        serialization.serialize(new_cv, again)
        assert again.getvalue() == text  # serialization is a fixed point
        clone = copy.deepcopy(cv)
        assert type(clone) is type(cv) and clone is not cv


def test_yaml_tags_match_reference():
    expected = {
        "Angle", "AttractionStrength", "CompositeRMSD", "Distance", "HelixAngleContent",
        "HelixHBondContent", "HelixRMSDContent", "HelixTorsionContent", "NumberOfContacts",
        "PathInRMSDSpace", "RadiusOfGyration", "RadiusOfGyrationSq", "ResidueCoordination", "RMSD",
        "SheetRMSDContent", "ShortestDistance", "Torsion", "TorsionSimilarity", "PathInCVSpace",
    }
    for name in expected:
        assert getattr(cvpack, name).yaml_tag == f"!cvpack.{name}"
    assert cvpack.path.Metric.yaml_tag == "!cvpack.path.Metric"
    assert units.Unit.yaml_tag == "!cvpack.Unit" and units.Quantity.yaml_tag == "!cvpack.Quantity"
    assert serialization.SerializableAtom.yaml_tag == "!cvpack.Atom"
    assert serialization.SerializableResidue.yaml_tag == "!cvpack.Residue"
    assert serialization.SerializableForce.yaml_tag == "!cvpack.Force"


def test_default_names_and_bounds():
    n, cvs = all_cvs()
    names = [cv.getName() for cv in cvs]
    assert names[:5] == ["distance", "angle", "torsion", "radius_of_gyration", "radius_of_gyration_sq"]
    assert "path_progress_in_rmsd_space" in names and "path_deviation_in_rmsd_space" in names
    bounds = cvpack.Torsion(0, 1, 2, 3).getPeriodicBounds()
    assert bounds.value == (-np.pi, np.pi) and bounds.unit.get_symbol() == "rad"
    assert cvpack.Distance(0, 1).getPeriodicBounds() is None
    assert n > 0


def test_exception_strings():
    rng = np.random.default_rng(1)
    topology = helpers.make_peptide(8)
    residues = list(topology.residues())
    n = topology.getNumAtoms()
    stripped = helpers.make_peptide(8)
    for residue in stripped.residues():
        residue._atoms = [a for a in residue._atoms if a.name != "CA"]  # pylint: disable=protected-access
    bad = list(stripped.residues())
    with pytest.raises(ValueError) as info:
        cvpack.HelixTorsionContent(bad)
    assert str(info.value) == "Could not find atom CA in residue ALA2"
    with pytest.raises(ValueError) as info:
        cvpack.HelixAngleContent(bad)
    assert str(info.value) == "Could not find atom CA in residue ALA1"
    with pytest.raises(ValueError) as info:
        cvpack.HelixRMSDContent(residues[:5], n)
    assert str(info.value) == "5 residues yield 0 blocks, which is not between 1 and 1024"
    with pytest.raises(ValueError) as info:
        cvpack.SheetRMSDContent(residues, n, blockSizes=[3, 3])
    assert str(info.value) == "The sum of block sizes (6) and the number of residues (8) must be equal."
    x = rng.normal(size=(n, 3))
    with pytest.raises(ValueError) as info:
        cvpack.PathInRMSDSpace(cvpack.path.progress, [{0: x[0]}], n, 0.1)
    assert str(info.value) == "At least two reference structures are required."
    with pytest.raises(ValueError) as info:
        cvpack.PathInRMSDSpace("progress", [{0: x[0]}, {0: x[1]}], n, 0.1)
    assert str(info.value) == "Invalid metric. Use 'cvpack.path.progress' or 'cvpack.path.deviation'."
    with pytest.raises(NotImplementedError):
        cvpack.NumberOfContacts([0], [1], helpers.make_nonbonded(4, rng, False), stepFunction="exp(-x)")


def test_force_groups_and_system():
    system = cvpack.System()
    for _ in range(4):
        system.addParticle(12.0 * mmunit.daltons)
    assert system.getParticleMass(2).value_in_unit(mmunit.daltons) == 12.0
    cvs = [cvpack.Distance(0, 1, name=f"d{k}") for k in range(32)]
    for k, cv in enumerate(cvs):
        cv.addToSystem(system)
        assert cv.getForceGroup() == k  # lowest unused group
    with pytest.raises(RuntimeError) as info:
        cvpack.Distance(0, 1).addToSystem(system)
    assert str(info.value) == "All force groups are already in use."
    assert system.getNumForces() == 32


def test_step_function_whitelist():
    assert parse_step_function("1/(1+x^6)") == (0, 6.0, 0.0)
    assert parse_step_function("1 / (1 + x**12)") == (0, 12.0, 0.0)
    assert parse_step_function("step(1-x)") == (1, 0.0, 0.0)
    assert parse_step_function("(1+x^4)/(1+x^4+x^8)") == (2, 4.0, 0.0)
    assert parse_step_function("(1-x^6)/(1-x^12)") == (3, 6.0, 12.0)
    for bad in ("(1+x^4)/(1+x^4+x^9)", "1/(1+y^6)", "exp(-x)", "1/(1+x^0)"):
        with pytest.raises(NotImplementedError):
            parse_step_function(bad)


def test_nonbonded_force_xml_round_trip():
    rng = np.random.default_rng(2)
    force = helpers.make_nonbonded(6, rng, periodic=True, cutoff=0.9, switch=0.7, num_exceptions=3)
    clone = cvpack.mm.XmlSerializer.deserialize(cvpack.mm.XmlSerializer.serialize(force))
    assert clone.getNumParticles() == 6 and clone.getNumExceptions() == 3
    assert clone.usesPeriodicBoundaryConditions() and clone.getUseSwitchingFunction()
    for k in range(6):
        a = [units.value_in_md_units(p) for p in force.getParticleParameters(k)]
        b = [units.value_in_md_units(p) for p in clone.getParticleParameters(k)]
        assert a == b  # repr round trip is exact
    assert [clone.getExceptionParameters(k)[:2] for k in range(3)] == [
        force.getExceptionParameters(k)[:2] for k in range(3)
    ]


def test_batch_context_validation_without_gpu():
    system = cvpack.System()
    for _ in range(3):
        system.addParticle(1.0)
    with pytest.raises(ValueError):
        cvpack.BatchContext(system, device="cpu")
    assert cvpack.shard_frames(10, 0, 4) == (0, 3)
    assert cvpack.shard_frames(10, 3, 4) == (8, 10)
    covered = sum((list(range(*cvpack.shard_frames(4097, r, 8))) for r in range(8)), [])
    assert covered == list(range(4097))
