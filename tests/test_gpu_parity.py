"""
GPU parity: every collective variable, evaluated by the CUDA library through the public Python API
(ctypes -> C ABI), against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): 1e-5 relative in FP32 mode, 1e-10 in FP64 mode.  Values are
compared relative to |value|; gradients relative to the largest gradient component of the frame
(per-element relative error is meaningless for components that cancel to ~0).  The oracle sees the
coordinates after rounding to the GPU's storage precision.
"""

import io

import numpy as np
import pytest
import torch

import cvpack_b200 as cvpack
from cvpack_b200 import _native, mmunit, serialization
from tests import helpers

pytestmark = pytest.mark.gpu

TOL = {"single": (1e-5, 1e-5), "double": (1e-10, 1e-10)}
ORTHO = np.diag([2.4, 2.6, 2.8])
TRICLINIC = np.array([[2.4, 0, 0], [0.3, 2.6, 0], [-0.4, 0.5, 2.8]])
BOXES = {"none": None, "ortho": ORTHO, "triclinic": TRICLINIC}


def frames(rng, num_frames, num_atoms, scale=2.4):
    return rng.uniform(0.0, scale, size=(num_frames, num_atoms, 3))


def run(cv, system, positions, box, precision):
    vt, gt = TOL[precision]
    return helpers.check_against_oracle(cv, system, positions, box, precision, vt, gt)


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("boxname", ["none", "ortho", "triclinic"])
def test_bonded(precision, boxname):
    rng = np.random.default_rng(11)
    box = BOXES[boxname]
    system = helpers.make_system(40, rng)
    x = frames(rng, 5, 40)
    pbc = box is not None
    run(cvpack.Distance(3, 17, pbc=pbc), system, x, box, precision)
    run(cvpack.Angle(3, 17, 5, pbc=pbc), system, x, box, precision)
    run(cvpack.Torsion(3, 17, 5, 9, pbc=pbc), system, x, box, precision)


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("boxname", ["none", "ortho", "triclinic"])
def test_torsion_similarity(precision, boxname):
    rng = np.random.default_rng(12)
    box = BOXES[boxname]
    system = helpers.make_system(40, rng)
    x = frames(rng, 6, 40)
    first = [list(rng.choice(40, size=4, replace=False)) for _ in range(9)]
    second = [list(rng.choice(40, size=4, replace=False)) for _ in range(9)]
    second[0] = first[0]  # a pair of identical torsions contributes exactly 1 and no gradient
    cv = cvpack.TorsionSimilarity(first, second, pbc=box is not None)
    assert cv.getName() == "torsion_similarity" and cv.getUnit().is_dimensionless()
    run(cv, system, x, box, precision)


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("metric", ["progress", "deviation"])
def test_path_in_cv_space(precision, metric):
    """Torsion (periodic) x Distance x RadiusOfGyration space, scaled; the inner variables run their
    own kernels and the combination applies the chain rule (path_in_cv_space.py:141-158)."""
    rng = np.random.default_rng(13)
    system = helpers.make_system(30, rng)
    x = frames(rng, 7, 30, scale=1.5)
    variables = [cvpack.Torsion(3, 17, 5, 9), cvpack.Distance(2, 11), cvpack.RadiusOfGyration(range(8, 20))]
    milestones = np.array([[3.0, 0.4, 0.5], [-3.0, 0.8, 0.6], [-2.0, 1.2, 0.7], [0.5, 1.6, 0.8], [1.5, 2.0, 0.9]])
    cv = cvpack.PathInCVSpace(getattr(cvpack.path, metric), variables, milestones, 0.8,
                              scales=[1.0, 0.5 * mmunit.nanometers, 0.2])
    assert cv.getName() == f"path_{metric}_in_cv_space" and cv.getUnit().is_dimensionless()
    vt, gt = TOL[precision]
    ctx = helpers.check_against_oracle(cv, system, x, None, precision, vt, gt)
    emass = cv.getEffectiveMass(ctx)
    assert torch.isfinite(emass._value).all()  # pylint: disable=protected-access
    with pytest.raises(ValueError, match="Wrong number of columns"):
        cvpack.PathInCVSpace(cvpack.path.progress, variables, milestones[:, :2], 0.8)
    with pytest.raises(ValueError, match="At least two rows"):
        cvpack.PathInCVSpace(cvpack.path.progress, variables, milestones[:1], 0.8)


def test_known_answers():
    """distance.py:37-50, angle.py:46-60, torsion.py:53-67 of the reference."""
    system = cvpack.System()
    for _ in range(4):
        system.addParticle(1.0)
    distance = cvpack.Distance(0, 1)
    angle = cvpack.Angle(0, 1, 2)
    torsion = cvpack.Torsion(0, 1, 2, 3)
    for cv in (distance, angle, torsion):
        cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, [[0, 0, 0], [1, 1, 1], [0, 0, 0], [0, 0, 0]], precision="double")
    assert float(distance.getValue(ctx).value_in_unit(mmunit.nanometers)[0]) == pytest.approx(np.sqrt(3))
    ctx.setPositions([[0, 0, 0], [0, 0, 1], [0, 1, 1], [0, 0, 0]])
    assert float(angle.getValue(ctx).value_in_unit(mmunit.radians)[0]) == pytest.approx(np.pi / 2)
    ctx.setPositions([[0, 0, 0], [1, 0, 0], [1, 1, 0], [1, 1, 1]])
    assert float(torsion.getValue(ctx).value_in_unit(mmunit.radians)[0]) == pytest.approx(np.pi / 2)
    assert str(distance.getValue(ctx).unit.get_symbol()) == "nm"


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("boxname", ["none", "ortho", "triclinic"])
@pytest.mark.parametrize("weigh", [False, True])
def test_radius_of_gyration(precision, boxname, weigh):
    rng = np.random.default_rng(12)
    box = BOXES[boxname]
    system = helpers.make_system(300, rng)
    x = frames(rng, 4, 300, scale=1.0) + 5.0  # far from the origin: exercises the cancellation
    pbc = box is not None
    group = rng.permutation(300)[:120]
    run(cvpack.RadiusOfGyration(group, pbc=pbc, weighByMass=weigh), system, x, box, precision)
    run(cvpack.RadiusOfGyrationSq(range(300), pbc=pbc, weighByMass=weigh), system, x, box, precision)


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("boxname", ["none", "ortho", "triclinic"])
@pytest.mark.parametrize("case", ["disjoint", "overlap", "same", "noswitch", "heaviside"])
def test_number_of_contacts(precision, boxname, case):
    rng = np.random.default_rng(13)
    box = BOXES[boxname]
    n = 700
    system = helpers.make_system(n, rng)
    x = frames(rng, 3, n)
    nb = helpers.make_nonbonded(n, rng, periodic=box is not None, num_exceptions=60)
    if case == "disjoint":
        g1, g2 = range(0, 300), range(300, 700)
    elif case == "overlap":
        g1, g2 = range(0, 400), range(250, 700)
    elif case == "same":
        g1 = g2 = range(100, 420)
    else:
        g1, g2 = rng.permutation(n)[:333], rng.permutation(n)[:290]
    kwargs = {}
    if case == "noswitch":
        kwargs["switchFactor"] = None
    if case == "heaviside":
        kwargs["stepFunction"] = "step(1-x)"
    cv = cvpack.NumberOfContacts(g1, g2, nb, thresholdDistance=0.3 * mmunit.nanometers, **kwargs)
    vt, gt = TOL[precision]
    if case in ("noswitch", "heaviside"):
        # a sharp cutoff makes the value discontinuous: pairs within rounding of r_c may flip
        vt = max(vt, 5e-4)
    helpers.check_against_oracle(cv, system, x, box, precision, vt, gt)


@pytest.mark.parametrize("precision", ["single", "double"])
def test_number_of_contacts_reference_context(precision):
    rng = np.random.default_rng(14)
    n = 200
    system = helpers.make_system(n, rng)
    x = frames(rng, 2, n, scale=1.2)
    nb = helpers.make_nonbonded(n, rng, periodic=False)
    plain = cvpack.NumberOfContacts(range(0, 100), range(100, 200), nb)
    plain.addToSystem(system)
    ctx = cvpack.BatchContext(system, x, precision=precision)
    raw = plain.getValue(ctx)
    normalized = cvpack.NumberOfContacts(range(0, 100), range(100, 200), nb, reference=ctx)
    normalized.addToSystem(system)
    ctx.reinitialize(preserveState=True)
    value = normalized.getValue(ctx)
    assert float(value._value[0]) == pytest.approx(1.0, rel=1e-5)
    assert float(value._value[1]) == pytest.approx(float(raw._value[1] / raw._value[0]), rel=1e-5)


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("boxname", ["none", "ortho"])
@pytest.mark.parametrize("contrast", [False, True])
def test_attraction_strength(precision, boxname, contrast):
    rng = np.random.default_rng(15)
    box = BOXES[boxname]
    n = 400
    system = helpers.make_system(n, rng)
    # keep atoms apart (a lattice with jitter): the soft-core LJ is stiff at tiny separations
    grid = np.stack(np.meshgrid(*[np.arange(8)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n] * 0.3
    x = grid[None] + rng.normal(scale=0.03, size=(3, n, 3))
    nb = helpers.make_nonbonded(n, rng, periodic=box is not None, cutoff=0.9, switch=0.7, num_exceptions=40)
    cv = cvpack.AttractionStrength(
        range(0, 120), range(120, 260), nb,
        contrastGroup=range(200, 400) if contrast else None, contrastScaling=0.5,
        reference=25.0 * mmunit.kilojoule_per_mole,
    )
    vt, gt = TOL[precision]
    helpers.check_against_oracle(cv, system, x, box, precision, vt, gt)


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("boxname", ["none", "ortho", "triclinic"])
def test_shortest_distance(precision, boxname):
    rng = np.random.default_rng(16)
    box = BOXES[boxname]
    n = 300
    system = helpers.make_system(n, rng)
    x = frames(rng, 4, n)
    cv = cvpack.ShortestDistance(range(0, 100), range(150, 300), n, beta=100, pbc=box is not None)
    run(cv, system, x, box, precision)


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("boxname", ["none", "ortho"])
@pytest.mark.parametrize("hydrogens,weigh", [(True, True), (False, False)])
def test_residue_coordination(precision, boxname, hydrogens, weigh):
    rng = np.random.default_rng(17)
    box = BOXES[boxname]
    topology = helpers.make_peptide(40, glycines=(3, 17))
    n = topology.getNumAtoms()
    system = helpers.make_system(n, rng)
    x = frames(rng, 3, n)
    residues = list(topology.residues())
    cv = cvpack.ResidueCoordination(
        residues[:15], residues[20:], pbc=box is not None, thresholdDistance=0.8,
        normalize=True, weighByMass=weigh, includeHydrogens=hydrogens,
    )
    run(cv, system, x, box, precision)
    assert cv.getReferenceValue()._value == 15 * 20


@pytest.mark.parametrize("precision", ["single", "double"])
def test_rmsd(precision):
    """The scipy check of tests/test_cvpack.py:275-338 (full / half / shuffled groups)."""
    from scipy.spatial.transform import Rotation

    rng = np.random.default_rng(18)
    n = 200
    system = helpers.make_system(n, rng)
    reference = rng.normal(size=(n, 3))
    rot = Rotation.random(5, random_state=3).as_matrix()
    x = np.einsum("bij,nj->bni", rot, reference) + 0.05 * rng.normal(size=(5, n, 3)) + 7.0
    for group in (np.arange(n), np.arange(n // 2), rng.permutation(n)[: n // 2]):
        for as_dict in (False, True):
            ref = dict(zip(group.tolist(), reference[group])) if as_dict else reference
            cv = cvpack.RMSD(ref, group, n)
            ctx = run(cv, system, x, None, precision)
            value = cv.getValue(ctx)._value.double().cpu().numpy()
            stored = ctx.getPositions().double().cpu().numpy()
            for b in range(5):
                xs = stored[b][group] - stored[b][group].mean(0)
                ys = reference[group] - reference[group].mean(0)
                _, rssd = Rotation.align_vectors(xs, ys)
                assert value[b] == pytest.approx(rssd / np.sqrt(len(group)), rel=TOL[precision][0] * 3)


@pytest.mark.parametrize("precision", ["single", "double"])
def test_rmsd_degenerate(precision):
    """msd < 1e-20 -> value 0 and zero gradient (OpenMM RMSDForce); FP32 noise keeps it ~0."""
    rng = np.random.default_rng(19)
    n = 64
    system = helpers.make_system(n, rng)
    reference = rng.normal(size=(n, 3))
    cv = cvpack.RMSD(reference, range(n), n)
    cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, reference[None], precision=precision)
    value, grad = cv.getValueAndGradient(ctx)
    assert float(value._value[0]) < (1e-3 if precision == "single" else 1e-7)
    assert np.isfinite(grad.cpu().numpy()).all()


@pytest.mark.parametrize("precision", ["single", "double"])
def test_composite_rmsd(precision):
    rng = np.random.default_rng(20)
    n = 260
    system = helpers.make_system(n, rng)
    reference = rng.normal(size=(n, 3))
    groups = [list(range(0, 90)), list(rng.permutation(np.arange(100, 220))[:70]), [230, 231, 240]]
    x = reference[None] + 0.05 * rng.normal(size=(4, n, 3))
    x[:, groups[1]] += 1.0  # translating one body must not matter (composite_rmsd.py:116-121)
    cv = cvpack.CompositeRMSD(reference, groups, n)
    ctx = run(cv, system, x, None, precision)
    # doctest invariance: 0.0 at the reference, still 0.0 after translating one body by 1 nm
    moved = reference.copy()
    moved[groups[1]] += 1.0
    ctx.setPositions(np.stack([reference, moved]))
    value = cv.getValue(ctx)._value.cpu().numpy()
    assert np.all(np.abs(value) < (2e-3 if precision == "single" else 1e-7))
    with pytest.raises(ValueError):
        cvpack.CompositeRMSD(reference, [[0, 1, 2], [2, 3, 4]], n)


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("normalize", [False, True])
def test_helix_rmsd_content(precision, normalize):
    rng = np.random.default_rng(21)
    topology = helpers.make_peptide(14, glycines=(4,))
    n = topology.getNumAtoms()
    system = helpers.make_system(n, rng)
    base = helpers.ideal_helix_positions(14, rng)
    x = base[None] + rng.normal(scale=0.02, size=(4, n, 3))
    cv = cvpack.HelixRMSDContent(list(topology.residues()), n, normalize=normalize)
    assert cv.getNumResidueBlocks() == 9
    run(cv, system, x, None, precision)
    # ideal coordinates: every block has rmsd ~ 0 -> S = 1 per block
    cv2 = cvpack.HelixRMSDContent(list(topology.residues()), n)
    cv2.addToSystem(system)
    ctx = cvpack.BatchContext(system, base[None], precision=precision)
    assert float(cv2.getValue(ctx)._value[0]) == pytest.approx(9.0, rel=2e-3)


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("parallel", [False, True])
def test_sheet_rmsd_content(precision, parallel):
    rng = np.random.default_rng(22)
    topology = helpers.make_peptide(12)
    n = topology.getNumAtoms()
    system = helpers.make_system(n, rng)
    x = frames(rng, 3, n, scale=0.8)
    cv = cvpack.SheetRMSDContent(list(topology.residues()), n, parallel=parallel, thresholdRMSD=0.5)
    run(cv, system, x, None, precision)
    cv = cvpack.SheetRMSDContent(list(topology.residues()), n, parallel=parallel, blockSizes=[6, 6], thresholdRMSD=0.5)
    run(cv, system, x, None, precision)


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("metric", ["progress", "deviation"])
def test_path_in_rmsd_space(precision, metric):
    rng = np.random.default_rng(23)
    n = 150
    system = helpers.make_system(n, rng)
    start, end = rng.normal(size=(n, 3)), rng.normal(size=(n, 3))
    atoms = rng.permutation(n)[:100]
    milestones = []
    for k in range(7):
        conf = start + (end - start) * k / 6
        order = rng.permutation(atoms) if k % 2 else atoms  # each milestone keeps its own order
        milestones.append({int(a): conf[a] for a in order})
    lam = np.linspace(0.1, 0.9, 5)[:, None, None]
    x = start[None] + (end - start)[None] * lam + 0.02 * rng.normal(size=(5, n, 3))
    cv = cvpack.PathInRMSDSpace(getattr(cvpack.path, metric), milestones, n, 0.3 * mmunit.nanometers)
    assert cv.getName() == f"path_{metric}_in_rmsd_space"
    vt, gt = TOL[precision]
    helpers.check_against_oracle(cv, system, x, None, precision, vt, gt)


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("normalize", [False, True])
def test_helix_bonded_contents(precision, normalize):
    rng = np.random.default_rng(24)
    topology = helpers.make_peptide(16)
    n = topology.getNumAtoms()
    system = helpers.make_system(n, rng)
    base = helpers.ideal_helix_positions(16, rng)
    x = base[None] + rng.normal(scale=0.02, size=(4, n, 3))
    residues = list(topology.residues())
    run(cvpack.HelixTorsionContent(residues, normalize=normalize), system, x, None, precision)
    run(cvpack.HelixAngleContent(residues, normalize=normalize), system, x, None, precision)
    run(cvpack.HelixHBondContent(residues, normalize=normalize), system, x, None, precision)


@pytest.mark.parametrize("precision", ["single", "double"])
def test_effective_mass_and_roundtrip(precision):
    """perform_common_tests of the reference (tests/test_cvpack.py:52-88), batched."""
    from oracle import cv_oracle

    rng = np.random.default_rng(25)
    n = 120
    system = helpers.make_system(n, rng)
    x = frames(rng, 6, n, scale=1.5)
    nb = helpers.make_nonbonded(n, rng, periodic=False, num_exceptions=10)
    cvs = [
        cvpack.RadiusOfGyration(range(10, 90)),
        cvpack.Torsion(1, 2, 3, 4),
        cvpack.NumberOfContacts(range(0, 60), range(60, 120), nb),
        cvpack.RMSD(x[0], range(0, 50), n),
    ]
    for cv in cvs:
        cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, x, precision=precision)
    for cv in cvs:
        unity = 1 * cv.getUnit()
        assert unity.value_in_unit_system(mmunit.md_unit_system) == 1
        pipe = io.StringIO()
        serialization.serialize(cv, pipe)
        pipe.seek(0)
        new_cv = serialization.deserialize(pipe)
        new_cv.addToSystem(ctx.getSystem())
        ctx.reinitialize(preserveState=True)
        value1, value2 = cv.getValue(ctx), new_cv.getValue(ctx)
        assert bool((value1 / value1.unit == value2 / value2.unit).all())  # bit-identical
        mass1, mass2 = cv.getEffectiveMass(ctx), new_cv.getEffectiveMass(ctx)
        assert bool((mass1 / mass1.unit == mass2 / mass2.unit).all())
        assert cv.getName() == new_cv.getName()
        assert cv.getPeriodicBounds() == new_cv.getPeriodicBounds()
        grad = cv.getGradient(ctx).double().cpu().numpy()
        emass = (mass1 / mass1.unit).double().cpu().numpy()
        for b in range(x.shape[0]):
            expected = cv_oracle.effective_mass(grad[b], system.getMasses())
            assert emass[b] == pytest.approx(expected, rel=1e-5 if precision == "single" else 1e-12)


def test_not_in_context_and_errors():
    rng = np.random.default_rng(26)
    system = helpers.make_system(30, rng)
    ctx = cvpack.BatchContext(system, frames(rng, 2, 30))
    cv = cvpack.RadiusOfGyration(range(30))
    with pytest.raises(RuntimeError) as excinfo:
        cv.getValue(ctx)
    assert str(excinfo.value) == "This force is not present in the given context."
    pbc_cv = cvpack.Distance(0, 1, pbc=True)
    pbc_cv.addToSystem(system)
    with pytest.raises(ValueError):
        pbc_cv.getValue(ctx)  # periodic CV but no box


def test_host_residency_matches_device():
    rng = np.random.default_rng(27)
    n = 500
    system = helpers.make_system(n, rng)
    x = frames(rng, 37, n)
    nb = helpers.make_nonbonded(n, rng, periodic=True)
    cv = cvpack.NumberOfContacts(range(0, 250), range(250, 500), nb)
    cv.addToSystem(system)
    dev = cvpack.BatchContext(system, x, ORTHO)
    host = cvpack.BatchContext(system, x, ORTHO, residency="host")
    v1, g1 = cv.getValueAndGradient(dev)
    v2, g2 = cv.getValueAndGradient(host)
    assert not v2._value.is_cuda and not g2.is_cuda
    assert bool((v1._value.cpu() == v2._value).all())
    assert bool((g1.cpu() == g2).all())


@pytest.mark.parametrize("precision", ["single", "double"])
def test_config0_solvated_dipeptide_single_frame(precision):
    """BASELINE.json configs[0] (openmmtools AlanineDipeptideExplicit, 2269 atoms, single frame:
    Distance, Torsion, RadiusOfGyration, RMSD, NumberOfContacts).  openmmtools' coordinates are not
    available offline, so the frame is the synthetic stand-in SURVEY 8(d) specifies:
    a 22-atom solute chain plus 749 three-site waters on a jittered lattice, seed 1, cubic box."""
    rng = np.random.default_rng(1)
    box_length = 2.84
    solute = [np.full(3, 0.5 * box_length)]
    for _ in range(21):
        step = rng.normal(size=3)
        solute.append(solute[-1] + 0.15 * step / np.linalg.norm(step))
    m = 10
    grid = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)
    sites = (grid[rng.permutation(m**3)[:749]] + 0.5) * (box_length / m)
    sites = sites + 0.03 * rng.normal(size=sites.shape)
    waters = []
    for oxygen in sites:
        waters.append(oxygen)
        for _ in range(2):
            h = rng.normal(size=3)
            waters.append(oxygen + 0.1 * h / np.linalg.norm(h))
    x = np.concatenate([np.array(solute), np.array(waters)])
    n = x.shape[0]
    assert n == 2269
    system = helpers.make_system(n, rng)
    box = np.diag([box_length] * 3)
    nb = helpers.make_nonbonded(n, rng, periodic=True, cutoff=0.9)
    reference = x[:22] + 0.02 * rng.normal(size=(22, 3))
    oxygens = range(22, n, 3)
    cvs = [
        cvpack.Distance(0, 21),
        cvpack.Torsion(4, 6, 8, 14),   # phi of alanine dipeptide
        cvpack.Torsion(6, 8, 14, 16),  # psi
        cvpack.RadiusOfGyration(range(22)),
        cvpack.RMSD({i: reference[i] for i in range(22)}, range(22), n),
        cvpack.NumberOfContacts(range(22), oxygens, nb, thresholdDistance=0.3),
    ]
    ctx = None
    for cv in cvs:
        cv.addToSystem(system)
    for cv in cvs:
        vt, gt = TOL[precision]
        if ctx is None:
            ctx = cvpack.BatchContext(system, x[None], box, precision=precision)
        helpers.check_against_oracle(cv, system, x[None], box, precision, vt, gt, context=ctx)
    emass = cvs[3].getEffectiveMass(ctx)
    assert np.isfinite(float(emass._value[0])) and float(emass._value[0]) > 0  # pylint: disable=protected-access


@pytest.mark.parametrize("precision", ["single", "double"])
def test_rmsd_content_one_cta_per_frame_equals_generic_kernels(precision):
    """The shared-memory path of the content CVs (`rmsd_blocks_kernel`) against the four generic
    kernels on the same frames: 200 residues (195 overlapping blocks), values and gradients; and
    against the oracle on a few frames."""
    rng = np.random.default_rng(61)
    nres = 200
    topology = helpers.make_peptide(nres, glycines=(7, 120))
    n = topology.getNumAtoms()
    system = helpers.make_system(n, rng)
    base = helpers.ideal_helix_positions(nres, rng)
    x = base[None] + rng.normal(scale=0.03, size=(70, n, 3))
    x[5] = base  # an ideal frame: every block at rmsd ~ 0
    cv = cvpack.HelixRMSDContent(list(topology.residues()), n)
    cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, x, precision=precision)
    native = ctx._native_cv(cv)  # pylint: disable=protected-access
    launches = _native.launch_count()
    value, grad = cv.getValueAndGradient(ctx)
    assert _native.launch_count() - launches == 1  # one kernel for the whole batch
    value, grad = value._value.clone(), grad.clone()
    assert native.set_option("blocks", 0) == 1
    value0, grad0 = cv.getValueAndGradient(ctx)
    native.set_option("blocks", 1)
    tol = 1e-5 if precision == "single" else 1e-10
    assert torch.allclose(value, value0._value, rtol=tol, atol=0)
    scale = float(grad0.abs().max())
    # (the two paths add the 13 sums of a block in different orders; near rmsd = 0 the gradient
    # (x - c - U^T y) / rmsd amplifies that, so the paths agree to 1e-8 in FP64 mode while each
    # meets the 1e-10 bar against the oracle on the frames checked below)
    assert float((grad - grad0).abs().max()) <= (tol if precision == "single" else 1e-8) * scale
    # atoms outside the blocks (H of every residue) carry exactly zero gradient
    outside = [a.index for a in topology.atoms() if a.name == "H"]
    assert float(grad[:, outside].abs().max()) == 0.0
    # (not the ideal frame: at rmsd ~ 1e-3 nm of fit residual the gradient is conditioned 1e4 worse)
    sub = cvpack.BatchContext(system, x[10:14], precision=precision)
    helpers.check_against_oracle(cv, system, x[10:14], None, precision, tol, tol * 5, context=sub)
    # value only
    only = cv.getValue(ctx)._value
    assert torch.equal(only, value)
