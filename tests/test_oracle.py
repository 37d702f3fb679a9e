"""
Pins the CPU oracle (oracle/cv_oracle.py, oracle/cv_oracle.c) against the reference's own
known-answer values (tests/golden/known_answers.json) and against independent recomputations that
restate what the reference's test-suite computes (cvpack/tests/test_cvpack.py), plus central finite
differences for every gradient.  No GPU needed.
"""

import itertools as it
import json
import os

import numpy as np
import pytest
from scipy.spatial.transform import Rotation
from scipy.special import logsumexp

from oracle import c_oracle
from oracle import cv_oracle as o

with open(os.path.join(os.path.dirname(__file__), "golden", "known_answers.json"), encoding="utf-8") as _f:
    GOLDEN = json.load(_f)

TRICLINIC = np.array([[2.0, 0, 0], [0.3, 2.2, 0], [-0.4, 0.5, 2.4]])


def fd_check(fn, x, atoms, tol=1e-6):
    value, grad = fn(x)
    fd = o.finite_difference(lambda y: fn(y)[0], x, atoms)
    scale = max(np.abs(fd).max(), 1e-12)
    assert np.abs(grad - fd).max() / scale < tol
    return value


def test_golden_bonded_terms():
    for name, fn in (("distance", o.distance), ("angle", o.angle), ("torsion", o.torsion)):
        entry = GOLDEN[name]
        x = np.array(entry["positions"], dtype=float)
        assert fn(x, *entry["atoms"])[0] == pytest.approx(entry["value"], rel=1e-14)


def test_torsion_sign_convention():
    """The atan2 recomputation of tests/test_cvpack.py:161-167."""
    rng = np.random.default_rng(0)
    for _ in range(20):
        x = rng.normal(size=(4, 3))
        r21, u23, r34 = x[0] - x[1], x[2] - x[1], x[3] - x[2]
        u23 = u23 / np.linalg.norm(u23)
        expected = np.arctan2(
            np.cross(r21, r34).dot(u23), r21.dot(r34) - r21.dot(u23) * r34.dot(u23)
        )
        assert o.torsion(x, 0, 1, 2, 3)[0] == pytest.approx(expected, rel=1e-13)


def test_torsion_similarity():
    """torsion_similarity.py:22-30: n/2 + 1/2 sum cos(phi1 - phi2); identical lists give n, and the
    `min(delta, 2 pi - delta)` of the reference's expression (:89) changes neither value nor slope."""
    rng = np.random.default_rng(5)
    x = rng.normal(size=(12, 3))
    first = [[0, 1, 2, 3], [2, 3, 4, 5], [6, 7, 8, 9]]
    second = [[1, 2, 3, 4], [5, 6, 7, 8], [8, 9, 10, 11]]
    assert o.torsion_similarity(x, first, first)[0] == pytest.approx(3.0, rel=1e-15)
    value, _ = o.torsion_similarity(x, first, second)
    phi1 = np.array([o.torsion(x, *t)[0] for t in first])
    phi2 = np.array([o.torsion(x, *t)[0] for t in second])
    assert value == pytest.approx(1.5 + 0.5 * np.cos(phi1 - phi2).sum(), rel=1e-14)
    for box in (None, TRICLINIC):
        fd_check(lambda y: o.torsion_similarity(y, first, second, box=box), x, range(12), tol=1e-6)


def test_path_in_cv_space_matches_logsumexp_and_finite_differences():
    """The softmin forms the reference's test-suite uses for the path CVs
    (tests/test_cvpack.py:807-860) applied to a path in (torsion, distance) space, with the
    periodic difference of path_in_cv_space.py:148-149; gradient by finite differences."""
    from scipy.special import logsumexp, softmax

    rng = np.random.default_rng(6)
    x = rng.normal(size=(6, 3))
    milestones = np.array([[3.0, 0.5], [-3.0, 1.0], [-2.0, 1.5], [0.5, 2.0]])
    sigma, periods, scales = 0.7, [2 * np.pi, 0.0], [1.0, 0.5]

    def evaluate(y, metric):
        inner = [o.torsion(y, 0, 1, 2, 3), o.distance(y, 1, 4)]
        return o.path_in_cv_space([v for v, _ in inner], [g for _, g in inner], milestones, sigma,
                                  metric, periods, scales)

    phi, dist = o.torsion(x, 0, 1, 2, 3)[0], o.distance(x, 1, 4)[0]
    dphi = np.abs(milestones[:, 0] - phi)
    dphi = np.minimum(dphi, 2 * np.pi - dphi)
    sq = dphi**2 + ((milestones[:, 1] - dist) / 0.5) ** 2
    exponents = -sq / (2 * sigma**2)
    assert evaluate(x, "progress")[0] == pytest.approx(
        np.dot(softmax(exponents), np.arange(4)) / 3, rel=1e-12)
    assert evaluate(x, "deviation")[0] == pytest.approx(-2 * sigma**2 * logsumexp(exponents), rel=1e-12)
    for metric in ("progress", "deviation"):
        fd_check(lambda y, metric=metric: evaluate(y, metric), x, range(6), tol=1e-6)


def test_radius_of_gyration_matches_reference_test():
    """tests/test_cvpack.py:172-229: about the centroid, and about the centre of mass but unweighted."""
    rng = np.random.default_rng(1)
    x = rng.normal(size=(30, 3))
    masses = rng.uniform(1, 16, 30)
    group = np.arange(30)
    centroid = x.mean(axis=0)
    com = masses @ x / masses.sum()
    assert o.radius_of_gyration(x, group)[0] == pytest.approx(
        np.sqrt(np.sum((x - centroid) ** 2) / 30)
    )
    assert o.radius_of_gyration(x, group, masses / masses.sum(), squared=True)[0] == pytest.approx(
        np.sum((x - com) ** 2) / 30
    )


def test_number_of_contacts_matches_reference_test():
    """tests/test_cvpack.py:232-272: pair set, exclusions, where(d <= 0.6, 1/(1+(d/0.3)^6), 0)."""
    rng = np.random.default_rng(2)
    x = rng.uniform(0, 1.5, size=(22, 3))
    group1, group2 = [0, 3, 4, 5, 8, 9, 12], [1, 3, 5, 6, 7, 12, 15, 20]
    exceptions = [(0, 1), (3, 5), (9, 20), (2, 11)]
    exclusions = {(i, j) if i < j else (j, i) for i, j in exceptions}
    pairs = set()
    for i, j in it.product(group1, group2):
        if j != i:
            pair = (i, j) if i < j else (j, i)
            if pair not in exclusions:
                pairs.add(pair)
    distances = np.array([np.linalg.norm(x[i] - x[j]) for i, j in pairs])
    contacts = np.where(distances <= 0.6, 1 / (1 + (distances / 0.3) ** 6), 0)
    oracle_pairs = o.interaction_pairs(group1, group2, exceptions)
    assert {tuple(p) for p in oracle_pairs.tolist()} == pairs  # integer-exact
    value, _ = o.number_of_contacts(x, oracle_pairs, switch_factor=None)
    assert value == pytest.approx(contacts.sum())


def test_attraction_strength_matches_reference_test():
    """The explicit double loop of tests/test_cvpack.py:688-732 on 20 random particles."""
    rng = np.random.default_rng(3)
    n, cutoff = 20, 1.0
    x = rng.uniform(0, 1.2, size=(n, 3))
    charge, sigma, eps = rng.normal(size=n), rng.uniform(0.1, 0.3, n), rng.uniform(0.1, 1, n)
    group1, group2 = range(0, 8), range(8, 20)
    strength = 0.0
    for i in group1:
        for j in group2:
            rij = np.linalg.norm(x[i] - x[j])
            if rij <= cutoff:
                xx = rij / cutoff
                y = abs((rij / (0.5 * (sigma[i] + sigma[j]))) ** 6 - 2) + 2
                strength += 4 * np.sqrt(eps[i] * eps[j]) * (1 / y**2 - 1 / y)
                strength += (
                    o.ONE_4PI_EPS0 * min(0.0, charge[i] * charge[j]) * (1 / xx + (xx**2 - 3) / 2)
                )
    pairs = o.interaction_pairs(group1, group2)
    value, _ = o.attraction_strength(x, pairs, charge, sigma, eps, cutoff)
    assert value == pytest.approx(-strength)


def test_residue_coordination_matches_reference_test():
    """tests/test_cvpack.py:735-770: numpy centroids, all pairs, 1/(1+d^6)."""
    rng = np.random.default_rng(4)
    x = rng.uniform(0, 3, size=(60, 3))
    g1 = [list(range(k, k + 5)) for k in range(0, 25, 5)]
    g2 = [list(range(k, k + 7)) for k in range(25, 60, 7)]
    c1 = np.array([x[g].mean(axis=0) for g in g1])
    c2 = np.array([x[g].mean(axis=0) for g in g2])
    d = np.linalg.norm(c1[:, None] - c2[None], axis=2)
    expected = np.sum(1 / (1 + d**6))
    value, _ = o.residue_coordination(
        x, g1, g2, [np.full(5, 0.2)] * 5, [np.full(7, 1 / 7)] * 5, r0=1.0
    )
    assert value == pytest.approx(expected)


def test_rmsd_matches_scipy():
    """tests/test_cvpack.py:275-338: rssd/sqrt(n) for full, half and shuffled groups."""
    rng = np.random.default_rng(5)
    n = 40
    reference = rng.normal(size=(n, 3))
    x = Rotation.random(random_state=5).apply(reference) + 0.1 * rng.normal(size=(n, 3)) + 3.0
    for group in (np.arange(n), np.arange(n // 2), rng.permutation(n)[: n // 2]):
        xs, ys = x[group] - x[group].mean(0), reference[group] - reference[group].mean(0)
        _, rssd = Rotation.align_vectors(xs, ys)
        assert o.rmsd(x, reference[group], group)[0] == pytest.approx(rssd / np.sqrt(len(group)))


def test_composite_rmsd_invariance():
    """composite_rmsd.py:116-121: 0.0 at the reference and after translating one body by 1 nm."""
    rng = np.random.default_rng(6)
    reference = rng.normal(size=(30, 3))
    bodies = [list(range(0, 12)), list(range(15, 30))]
    ref_rows = reference[sum(bodies, [])]
    assert o.composite_rmsd(reference, ref_rows, bodies)[0] == GOLDEN["composite_rmsd"]["value"]
    moved = reference.copy()
    moved[bodies[1]] += 1.0
    assert o.composite_rmsd(moved, ref_rows, bodies)[0] == pytest.approx(0.0, abs=1e-7)
    # one body == plain RMSD
    x = reference + 0.1 * rng.normal(size=reference.shape)
    assert o.composite_rmsd(x, reference[bodies[0]], [bodies[0]])[0] == pytest.approx(
        o.rmsd(x, reference[bodies[0]], bodies[0])[0]
    )


def test_path_matches_logsumexp():
    """tests/test_cvpack.py:807-860: logsumexp forms of progress and deviation."""
    rng = np.random.default_rng(7)
    n, sigma = 25, 0.2
    atoms = list(range(n))
    refs = [rng.normal(size=(n, 3)) * 0.3 + k * 0.05 for k in range(7)]
    x = refs[3] + 0.05 * rng.normal(size=(n, 3))
    milestones = [(atoms, r) for r in refs]
    sq = np.array([o.rmsd(x, r, atoms)[0] ** 2 for r in refs])
    exponents = -sq / (2 * sigma**2)
    expected_s = np.exp(logsumexp(exponents, b=np.arange(7)) - logsumexp(exponents)) / 6
    expected_z = -2 * sigma**2 * logsumexp(exponents)
    assert o.path_in_rmsd_space(x, milestones, sigma, "progress")[0] == pytest.approx(expected_s)
    assert o.path_in_rmsd_space(x, milestones, sigma, "deviation")[0] == pytest.approx(expected_z)


def test_shortest_distance_matches_logsumexp():
    """tests/test_cvpack.py:890-928: softmin including the r_c guard term."""
    rng = np.random.default_rng(8)
    x = rng.uniform(0, 1.5, size=(30, 3))
    g1, g2 = range(0, 12), range(12, 30)
    beta, rc = 100.0, 1.0
    r = np.array([np.linalg.norm(x[i] - x[j]) for i in g1 for j in g2])
    r = r[r < rc]
    exponents = np.append(beta * (1 - r / rc), 0.0)
    expected = np.exp(logsumexp(exponents, b=np.append(r, rc)) - logsumexp(exponents))
    value, _ = o.shortest_distance(x, o.interaction_pairs(g1, g2), beta, rc)
    assert value == pytest.approx(expected)


def test_effective_mass():
    """utils.py:172-184: restricted to atoms with non-zero gradient; inf if none."""
    grad = np.zeros((5, 3))
    masses = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    assert o.effective_mass(grad, masses) == np.inf
    grad[1] = [1, 0, 0]
    grad[3] = [0, 2, 0]
    assert o.effective_mass(grad, masses) == pytest.approx(1 / (1 / 2 + 4 / 4))


@pytest.mark.parametrize("box", [None, np.diag([2.0, 2.2, 2.4]), TRICLINIC])
def test_gradients_by_finite_differences(box):
    rng = np.random.default_rng(9)
    x = rng.uniform(0, 2.0, size=(40, 3))
    atoms = range(40)
    fd_check(lambda y: o.distance(y, 3, 17, box=box), x, atoms)
    fd_check(lambda y: o.angle(y, 3, 17, 5, box=box), x, atoms)
    fd_check(lambda y: o.torsion(y, 3, 17, 5, 9, box=box), x, atoms)
    w = rng.uniform(1, 16, 20)
    w /= w.sum()
    fd_check(lambda y: o.radius_of_gyration(y, list(range(5, 25)), box=box), x, atoms)
    fd_check(lambda y: o.radius_of_gyration(y, list(range(5, 25)), w, box, squared=True), x, atoms)
    pairs = o.interaction_pairs(range(0, 25), range(15, 40), [(0, 16), (20, 21), (17, 3)])
    fd_check(lambda y: o.number_of_contacts(y, pairs, box=box), x, atoms)
    q, s, e = rng.normal(size=40), rng.uniform(0.2, 0.4, 40), rng.uniform(0.1, 1, 40)
    sign = np.where(np.arange(40) > 30, -1.0, 1.0)
    fd_check(
        lambda y: o.attraction_strength(y, pairs, q, s, e, 0.9, sign, 0.7, box=box), x, atoms, 1e-5
    )
    fd_check(lambda y: o.shortest_distance(y, pairs, 100, 0.95, box), x, atoms, 1e-5)
    g1 = [list(range(k, k + 4)) for k in range(0, 16, 4)]
    g2 = [list(range(k, k + 3)) for k in range(20, 38, 3)]
    fd_check(
        lambda y: o.residue_coordination(
            y, g1, g2, [np.full(4, 0.25)] * 4, [np.array([0.5, 0.3, 0.2])] * 6, r0=0.8, box=box
        ),
        x,
        atoms,
    )
    terms = [[1, 2, 3, 4], [5, 6, 7, 8]]
    fd_check(lambda y: o.bonded_sum(y, terms, "boxcar", [-1.1, -0.7], 0.44, 3, 0.5, box), x, atoms)


def test_rmsd_family_gradients_by_finite_differences():
    rng = np.random.default_rng(10)
    x = rng.normal(size=(40, 3))
    atoms = range(40)
    ref = rng.normal(size=(20, 3))
    fd_check(lambda y: o.rmsd(y, ref, list(range(10, 30))), x, atoms)
    fd_check(lambda y: o.composite_rmsd(y, ref, [list(range(0, 8)), list(range(20, 32))]), x, atoms)
    ideal = rng.normal(size=(10, 3))
    blocks = [list(range(k, k + 10)) for k in range(0, 30, 3)]
    for step in ("rational", "rational_pair", "plumed"):
        fd_check(lambda y: o.rmsd_content(y, blocks, ideal, r0=1.0, step=step, n=4, m=8), x, atoms, 1e-5)
    milestones = [(list(range(5, 30)), x[5:30] + 0.1 * k * rng.normal(size=(25, 3))) for k in range(1, 6)]
    for metric in ("progress", "deviation"):
        fd_check(lambda y: o.path_in_rmsd_space(y, milestones, 0.1, metric), x, atoms, 1e-5)


def test_c_oracle_matches_numpy():
    rng = np.random.default_rng(11)
    x = rng.uniform(0, 2, size=(3, 200, 3))
    g1, g2 = list(range(0, 120)), list(range(80, 200))
    ex = [(0, 100), (90, 95), (5, 150)]
    for box in (None, np.diag([2.0, 2.1, 2.2]), TRICLINIC):
        value, grad = c_oracle.contacts_batch(x, g1, g2, ex, box)
        for b in range(3):
            ref_value, ref_grad = o.number_of_contacts(x[b], o.interaction_pairs(g1, g2, ex), box=box)
            assert value[b] == pytest.approx(ref_value, rel=1e-12)
            assert np.abs(grad[b] - ref_grad).max() < 1e-12
    group = rng.permutation(200)[:50]
    ref = rng.normal(size=(50, 3))
    value, grad = c_oracle.rmsd_batch(x, group, ref)
    weights = rng.uniform(1, 2, 50)
    weights /= weights.sum()
    gvalue, ggrad = c_oracle.gyration_batch(x, group, weights)
    for b in range(3):
        ref_value, ref_grad = o.rmsd(x[b], ref, group)
        assert value[b] == pytest.approx(ref_value, rel=1e-12)
        assert np.abs(grad[b] - ref_grad).max() < 1e-12
        ref_value, ref_grad = o.radius_of_gyration(x[b], group, weights)
        assert gvalue[b] == pytest.approx(ref_value, rel=1e-12)
        assert np.abs(ggrad[b] - ref_grad).max() < 1e-12


def test_c_oracle_cell_list_matches_all_pairs():
    """The cell-list CPU arm of bench.py (neighbour-list analogue of OpenMM's CPU platform) visits the
    same pairs as the all-pairs Reference-platform loop: same values, gradients to summation order."""
    rng = np.random.default_rng(12)
    n = 900
    x = rng.uniform(-0.3, 2.9, size=(4, n, 3))  # some atoms outside the box: wrapped
    g1, g2 = list(range(0, 500)), list(range(400, 900))
    ex = [(0, 450), (410, 420), (5, 700), (499, 500)]
    boxes = [None, np.diag([2.6, 2.7, 2.8]), np.stack([np.diag([2.6 + 0.1 * k] * 3) for k in range(4)])]
    for box in boxes:
        for rs in (0.6, None):
            ref_value, ref_grad = c_oracle.contacts_batch(x, g1, g2, ex, box, r0=0.4, rc=0.8, rs=rs)
            value, grad = c_oracle.contacts_batch(x, g1, g2, ex, box, r0=0.4, rc=0.8, rs=rs, cells=True)
            assert np.allclose(value, ref_value, rtol=1e-13, atol=0)
            assert np.abs(grad - ref_grad).max() <= 1e-12 * np.abs(ref_grad).max()
            assert np.array_equal(grad != 0, ref_grad != 0)
    # group 1 far from the bounding box of group 2 (no box): border cells still reached
    y = x.copy()
    y[:, :400] += np.array([3.3, 0.0, 0.0])
    ref_value, ref_grad = c_oracle.contacts_batch(y, g1, g2, ex, None, r0=0.4, rc=0.8)
    value, grad = c_oracle.contacts_batch(y, g1, g2, ex, None, r0=0.4, rc=0.8, cells=True)
    assert np.allclose(value, ref_value, rtol=1e-13) and np.abs(grad - ref_grad).max() <= 1e-12
    with pytest.raises(ValueError):
        c_oracle.contacts_batch(x, g1, g2, ex, TRICLINIC, cells=True)
    with pytest.raises(ValueError):
        c_oracle.contacts_batch(x, g1, g2, ex, np.diag([2.0, 2.0, 2.0]), r0=0.4, rc=0.8, cells=True)
