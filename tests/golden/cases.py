"""
The case table behind ``tests/golden/reference_vectors.*``.

Every case builds the *same* collective variable on the *same* seeded synthetic system through a
namespace ``ns`` -- either the reference package running over the OpenMM stand-in
(``oracle.refshim``; used once, by ``make_reference_golden.py``, to produce the golden vectors) or
``cvpack_b200`` (used by the tests to compare the CPU oracle and the CUDA path with them).  The
constructors are called with identical arguments in both worlds: that is the drop-in claim.

Coordinates, box vectors and reference structures are rounded to FP32-representable values so that
the FP32 and FP64 modes of the GPU path and the FP64 generators all see the same numbers.
"""

from __future__ import annotations

import typing as t

import numpy as np

BACKBONE = ("N", "H", "CA", "CB", "C", "O")
ORTHO = np.diag([2.5, 2.75, 3.0])
TRICLINIC = np.array([[2.5, 0, 0], [0.25, 2.75, 0], [-0.5, 0.375, 3.0]])
BOXES = {"none": None, "ortho": ORTHO, "tric": TRICLINIC}


def f32(array: t.Any) -> np.ndarray:
    """Round to FP32-representable doubles."""
    return np.asarray(array, dtype=np.float64).astype(np.float32).astype(np.float64)


class Built(t.NamedTuple):
    """What a case builder returns."""

    cv: t.Any
    system: t.Any
    positions: np.ndarray  # [F, N, 3]
    box: t.Optional[np.ndarray]  # [3, 3] or None


class Case(t.NamedTuple):
    """A named builder; ``needs_context`` cases take a context of frame 0 as ``reference``."""

    name: str
    build: t.Callable[..., Built]
    row: str  # SURVEY.md section 8 row the case pins
    needs_context: bool = False
    fp32_factor: float = 1.0  # see CONDITIONING below


# ---- ns-generic builders ---------------------------------------------------------------------
def make_system(ns, num_atoms: int, rng: np.random.Generator, box=None):
    system = ns.System()
    for mass in rng.uniform(1.0, 16.0, num_atoms):
        system.addParticle(float(mass))
    if box is not None:
        system.setDefaultPeriodicBoxVectors(*box)
    return system


def make_nonbonded(ns, num_atoms, rng, periodic, cutoff=0.9, switch=None, num_exceptions=0):
    force = ns.NonbondedForce()
    for _ in range(num_atoms):
        force.addParticle(float(rng.normal() * 0.5), float(rng.uniform(0.2, 0.4)), float(rng.uniform(0.1, 1.0)))
    force.setNonbondedMethod(force.CutoffPeriodic if periodic else force.CutoffNonPeriodic)
    force.setCutoffDistance(cutoff)
    if switch is not None:
        force.setUseSwitchingFunction(True)
        force.setSwitchingDistance(switch)
    seen: t.Set[t.Tuple[int, int]] = set()
    while len(seen) < num_exceptions:
        i, j = (int(v) for v in rng.integers(0, num_atoms, 2))
        if i != j and (min(i, j), max(i, j)) not in seen:
            seen.add((min(i, j), max(i, j)))
            force.addException(i, j, 0.0, 1.0, 0.0)
    return force


def make_peptide(ns, num_residues: int, glycines: t.Sequence[int] = ()):
    topology = ns.app.Topology()
    chain = topology.addChain()
    elements = {"N": ns.app.nitrogen, "H": ns.app.hydrogen, "CA": ns.app.carbon, "CB": ns.app.carbon,
                "C": ns.app.carbon, "O": ns.app.oxygen}
    for index in range(num_residues):
        gly = index in glycines
        residue = topology.addResidue("GLY" if gly else "ALA", chain)
        for name in BACKBONE:
            element = elements[name]
            if gly and name == "CB":
                name, element = "HA2", ns.app.hydrogen
            topology.addAtom(name, element, residue)
    return topology


def uniform_frames(rng, num_frames, num_atoms, scale=2.4) -> np.ndarray:
    return f32(rng.uniform(0.0, scale, size=(num_frames, num_atoms, 3)))


def helix_frames(rng, num_residues, num_frames, noise) -> np.ndarray:
    from tests import helpers  # pylint: disable=import-outside-toplevel

    base = helpers.ideal_helix_positions(num_residues, rng, 0.0)
    return f32(np.stack([base + noise * rng.normal(size=base.shape) for _ in range(num_frames)]))


# ---- cases -------------------------------------------------------------------------------------
def _bonded(kind: str, boxname: str):
    def build(ns, context=None) -> Built:
        rng = np.random.default_rng(101)
        box = BOXES[boxname]
        system = make_system(ns, 24, rng, box)
        x = uniform_frames(rng, 4, 24)
        pbc = box is not None
        cv = {
            "distance": lambda: ns.cvpack.Distance(3, 17, pbc=pbc),
            "angle": lambda: ns.cvpack.Angle(3, 17, 5, pbc=pbc),
            "torsion": lambda: ns.cvpack.Torsion(3, 17, 5, 9, pbc=pbc),
        }[kind]()
        return Built(cv, system, x, box)

    return build


def _gyration(squared: bool, pbc: bool, weigh: bool):
    def build(ns, context=None) -> Built:
        rng = np.random.default_rng(102)
        box = ORTHO if pbc else None
        system = make_system(ns, 60, rng, box)
        x = uniform_frames(rng, 3, 60, scale=1.1 if pbc else 2.4)
        # RadiusOfGyrationSq in the reference indexes the *groups* by atom index
        # (radius_of_gyration_sq.py:88-89): it is only meaningful for group == range(n)
        group = list(range(40)) if squared else [int(i) for i in rng.permutation(60)[:40]]
        cls = ns.cvpack.RadiusOfGyrationSq if squared else ns.cvpack.RadiusOfGyration
        return Built(cls(group, pbc=pbc, weighByMass=weigh), system, x, box)

    return build


def _contacts(variant: str):
    def build(ns, context=None) -> Built:
        rng = np.random.default_rng(103)
        n = 160
        boxname = {"tric": "tric", "nopbc": "none"}.get(variant, "ortho")
        box = BOXES[boxname]
        system = make_system(ns, n, rng, box)
        x = uniform_frames(rng, 3, n, scale=2.4)
        nb = make_nonbonded(ns, n, rng, periodic=box is not None,
                            num_exceptions=40 if variant in ("exclusions", "overlap") else 0)
        g1, g2 = list(range(0, 80)), list(range(80, 160))
        kwargs: t.Dict[str, t.Any] = {}
        if variant == "overlap":
            g1, g2 = list(range(0, 100)), list(range(60, 160))
        elif variant == "same":
            g1 = g2 = list(range(20, 120))
        elif variant == "sharp":
            kwargs = {"switchFactor": None, "thresholdDistance": 0.35 * ns.unit.nanometers}
        elif variant == "step12":
            kwargs = {"stepFunction": "(1-x^6)/(1-x^12)", "thresholdDistance": 0.4, "cutoffFactor": 2.5,
                      "switchFactor": 2.0}
        elif variant == "heaviside":
            kwargs = {"stepFunction": "step(1-x)", "thresholdDistance": 0.5, "switchFactor": None}
        elif variant == "normalised":
            kwargs = {"reference": 37.5}
        elif variant == "context":
            kwargs = {"reference": context}
        if variant in ("ortho", "tric", "nopbc"):
            kwargs = {"thresholdDistance": 0.45 * ns.unit.nanometers}
        cv = ns.cvpack.NumberOfContacts(g1, g2, nb, **kwargs)
        return Built(cv, system, x, box)

    return build


def _residue_coordination(pbc: bool, weigh: bool, hydrogens: bool, normalize: bool):
    def build(ns, context=None) -> Built:
        rng = np.random.default_rng(104)
        topology = make_peptide(ns, 14)
        n = topology.getNumAtoms()
        box = ORTHO if pbc else None
        system = make_system(ns, n, rng, box)
        x = uniform_frames(rng, 3, n)
        residues = list(topology.residues())
        cv = ns.cvpack.ResidueCoordination(
            residues[:6], residues[7:], pbc=pbc, thresholdDistance=0.9 * ns.unit.nanometers,
            normalize=normalize, weighByMass=weigh, includeHydrogens=hydrogens,
        )
        return Built(cv, system, x, box)

    return build


def _attraction(variant: str):
    def build(ns, context=None) -> Built:
        rng = np.random.default_rng(105)
        n = 150
        periodic = variant != "nopbc"
        box = ORTHO if periodic else None
        system = make_system(ns, n, rng, box)
        x = uniform_frames(rng, 3, n)
        nb = make_nonbonded(ns, n, rng, periodic=periodic, cutoff=1.0,
                            switch=0.8 if variant in ("switch", "contrast") else None,
                            num_exceptions=30 if variant != "plain" else 0)
        g1, g2 = list(range(0, 50)), list(range(50, 100))
        kwargs: t.Dict[str, t.Any] = {}
        if variant == "contrast":
            kwargs = {"contrastGroup": list(range(100, 150)), "contrastScaling": 0.75,
                      "reference": 2.5 * ns.unit.kilojoule_per_mole}
        elif variant == "self":
            g2 = g1
        elif variant == "context":
            kwargs = {"reference": context, "contrastGroup": list(range(100, 150))}
        return Built(ns.cvpack.AttractionStrength(g1, g2, nb, **kwargs), system, x, box)

    return build


def _shortest_distance(variant: str):
    def build(ns, context=None) -> Built:
        rng = np.random.default_rng(106)
        n = 120
        pbc = variant != "nopbc"
        box = ORTHO if pbc else None
        system = make_system(ns, n, rng, box)
        x = uniform_frames(rng, 3, n, scale=2.4)
        if variant == "far":
            # the two groups are kept 0.9-1.0 rc apart: exercises the e^{-beta} regime
            x[:, :60, 0] = f32(0.2 + 0.1 * rng.uniform(size=(3, 60)))
            x[:, 60:, 0] = f32(1.26 + 0.05 * rng.uniform(size=(3, 60)))
        cv = ns.cvpack.ShortestDistance(list(range(0, 60)), list(range(60, 120)), n, pbc=pbc)
        return Built(cv, system, x, box)

    return build


def _rmsd(variant: str):
    def build(ns, context=None) -> Built:
        rng = np.random.default_rng(107)
        n = 90
        system = make_system(ns, n, rng)
        ref = f32(rng.normal(size=(n, 3)))
        x = []
        for _ in range(3):
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            a, b, c, d = q
            rot = np.array([[a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c)],
                            [2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b)],
                            [2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d]])
            x.append(ref @ rot.T + rng.normal(size=3) + 0.1 * rng.normal(size=(n, 3)))
        x = f32(np.stack(x))
        if variant == "all":
            cv = ns.cvpack.RMSD(ref, range(n), n)
        elif variant == "shuffled":
            group = [int(i) for i in rng.permutation(n)[:50]]
            cv = ns.cvpack.RMSD({i: ref[i] for i in group}, group, n)
        elif variant == "composite2":
            cv = ns.cvpack.CompositeRMSD(ref, [list(range(0, 30)), list(range(40, 75))], n)
            x[:, 40:75] += f32(rng.normal(size=(3, 1, 3)))  # bodies drift apart: invisible to the CV
        else:
            cv = ns.cvpack.CompositeRMSD(ref, [[3, 1, 4, 15, 9, 2, 6], list(range(20, 45)), list(range(50, 90))], n)
        return Built(cv, system, f32(x), None)

    return build


def _helix(kind: str, pbc: bool, normalize: bool):
    def build(ns, context=None) -> Built:
        rng = np.random.default_rng(108)
        nres = 12
        topology = make_peptide(ns, nres, glycines=(4,))
        n = topology.getNumAtoms()
        box = np.diag([3.0, 3.25, 3.5]) if pbc else None
        system = make_system(ns, n, rng, box)
        x = helix_frames(rng, nres, 3, 0.03)
        residues = list(topology.residues())
        cls = {"torsion": ns.cvpack.HelixTorsionContent, "angle": ns.cvpack.HelixAngleContent,
               "hbond": ns.cvpack.HelixHBondContent}[kind]
        return Built(cls(residues, pbc=pbc, normalize=normalize), system, x, box)

    return build


def _rmsd_content(kind: str, normalize: bool):
    def build(ns, context=None) -> Built:
        rng = np.random.default_rng(109)
        nres = {"helix": 12, "helix_chunks": 40, "sheet_anti": 14, "sheet_para": 15, "sheet_blocks": 16}[kind]
        topology = make_peptide(ns, nres, glycines=(2, 7))
        n = topology.getNumAtoms()
        system = make_system(ns, n, rng)
        x = helix_frames(rng, nres, 3, 0.02)
        residues = list(topology.residues())
        if kind.startswith("helix"):
            cv = ns.cvpack.HelixRMSDContent(residues, n, normalize=normalize)
        elif kind == "sheet_blocks":
            cv = ns.cvpack.SheetRMSDContent(residues, n, blockSizes=[5, 6, 5], normalize=normalize,
                                            thresholdRMSD=0.3 * ns.unit.nanometers)
        else:
            cv = ns.cvpack.SheetRMSDContent(residues, n, parallel=kind == "sheet_para", normalize=normalize,
                                            thresholdRMSD=0.3 * ns.unit.nanometers)
        return Built(cv, system, x, None)

    return build


def _path_rmsd(metric: str, variant: str):
    def build(ns, context=None) -> Built:
        rng = np.random.default_rng(110)
        n = 70
        system = make_system(ns, n, rng)
        a, b = rng.normal(size=(n, 3)), rng.normal(size=(n, 3))
        count = 6
        atoms = list(range(5, 55))
        milestones = []
        for k in range(count):
            s = k / (count - 1)
            structure = f32((1 - s) * a + s * b + 0.02 * rng.normal(size=(n, 3)))
            members = atoms if variant == "same" else [int(i) for i in rng.permutation(n)[: 30 + 3 * k]]
            milestones.append({i: structure[i] for i in members})
        x = f32(np.stack([(1 - s) * a + s * b + 0.05 * rng.normal(size=(n, 3)) for s in (0.1, 0.45, 0.8)]))
        cv = ns.cvpack.PathInRMSDSpace(getattr(ns.cvpack.path, metric), milestones, n, 0.5 * ns.unit.nanometers)
        return Built(cv, system, x, None)

    return build


def _path_cv(metric: str):
    def build(ns, context=None) -> Built:
        rng = np.random.default_rng(111)
        system = make_system(ns, 30, rng)
        x = uniform_frames(rng, 4, 30, scale=1.5)
        variables = [ns.cvpack.Torsion(3, 17, 5, 9), ns.cvpack.Distance(2, 11),
                     ns.cvpack.RadiusOfGyration(range(8, 20))]
        milestones = np.array([[3.0, 0.4, 0.5], [-3.0, 0.8, 0.6], [-2.0, 1.2, 0.7], [0.5, 1.6, 0.8], [1.5, 2.0, 0.9]])
        cv = ns.cvpack.PathInCVSpace(getattr(ns.cvpack.path, metric), variables, milestones, 0.8,
                                     scales=[1.0, 0.5, 0.2])
        return Built(cv, system, x, None)

    return build


def _torsion_similarity(boxname: str):
    def build(ns, context=None) -> Built:
        rng = np.random.default_rng(112)
        box = BOXES[boxname]
        system = make_system(ns, 40, rng, box)
        x = uniform_frames(rng, 3, 40)
        first = [[int(i) for i in rng.choice(40, size=4, replace=False)] for _ in range(9)]
        second = [[int(i) for i in rng.choice(40, size=4, replace=False)] for _ in range(9)]
        second[0] = first[0]
        return Built(ns.cvpack.TorsionSimilarity(first, second, pbc=box is not None), system, x, box)

    return build


def _meta(variant: str):
    def build(ns, context=None) -> Built:
        rng = np.random.default_rng(113)
        system = make_system(ns, 30, rng)
        x = uniform_frames(rng, 4, 30, scale=1.5)
        phi = ns.cvpack.Torsion(3, 17, 5, 9, name="phi")
        dist = ns.cvpack.Distance(2, 11, name="dist")
        rg = ns.cvpack.RadiusOfGyration(range(8, 20), name="rg")
        unit = ns.unit
        if variant == "harmonic":
            cv = ns.cvpack.MetaCollectiveVariable(
                "0.5*kappa*min(delta,6.283185307179586-delta)^2; delta=abs(phi-phi0)", [phi],
                unit.kilojoules_per_mole, kappa=1000 * unit.kilojoules_per_mole / unit.radian**2,
                phi0=120 * unit.degrees,
            )
        elif variant == "ratio":
            cv = ns.cvpack.MetaCollectiveVariable("dist/(rg+shift)", [dist, rg], unit.dimensionless,
                                                  shift=0.1 * unit.nanometers)
        else:
            cv = ns.cvpack.MetaCollectiveVariable(
                "a*sin(phi)+dist^2*exp(-rg/b)", [phi, dist, rg], unit.nanometers**2,
                a=0.3 * unit.nanometers**2, b=2.0 * unit.angstroms,
            )
        return Built(cv, system, x, None)

    return build


# CONDITIONING (why three families carry a factor on the FP32 tolerance; the FP64 mode is held to
# the plain 1e-10): written next to each entry.
CASES: t.List[Case] = (
    [Case(f"{kind}_{box}", _bonded(kind, box), "a11") for kind in ("distance", "angle", "torsion")
     for box in ("none", "ortho", "tric")]
    + [Case(f"rg{'sq' if sq else ''}_{'pbc' if pbc else 'nopbc'}_{'mass' if w else 'geom'}",
            _gyration(sq, pbc, w), "a10")
       for sq in (False, True) for pbc in (False, True) for w in (False, True)]
    + [Case(f"contacts_{v}", _contacts(v), "a4", needs_context=v == "context")
       for v in ("ortho", "tric", "nopbc", "exclusions", "overlap", "same", "sharp", "step12",
                 "heaviside", "normalised", "context")]
    + [Case(f"rescoord_{'pbc' if p else 'nopbc'}_{'mass' if w else 'geom'}_{'allH' if h else 'noH'}"
            f"{'_norm' if nz else ''}", _residue_coordination(p, w, h, nz), "a5")
       for p, w, h, nz in ((True, True, True, False), (False, False, True, True), (True, True, False, False),
                           (True, False, False, True))]
    + [Case(f"attraction_{v}", _attraction(v), "a6", needs_context=v == "context")
       for v in ("plain", "nopbc", "switch", "contrast", "self", "context")]
    + [Case(f"shortest_{v}", _shortest_distance(v), "a7") for v in ("pbc", "nopbc", "far")]
    + [Case(f"rmsd_{v}", _rmsd(v), "a8") for v in ("all", "shuffled")]
    + [Case(f"rmsd_{v}", _rmsd(v), "a9") for v in ("composite2", "composite3")]
    + [Case(f"helix_{k}_{'pbc' if p else 'nopbc'}{'_norm' if nz else ''}", _helix(k, p, nz), "a12")
       for k in ("torsion", "angle", "hbond") for p, nz in ((False, False), (True, True))]
    + [Case(f"content_{k}{'_norm' if nz else ''}", _rmsd_content(k, nz), "a13")
       for k, nz in (("helix", False), ("helix", True), ("helix_chunks", False), ("sheet_anti", False),
                     ("sheet_para", True), ("sheet_blocks", False))]
    + [Case(f"path_rmsd_{m}_{v}", _path_rmsd(m, v), "a14")
       for m in ("progress", "deviation") for v in ("same", "ragged")]
    + [Case(f"path_cv_{m}", _path_cv(m), "f2") for m in ("progress", "deviation")]
    + [Case(f"torsion_similarity_{b}", _torsion_similarity(b), "f1") for b in ("none", "ortho", "tric")]
    + [Case(f"meta_{v}", _meta(v), "f4") for v in ("harmonic", "ratio", "mixed")]
)

CASES_BY_NAME = {case.name: case for case in CASES}
