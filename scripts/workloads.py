"""
Synthetic workloads of BASELINE.json's configs, shared by bench.py and the scripts.

Every builder returns ``(cv(s), system, positions [B, N, 3] float32 on the device, box or None,
info)``; the coordinates are generated on the GPU (torch generators, fixed seeds) so that a bench
run does not spend its time in NumPy.  Nothing here touches ``oracle/``.
"""

from __future__ import annotations

import typing as t

import numpy as np
import torch

import cvpack_b200 as cvpack

CONFIGS = {
    "contacts": "NumberOfContacts 10k x 10k atoms, cubic PBC L=5.844 nm, r0=0.6 rc=1.2 rs=0.9 nm, "
    "4096 frames/GPU of water density (100 atoms/nm^3)",
    "residue_coordination": "ResidueCoordination between two groups of 1000 residues x 10 atoms (the 10k + 10k "
    "atoms of the contact workload), r0=1.0 nm, PBC, weighByMass=False, same frames",
    "rmsd": "RMSD (quaternion alignment) of a 100k-atom synthetic protein, 2048 frames",
    "rg": "RadiusOfGyration of a 100k-atom synthetic protein, 2048 frames",
    "path": "PathInRMSDSpace (progress), 64 milestones x 20k atoms, sigma 0.05 nm, 1024 frames",
    "helix_rmsd": "HelixRMSDContent, 495 blocks of a 500-residue helix, 1024 frames",
    "helix_torsion": "HelixTorsionContent, 996 torsions of a 500-residue helix, 1024 frames",
    "walkers": "NumberOfContacts 2.5k x 2.5k (PBC L=6.2 nm) + CompositeRMSD 1.5k + 1k atoms on "
    "23k-atom walkers",
}


def unit_system(num_atoms: int, box: t.Optional[float] = None) -> cvpack.System:
    """A system of unit-mass particles."""
    system = cvpack.System()
    for _ in range(num_atoms):
        system.addParticle(1.0)
    if box is not None:
        system.setDefaultPeriodicBoxVectors(*np.diag([box] * 3))
    return system


def lattice_frames(num_frames: int, num_atoms: int, box: float, seed: int, device) -> t.Tuple[torch.Tensor, torch.Tensor]:
    """Jittered cubic lattice at the density num_atoms / box^3, independent jitter per frame."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    m = int(round(num_atoms ** (1.0 / 3.0)))
    while m**3 < num_atoms:
        m += 1
    cell = box / m
    grid = torch.stack(
        torch.meshgrid(*[torch.arange(m, device=device, dtype=torch.float32)] * 3, indexing="ij"), -1
    ).reshape(-1, 3)
    perm = torch.randperm(m**3, generator=gen, device=device)[:num_atoms]
    lattice = (grid[perm] + 0.5) * cell
    x = lattice.unsqueeze(0) + 0.3 * cell * torch.randn(
        (num_frames, num_atoms, 3), generator=gen, device=device, dtype=torch.float32
    )
    return x.contiguous(), lattice


def blob_frames(num_frames: int, num_atoms: int, seed: int, device) -> t.Tuple[torch.Tensor, torch.Tensor]:
    """Protein-like blob (sigma 3 nm): reference + 0.05 nm noise, random rigid rotation + translation."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    reference = 3.0 * torch.randn((num_atoms, 3), generator=gen, device=device, dtype=torch.float32)
    q = torch.randn((num_frames, 4), generator=gen, device=device, dtype=torch.float32)
    q = q / q.norm(dim=1, keepdim=True)
    a, b, c, d = q.unbind(1)
    rot = torch.stack(
        [
            a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c),
            2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b),
            2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d,
        ],
        dim=1,
    ).reshape(num_frames, 3, 3)
    x = torch.empty((num_frames, num_atoms, 3), device=device, dtype=torch.float32)
    chunk = 64
    for first in range(0, num_frames, chunk):
        last = min(first + chunk, num_frames)
        noisy = reference.unsqueeze(0) + 0.05 * torch.randn(
            (last - first, num_atoms, 3), generator=gen, device=device, dtype=torch.float32
        )
        x[first:last] = torch.einsum("bij,bnj->bni", rot[first:last], noisy) + 2.0 * torch.randn(
            (last - first, 1, 3), generator=gen, device=device, dtype=torch.float32
        )
    return x.contiguous(), reference


def nonbonded(num_atoms: int, periodic: bool = True) -> cvpack.NonbondedForce:
    force = cvpack.NonbondedForce()
    for _ in range(num_atoms):
        force.addParticle(0.0, 0.3, 0.0)
    force.setNonbondedMethod(force.CutoffPeriodic if periodic else force.CutoffNonPeriodic)
    return force


def contacts(num_frames: int, seed: int, device, num_atoms: int = 20000, box: float = 5.844, r0: float = 0.6):
    """configs[1]: NumberOfContacts between the two halves of a water-density lattice."""
    x, lattice = lattice_frames(num_frames, num_atoms, box, seed, device)
    system = unit_system(num_atoms)
    cv = cvpack.NumberOfContacts(
        range(0, num_atoms // 2), range(num_atoms // 2, num_atoms), nonbonded(num_atoms),
        thresholdDistance=r0, cutoffFactor=2.0, switchFactor=1.5,
    )
    cv.addToSystem(system)
    return cv, system, x, [box] * 3, {"lattice": lattice, "r0": r0, "box": box}


def residue_coordination(num_frames: int, seed: int, device, num_atoms: int = 20000, box: float = 5.844,
                         residue_size: int = 10):
    """configs[1], SURVEY 8(d) variant: every 10 consecutive atoms are a residue; coordination between
    the residues of the first and of the second half (1000 x 1000 centroid pairs, no cutoff)."""
    from cvpack_b200 import app

    x, lattice = lattice_frames(num_frames, num_atoms, box, seed, device)
    system = unit_system(num_atoms)
    topology = app.Topology()
    chain = topology.addChain()
    residues = []
    for _ in range(num_atoms // residue_size):
        residue = topology.addResidue("ALA", chain)
        for k in range(residue_size):
            topology.addAtom(f"C{k}", app.Element.getBySymbol("C"), residue)
        residues.append(residue)
    half = len(residues) // 2
    cv = cvpack.ResidueCoordination(residues[:half], residues[half:], pbc=True, thresholdDistance=1.0,
                                    weighByMass=False, includeHydrogens=True)
    cv.addToSystem(system)
    return cv, system, x, [box] * 3, {"lattice": lattice, "residues": half}


def rmsd(num_frames: int, seed: int, device, num_atoms: int = 100000):
    """configs[2]: RMSD of all atoms of a 100k-atom blob."""
    x, reference = blob_frames(num_frames, num_atoms, seed, device)
    system = unit_system(num_atoms)
    cv = cvpack.RMSD(reference.cpu().numpy().astype(np.float64), range(num_atoms), num_atoms)
    cv.addToSystem(system)
    return cv, system, x, None, {"reference": reference}


def gyration(num_frames: int, seed: int, device, num_atoms: int = 100000):
    """configs[2]: radius of gyration of all atoms."""
    x, reference = blob_frames(num_frames, num_atoms, seed, device)
    system = unit_system(num_atoms)
    cv = cvpack.RadiusOfGyration(range(num_atoms))
    cv.addToSystem(system)
    return cv, system, x, None, {"reference": reference}


def path_rmsd(num_frames: int, seed: int, device, num_atoms: int = 20000, milestones: int = 64,
              sigma: float = 0.05):
    """configs[3]: PathInRMSDSpace; frames are generated at known positions 20..40 along the path."""
    rng = np.random.default_rng(seed)
    a, b = 2.0 * rng.normal(size=(num_atoms, 3)), 2.0 * rng.normal(size=(num_atoms, 3))
    b = a + 0.3 * (b - a) / np.linalg.norm(b - a, axis=1, keepdims=True)
    structures = [
        a + (b - a) * k / (milestones - 1) + 0.02 * rng.normal(size=(num_atoms, 3)) for k in range(milestones)
    ]
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 1)
    t_gen = 20 + 20 * torch.rand(num_frames, generator=gen, device=device)
    ta = torch.tensor(a, device=device, dtype=torch.float32)
    tb = torch.tensor(b, device=device, dtype=torch.float32)
    x = ta[None] + (tb - ta)[None] * (t_gen / (milestones - 1))[:, None, None]
    x = (x + 0.02 * torch.randn(x.shape, generator=gen, device=device)).contiguous()
    system = unit_system(num_atoms)
    refs = [dict(enumerate(structure)) for structure in structures]
    cv = cvpack.PathInRMSDSpace(cvpack.path.progress, refs, num_atoms, sigma)
    cv.addToSystem(system)
    return cv, system, x, None, {"generated_position": t_gen, "structures": structures, "sigma": sigma}


# ---- helix ------------------------------------------------------------------------------------------
_BACKBONE = (("N", "nitrogen"), ("H", "hydrogen"), ("CA", "carbon"), ("CB", "carbon"), ("C", "carbon"),
             ("O", "oxygen"))


def peptide(num_residues: int) -> cvpack.app.Topology:
    """One chain of ALA residues carrying N, H, CA, CB, C, O."""
    topology = cvpack.app.Topology()
    chain = topology.addChain()
    for _ in range(num_residues):
        residue = topology.addResidue("ALA", chain)
        for name, element in _BACKBONE:
            topology.addAtom(name, getattr(cvpack.app, element), residue)
    return topology


def ideal_helix(num_residues: int) -> np.ndarray:
    """N, H, CA, CB, C, O per residue on the screw axis fitted to the tabulated ideal alpha helix."""
    ideal = cvpack.helix_rmsd_content.ALPHA_POSITIONS.reshape(6, 5, 3)  # N CA CB C O
    a, b = ideal[:-1].reshape(-1, 3), ideal[1:].reshape(-1, 3)
    ca, cb = a.mean(0), b.mean(0)
    u, _, vt = np.linalg.svd((a - ca).T @ (b - cb))
    d = np.sign(np.linalg.det(vt.T @ u.T))
    rot = vt.T @ np.diag([1, 1, d]) @ u.T
    shift = cb - rot @ ca
    res, out = ideal[0], []
    for _ in range(num_residues):
        n, calpha, cbeta, c, o = res
        h = n + 0.1 * (n - calpha) / np.linalg.norm(n - calpha)
        out.append(np.stack([n, h, calpha, cbeta, c, o]))
        res = res @ rot.T + shift
    return np.concatenate(out)


def helix(num_frames: int, seed: int, device, num_residues: int = 500):
    """configs[3]: HelixRMSDContent and HelixTorsionContent on a noisy ideal helix."""
    topology = peptide(num_residues)
    base = torch.tensor(ideal_helix(num_residues), dtype=torch.float32, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    x = (base[None] + 0.02 * torch.randn((num_frames,) + tuple(base.shape), generator=gen, device=device)).contiguous()
    n = int(base.shape[0])
    system = unit_system(n)
    residues = list(topology.residues())
    cvs = {
        "helix_rmsd": cvpack.HelixRMSDContent(residues, n),
        "helix_torsion": cvpack.HelixTorsionContent(residues),
    }
    for cv in cvs.values():
        cv.addToSystem(system)
    return cvs, system, x, None, {"active_atoms": {"helix_rmsd": 5 * num_residues, "helix_torsion": 3 * num_residues}}


def walkers(num_frames: int, seed: int, device, num_atoms: int = 23000, box: float = 6.2):
    """configs[4]: one GPU's share of the multiple-walker batch."""
    x, lattice = lattice_frames(num_frames, num_atoms, box, seed, device)
    system = unit_system(num_atoms)
    contacts_cv = cvpack.NumberOfContacts(
        range(0, 2500), range(2500, 5000), nonbonded(num_atoms), thresholdDistance=0.6
    )
    composite = cvpack.CompositeRMSD(
        lattice.cpu().numpy().astype(np.float64), [range(0, 1500), range(1500, 2500)], num_atoms
    )
    for cv in (contacts_cv, composite):
        cv.addToSystem(system)
    return {"contacts": contacts_cv, "composite": composite}, system, x, [box] * 3, {"lattice": lattice}
