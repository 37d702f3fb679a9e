#!/usr/bin/env python
"""The other pair-family CVs at the size of configs[1] (SURVEY 8d lists the ResidueCoordination
variant: 1000 residues x 10 atoms per group, r0 = 1.0, PBC): frames/s and kernel breakdown."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cvpack_b200 as cvpack  # noqa: E402
from cvpack_b200 import app  # noqa: E402
from scripts import workloads  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
device = torch.device("cuda:0")
n, box = 20000, 5.844
x, _ = workloads.lattice_frames(frames, n, box, 7, device)
system = workloads.unit_system(n)
rng = np.random.default_rng(7)


def timed(cv, label, reps=3, kernels=("sweep", "cell_sort")):
    ctx = cvpack.BatchContext(system, x, [box] * 3, device=device)
    native = ctx._native_cv(cv) if hasattr(ctx, "_native_cv") else None
    value = torch.empty(frames, device=device)
    grad = torch.empty((frames, n, 3), device=device)
    for _ in range(2):
        ctx.evaluateInto(cv, value, grad)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ctx.evaluateInto(cv, value, grad)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(json.dumps({"cv": label, "frames": frames, "ms_per_batch": ms, "frames_per_s": frames / ms * 1e3,
                      "value_mean": float(value.double().mean())}))


# ResidueCoordination: one topology of 2000 residues x 10 atoms = the 20k atoms in order
topology = app.Topology()
chain = topology.addChain()
all_residues = []
for r in range(2000):
    residue = topology.addResidue("ALA", chain)
    for k in range(10):
        topology.addAtom(f"C{k}", app.Element.getBySymbol("C"), residue)
    all_residues.append(residue)
cv = cvpack.ResidueCoordination(all_residues[:1000], all_residues[1000:], pbc=True, thresholdDistance=1.0,
                                weighByMass=False, includeHydrogens=True)
cv.addToSystem(system)
timed(cv, "ResidueCoordination 1000 x 1000 residues of 10 atoms, r0 = 1.0 nm, PBC")

nb = cvpack.NonbondedForce()
for _ in range(n):
    nb.addParticle(float(rng.normal() * 0.5), float(rng.uniform(0.2, 0.4)), float(rng.uniform(0.1, 1.0)))
nb.setNonbondedMethod(nb.CutoffPeriodic)
nb.setCutoffDistance(1.2)
nb.setUseSwitchingFunction(True)
nb.setSwitchingDistance(0.9)
cv = cvpack.AttractionStrength(range(0, n // 2), range(n // 2, n), nb)
cv.addToSystem(system)
timed(cv, "AttractionStrength 10k x 10k, rc = 1.2 nm")

cv = cvpack.ShortestDistance(range(0, n // 2), range(n // 2, n), n, cutoffDistance=1.2)
cv.addToSystem(system)
timed(cv, "ShortestDistance 10k x 10k, rc = 1.2 nm")
