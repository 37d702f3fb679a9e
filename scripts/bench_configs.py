#!/usr/bin/env python
"""
BASELINE.json configs[3] and configs[4] on one GPU (device-resident batches, CUDA events):

  path     PathInRMSDSpace, 64 milestones x 20k atoms, 1024 frames (progress and deviation)
  helix    HelixRMSDContent (495 blocks) + HelixTorsionContent (996 torsions), 500-residue helix,
           1024 frames
  walkers  NumberOfContacts (2 500 x 2 500 atoms, PBC) + CompositeRMSD (1 500 + 1 000 atoms) on
           23k-atom walkers, 1024 walkers (one GPU's share of 8192)

    python scripts/bench_configs.py [path|helix|walkers ...] [--frames F] [--reps R]

Prints one JSON line per collective variable: frames/s for value+gradient and the algorithmic
bytes (coordinates of the CV's atoms read once + their gradient written once) per second.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cvpack_b200 as cvpack  # noqa: E402
from tests import helpers  # noqa: E402


def unit_system(n):
    system = cvpack.System()
    for _ in range(n):
        system.addParticle(1.0)
    return system


def timed(ctx, cv, reps):
    b, n = ctx.getNumFrames(), ctx.getSystem().getNumParticles()
    value = torch.empty(b, device=ctx.getDevice())
    grad = torch.empty((b, n, 3), device=ctx.getDevice())
    for _ in range(3):
        ctx.evaluateInto(cv, value, grad)
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(reps):
        ctx.evaluateInto(cv, value, grad)
    stop.record()
    torch.cuda.synchronize()
    return start.elapsed_time(stop) / reps, value, grad


def report(config, name, ctx, cv, active_atoms, reps, extra=None):
    ms, value, grad = timed(ctx, cv, reps)
    b = ctx.getNumFrames()
    line = {
        "config": config, "cv": name, "frames": b, "atoms": ctx.getSystem().getNumParticles(),
        "ms_per_batch": ms, "frames_per_s": b / ms * 1e3,
        "algorithmic_GBps": b * active_atoms * 24 / ms / 1e6,
        "value_mean": float(value.double().mean()), "grad_absmax": float(grad.abs().max()),
    }
    line.update(extra or {})
    print(json.dumps(line), flush=True)
    return value, grad


def run_path(frames, reps, device):
    rng = np.random.default_rng(4)
    n, m = 20000, 64
    a, b = 2.0 * rng.normal(size=(n, 3)), 2.0 * rng.normal(size=(n, 3))
    b = a + 0.3 * (b - a) / np.linalg.norm(b - a, axis=1, keepdims=True)
    milestones = [a + (b - a) * k / (m - 1) + 0.02 * rng.normal(size=(n, 3)) for k in range(m)]
    gen = torch.Generator(device=device).manual_seed(5)
    t = 20 + 20 * torch.rand(frames, generator=gen, device=device)
    ta, tb = torch.tensor(a, device=device, dtype=torch.float32), torch.tensor(b, device=device, dtype=torch.float32)
    x = ta[None] + (tb - ta)[None] * (t / (m - 1))[:, None, None]
    x = x + 0.02 * torch.randn(x.shape, generator=gen, device=device)
    system = unit_system(n)
    refs = [{i: row for i, row in enumerate(ms)} for ms in milestones]
    out = {}
    for metric, label in ((cvpack.path.progress, "progress"), (cvpack.path.deviation, "deviation")):
        cv = cvpack.PathInRMSDSpace(metric, refs, n, 0.05)
        cv.addToSystem(system)
        ctx = cvpack.BatchContext(system, x, None, device=device)
        value, _ = report("configs[3]", f"PathInRMSDSpace[{label}] 64 x 20k", ctx, cv, n, reps)
        out[label] = value
    s = out["progress"].double().cpu().numpy() * (m - 1)
    # the progress variable must recover where along the path each frame was generated
    err = float(np.abs(s - t.double().cpu().numpy()).max())
    print(json.dumps({"config": "configs[3]", "check": "max |progress*(m-1) - generated position|", "value": err}))


def run_helix(frames, reps, device):
    rng = np.random.default_rng(6)
    nres = 500
    topology = helpers.make_peptide(nres)
    base = helpers.ideal_helix_positions(nres, rng)
    x = base[None] + 0.02 * rng.normal(size=(frames,) + base.shape)
    n = base.shape[0]
    system = unit_system(n)
    residues = list(topology.residues())
    cvs = [
        ("HelixRMSDContent 495 blocks", cvpack.HelixRMSDContent(residues, n), 5 * nres),
        ("HelixTorsionContent 996 torsions", cvpack.HelixTorsionContent(residues), 3 * nres),
        ("HelixAngleContent", cvpack.HelixAngleContent(residues), nres),
        ("HelixHBondContent", cvpack.HelixHBondContent(residues), 2 * nres),
    ]
    for _, cv, _ in cvs:
        cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, x, None, device=device)
    for name, cv, active in cvs:
        report("configs[3]", name, ctx, cv, active, reps)


def run_walkers(frames, reps, device):
    rng = np.random.default_rng(7)
    n, box = 23000, 6.2
    m = int(np.ceil(n ** (1 / 3)))
    grid = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)
    sites = torch.tensor((grid[rng.permutation(m**3)[:n]] + 0.5) * (box / m), dtype=torch.float32, device=device)
    gen = torch.Generator(device=device).manual_seed(8)
    x = sites[None] + 0.3 * (box / m) * torch.randn((frames, n, 3), generator=gen, device=device)
    system = unit_system(n)
    nb = cvpack.NonbondedForce()
    for _ in range(n):
        nb.addParticle(0.0, 0.3, 0.0)
    nb.setNonbondedMethod(nb.CutoffPeriodic)
    contacts = cvpack.NumberOfContacts(range(0, 2500), range(2500, 5000), nb, thresholdDistance=0.6)
    reference = sites.cpu().numpy().astype(np.float64)
    composite = cvpack.CompositeRMSD(reference, [range(0, 1500), range(1500, 2500)], n)
    for cv in (contacts, composite):
        cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, x, [box] * 3, device=device)
    report("configs[4]", "NumberOfContacts 2500 x 2500 of 23k atoms", ctx, contacts, 5000, reps)
    report("configs[4]", "CompositeRMSD 1500 + 1000 of 23k atoms", ctx, composite, 2500, reps)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("configs", nargs="*", default=["path", "helix", "walkers"])
    parser.add_argument("--frames", type=int, default=1024)
    parser.add_argument("--reps", type=int, default=5)
    args = parser.parse_args()
    device = torch.device("cuda:0")
    for name in args.configs:
        {"path": run_path, "helix": run_helix, "walkers": run_walkers}[name](args.frames, args.reps, device)


if __name__ == "__main__":
    main()
