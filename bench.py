#!/usr/bin/env python
"""
Benchmark of the hot path: CV value + gradient frames/s on synthetic batches.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the hot path (value + per-atom gradient of the workload's CV) over one
batch of synthetic frames per GPU.  Default workload = BASELINE.json configs[1]: NumberOfContacts
between two 10k-atom groups, cubic periodic box, 1.2 nm cutoff, 4096 frames of water density.
For N > 1 launch with torchrun (one rank per GPU); frames are sharded by rank with no data-path
collective (weak scaling: every GPU processes its own 4096 frames); only the [B] values are
all-gathered over NCCL.

Prints ONE JSON line (rank 0): `value` = whole-job frames/s with inputs resident in HBM (CUDA
events, max over ranks); `e2e` = the same through the host-buffer API (pinned host -> device ->
pinned host every step); `roofline` for the dominant kernel timed live with CUDA events;
`cpu_baseline` = the C oracle (OpenMP, all host cores) on a bounded sample of the same workload.

`--impl reference` times that CPU oracle alone (the reference's own CPU algorithm restated: all
|g1| x |g2| pairs per frame) on the same config, each step a bounded sample of frames.
"""

from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# ---- workloads (BASELINE.json configs) --------------------------------------------------------------
WORKLOADS = {
    # configs[1]: the headline
    "contacts": dict(
        name="NumberOfContacts 10k x 10k atoms, cubic PBC L=5.844 nm, r0=0.6 rc=1.2 rs=0.9 nm, "
        "4096 frames/GPU of water density (100 atoms/nm^3)",
        frames=4096, atoms=20000, box=5.844, r0=0.6, cpu_sample_frames=32,
    ),
    # configs[2]
    "rmsd": dict(
        name="RMSD (quaternion alignment) of a 100k-atom synthetic protein, 2048 frames/GPU",
        frames=2048, atoms=100000, cpu_sample_frames=256,
    ),
    "rg": dict(
        name="RadiusOfGyration of a 100k-atom synthetic protein, 2048 frames/GPU",
        frames=2048, atoms=100000, cpu_sample_frames=256,
    ),
}


def make_positions(workload: str, spec: dict, num_frames: int, seed: int, device=None):
    """Synthetic coordinates [B, N, 3] float32 (torch, on `device` or CPU)."""
    import torch

    gen = torch.Generator(device="cpu" if device is None else device)
    gen.manual_seed(seed)
    dev = "cpu" if device is None else device
    n = spec["atoms"]
    if workload == "contacts":
        # jittered cubic lattice at water density, independent jitter per frame
        m = int(round(n ** (1.0 / 3.0)))
        while m**3 < n:
            m += 1
        cell = spec["box"] / m
        grid = torch.stack(
            torch.meshgrid(*[torch.arange(m, device=dev, dtype=torch.float32)] * 3, indexing="ij"), -1
        ).reshape(-1, 3)
        perm = torch.randperm(m**3, generator=gen, device=dev)[:n]
        lattice = (grid[perm] + 0.5) * cell
        x = lattice.unsqueeze(0) + 0.3 * cell * torch.randn(
            (num_frames, n, 3), generator=gen, device=dev, dtype=torch.float32
        )
        return x.contiguous(), lattice
    # protein-like blob (sigma = 3 nm) + per-frame noise and rigid rotation + translation
    reference = 3.0 * torch.randn((n, 3), generator=gen, device=dev, dtype=torch.float32)
    q = torch.randn((num_frames, 4), generator=gen, device=dev, dtype=torch.float32)
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
    x = torch.empty((num_frames, n, 3), device=dev, dtype=torch.float32)
    chunk = 64
    for first in range(0, num_frames, chunk):
        last = min(first + chunk, num_frames)
        noisy = reference.unsqueeze(0) + 0.05 * torch.randn(
            (last - first, n, 3), generator=gen, device=dev, dtype=torch.float32
        )
        x[first:last] = torch.einsum("bij,bnj->bni", rot[first:last], noisy) + 2.0 * torch.randn(
            (last - first, 1, 3), generator=gen, device=dev, dtype=torch.float32
        )
    return x.contiguous(), reference


def build_cv(workload: str, spec: dict, reference):
    """The cvpack-API collective variable of a workload and its system."""
    import cvpack_b200 as cvpack

    n = spec["atoms"]
    system = cvpack.System()
    for _ in range(n):
        system.addParticle(1.0)
    if workload == "contacts":
        nb = cvpack.NonbondedForce()
        for _ in range(n):
            nb.addParticle(0.0, 0.3, 0.0)
        nb.setNonbondedMethod(nb.CutoffPeriodic)
        cv = cvpack.NumberOfContacts(
            range(0, n // 2), range(n // 2, n), nb, thresholdDistance=spec["r0"],
            cutoffFactor=2.0, switchFactor=1.5,
        )
    elif workload == "rmsd":
        cv = cvpack.RMSD(reference.cpu().numpy().astype(np.float64), range(n), n)
    else:
        cv = cvpack.RadiusOfGyration(range(n))
    cv.addToSystem(system)
    return cv, system


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    QUERY = (
        "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
        "clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, device_index: int) -> None:
        self.samples = []
        self.proc = None
        self.device_index = device_index
        self.thread = None

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device_index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self) -> None:
        assert self.proc is not None and self.proc.stdout is not None
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self) -> dict:
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, sm_max, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for parts in self.samples:
            try:
                sm.append(float(parts[0]))
                sm_max.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(sm_max) if sm_max else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


def measured_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path, encoding="utf-8") as file:
            data = json.load(file)
        return {"hbm_gbs": data.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "source": "fallback"}


# DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from the
# committed `ncu --set full` captures, per frame, so that it can be quoted per launch of this run.
NCU_TRAFFIC = {
    "contacts": dict(bytes=2.145803e9 + 1.389061e9, frames=1483, source="profiles/r1_ncu_cell_sweep_sym.txt"),
    "rmsd": dict(bytes=3.373123e9 + 2.435028e9, frames=2048, source="profiles/r1_ncu_rmsd_fused_v3.txt"),
    "rg": dict(bytes=3.116965e9 + 2.407500e9, frames=2048, source="profiles/r1_ncu_rg_fused_v3.txt"),
}


def ncu_traffic(workload: str, frames_per_launch: float):
    """(bytes per launch, provenance) from the committed ncu capture, scaled by frames per launch."""
    entry = NCU_TRAFFIC.get(workload)
    if entry is None or not frames_per_launch:
        return None, None
    return entry["bytes"] / entry["frames"] * frames_per_launch, entry["source"]


# ---- CPU oracle arm ------------------------------------------------------------------------------------
def cpu_oracle_run(workload: str, spec: dict, positions: np.ndarray, reference) -> float:
    """Seconds the C oracle (OpenMP over frames, all host threads) needs for `positions`."""
    from oracle import c_oracle

    n = spec["atoms"]
    x = np.ascontiguousarray(positions, dtype=np.float64)
    start = time.perf_counter()
    if workload == "contacts":
        box = np.diag([spec["box"]] * 3)
        c_oracle.contacts_batch(
            x, range(0, n // 2), range(n // 2, n), (), box, r0=spec["r0"], rc=2 * spec["r0"],
            rs=1.5 * spec["r0"],
        )
    elif workload == "rmsd":
        c_oracle.rmsd_batch(x, np.arange(n), np.asarray(reference, dtype=np.float64))
    else:
        c_oracle.gyration_batch(x, np.arange(n))
    return time.perf_counter() - start


def run_reference(args, spec: dict) -> None:
    """`--impl reference`: the CPU oracle on the same config, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle

    c_oracle.use_all_threads()
    sample = spec["cpu_sample_frames"]
    x, reference = make_positions(args.workload, spec, sample, seed=1234)
    x = x.numpy()
    reference = reference.numpy()
    for _ in range(args.warmup):
        cpu_oracle_run(args.workload, spec, x[: max(1, sample // 4)], reference)
    seconds = [cpu_oracle_run(args.workload, spec, x, reference) for _ in range(args.steps)]
    total = sum(seconds)
    fps = sample * args.steps / total
    cores = c_oracle.num_threads()
    line = {
        "impl": "reference",
        "metric": "CV value+gradient frames/s",
        "value": fps,
        "unit": "frames/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": spec["name"], "frames_per_step": sample, "atoms": spec["atoms"]},
        "cpu_baseline": {
            "value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{sample} frames per step x {args.steps} steps, C oracle "
            "(all-pairs Reference-platform algorithm restated), OpenMP over frames",
        },
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- GPU arm ---------------------------------------------------------------------------------------------
def run_ours(args, spec: dict) -> None:
    import torch
    import torch.distributed as dist

    import cvpack_b200 as cvpack
    from cvpack_b200 import _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL's banner ("NCCL version ...", printed to stdout
        # at NCCL_DEBUG=VERSION and above) goes to a file unless the caller asked for a debug file
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO") and "NCCL_DEBUG_FILE" not in os.environ:
            os.environ["NCCL_DEBUG_FILE"] = os.path.join(os.environ.get("TMPDIR", "/tmp"), "nccl_debug.%h.%p.log")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    device = torch.device(f"cuda:{local_rank}")

    frames = args.frames or spec["frames"]
    n = spec["atoms"]
    # every rank generates its own shard (different seed): weak scaling, no data-path collective
    x, reference = make_positions(args.workload, spec, frames, seed=100 + rank, device=device)
    cv, system = build_cv(args.workload, spec, reference)
    box = None if args.workload != "contacts" else [spec["box"]] * 3
    ctx = cvpack.BatchContext(system, x, box, precision="single", device=device)
    native = ctx._native_cv(cv)  # pylint: disable=protected-access
    gathered = torch.empty(world * frames, dtype=torch.float32, device=device) if distributed else None

    def step():
        value, grad = cv.getValueAndGradient(ctx)
        if distributed:
            dist.all_gather_into_tensor(gathered, value._value)  # pylint: disable=protected-access
        return value, grad

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    native.set_option("time_kernels", 1)
    kernel_names = {"contacts": ["sweep"], "rmsd": ["rmsd_sums", "rmsd_gradient", "rmsd_fused"],
                    "rg": ["gyration_sums", "gyration_gradient", "gyration_fused"]}[args.workload]
    for name in kernel_names:
        native.kernel_time(name)
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local_rank)
    if rank == 0:
        sampler.start()
    launches_before = _native.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for _ in range(args.steps):
        step()
    stop.record()
    barrier()
    elapsed_ms = start.elapsed_time(stop)
    launches = _native.launch_count() - launches_before
    kernel_ms = {name: native.kernel_time(name) for name in kernel_names}
    native.set_option("time_kernels", 0)
    if distributed:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    # nvidia-smi answers every 100 ms at best: a timed region shorter than about half a second has
    # no clock sample of its own, so the same steps keep running (untimed, same count on every
    # rank) until the sampler has seen the load for a second
    clock_window = "timed region"
    if elapsed_ms < 500.0:
        extra = int(math.ceil(1000.0 / max(elapsed_ms / args.steps, 1e-3)))
        for _ in range(extra):
            step()
        barrier()
        clock_window = f"timed region + {extra} more untimed steps of the same load (region < 0.5 s)"
    clocks = sampler.stop() if rank == 0 else {}
    if rank == 0:
        clocks["window"] = clock_window
    fps = world * frames * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the host-buffer API (pinned host -> device -> pinned host) ----
    e2e_frames = frames
    host_ctx = cvpack.BatchContext(system, x.cpu(), box, precision="single", device=device, residency="host")
    del x
    value_h = torch.empty(frames, dtype=torch.float32, pin_memory=True)
    grad_h = torch.empty((frames, n, 3), dtype=torch.float32, pin_memory=True)
    host_ctx.evaluateInto(cv, value_h, grad_h)  # warm-up (also sizes the workspace)
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_ctx.evaluateInto(cv, value_h, grad_h)  # synchronous: results are in host memory
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if distributed:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_fps = world * e2e_frames * e2e_steps / e2e_s
    h2d = e2e_frames * n * 12 + (36 if box is not None else 0)
    d2h = e2e_frames * 4 + e2e_frames * n * 12
    checksum = float(value_h.double().sum())
    assert np.isfinite(checksum) and grad_h.shape == (e2e_frames, n, 3)

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peaks = measured_peaks()
    if args.workload == "contacts":
        fp32_peak = _native.measure_fp32_peak(local_rank)
        ms, count = kernel_ms["sweep"]
        volume = spec["box"] ** 3
        p_in = (4.0 / 3.0) * np.pi * (2 * spec["r0"]) ** 3 / volume
        # algorithmic work = |g1| x |g2| pair evaluations per frame x W flop, W = 21 per
        # pair test + 39 per in-range pair (both gradient sides); the two sweeps of a step share it
        flop_per_pair = 21.0 + 39.0 * p_in
        flop_per_step = frames * (n // 2) * (n // 2) * flop_per_pair
        achieved = args.steps * flop_per_step / (ms * 1e-3) / 1e12 if count else None
        traffic, traffic_source = ncu_traffic("contacts", frames * args.steps / count if count else 0)
        roofline = {
            "bound": "fp32", "kernel": "cell_sweep_kernel (or pair_sweep_kernel when cells=0)",
            "achieved": achieved, "peak": fp32_peak,
            "unit": "TFLOP/s", "frac": achieved / fp32_peak if achieved else None,
            "traffic": traffic, "traffic_source": traffic_source,
            "peak_source": "FFMA microbenchmark in this run (cvp_measure_fp32_peak)",
            "launches": count, "avg_launch_ms": ms / count if count else None,
            "kernel_share_of_step": ms / elapsed_ms,
            "flop_per_pair": flop_per_pair, "algorithmic_gflop_per_frame": flop_per_step / frames / 1e9,
            "note": "achieved = all-pairs-equivalent algorithmic flop / sweep time; the cell path "
            "skips out-of-range clusters, so this is work avoided as well as work done (DESIGN.md)",
        }
    else:
        used = {name: mc for name, mc in kernel_ms.items() if mc[1] > 0}
        total_ms = sum(mc[0] for mc in used.values())
        total_launches = sum(mc[1] for mc in used.values())
        # algorithmic bytes: 12 B/atom read once + 12 B/atom gradient written = 24 B/atom/frame
        bytes_per_step = frames * n * 24.0
        achieved = args.steps * bytes_per_step / (total_ms * 1e-3) / 1e9
        fused = any(name.endswith("_fused") for name in used)
        traffic, traffic_source = ncu_traffic(
            args.workload, frames * args.steps / total_launches if fused and total_launches else 0)
        roofline = {
            "bound": "hbm", "kernel": "+".join(sorted(used)), "achieved": achieved,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            "traffic": traffic, "traffic_source": traffic_source,
            "peak_source": f"MEASURED_PEAKS.json ({peaks['source']})",
            "launches": total_launches, "avg_launch_ms": total_ms / max(total_launches, 1),
            "kernel_share_of_step": total_ms / elapsed_ms,
            "algorithmic_bytes_per_frame": n * 24,
        }

    # ---- CPU baseline on a bounded sample ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle

        c_oracle.use_all_threads()
        sample = spec["cpu_sample_frames"]
        xs, ref = make_positions(args.workload, spec, sample, seed=1234)
        # repeat the sample until about 10 s of CPU work have been timed
        seconds, passes = 0.0, 0
        while seconds < 10.0 and passes < 4096:
            seconds += cpu_oracle_run(args.workload, spec, xs.numpy(), ref.numpy())
            passes += 1
        cpu = {
            "value": sample * passes / seconds, "unit": "frames/s", "cores": c_oracle.num_threads(),
            "kind": "port",
            "sample": f"{sample} frames of the same workload x {passes} passes, C oracle "
            f"(Reference-platform algorithm restated, OpenMP over frames), {seconds:.1f} s",
        }

    line = {
        "metric": "CV value+gradient frames/s",
        "value": fps,
        "unit": "frames/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(args.warmup, 3),
        "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": spec["name"], "frames_per_gpu": frames, "atoms": n,
            "l2": "inputs (>= 0.98 GB per step) are larger than the 126 MB L2",
            "parallelism": f"frames sharded over {world} GPU(s), values all-gathered",
        },
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


def main() -> None:
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=5)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", choices=["ours", "reference"], default="ours")
    parser.add_argument("--workload", choices=sorted(WORKLOADS), default="contacts")
    parser.add_argument("--frames", type=int, default=0, help="frames per GPU (default: the config's)")
    parser.add_argument("--no-cpu-baseline", action="store_true")
    args = parser.parse_args()
    spec = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, spec)
    else:
        run_ours(args, spec)


if __name__ == "__main__":
    main()
