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

Prints ONE JSON line (rank 0):
  value        whole-job frames/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e          the same through the host-buffer API (pinned host -> device -> pinned host each step)
  roofline     of the dominant kernel, timed live with CUDA events on its stream
  parity       the CPU-baseline sample frames evaluated on the GPU *inside a batch that takes the
               benchmarked kernel path* and compared with the C oracle (the run fails above 1e-5)
  cpu_baseline the C oracle on a bounded sample: cell-list algorithm (the analogue of OpenMM's CPU
               platform), with the all-pairs Reference-platform loop quoted beside it
  extra        (N = 1) one record per other BASELINE config -- RMSD and Rg on 100k atoms, the path CV
               with 64 x 20k references, helix contents, the walker batch, configs[1] in FP64 mode --
               each with frames/s, its own roofline, clock sample, CPU baseline and parity
  scale_extra  (N > 1) strong scaling of the fixed 4096-frame batch and configs[4] (8192 walkers
               sharded by walker) with the gradient gathered dense / compact, over NCCL and through
               the library's own peer-memory push kernel, with and without overlap

`--impl reference` times the CPU oracle alone on the same config (cell list, all host threads),
each step a bounded sample of frames.
"""

from __future__ import annotations

import argparse
import hashlib
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

RECORD: list = []  # the JSON line, printed by main() once stdout is ours again
HEADLINE = dict(frames=4096, atoms=20000, box=5.844, r0=0.6, cpu_sample_frames=32)
TOLERANCE = 1e-5  # north_star: FP32 mode within 1e-5 relative of the Reference-platform result


# ---- clocks --------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the benchmark runs; windows are cut out of
    the one stream of samples by time stamp."""

    QUERY = (
        "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
        "clocks_event_reasons.sw_power_cap"
    )
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

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
                self.samples.append((time.monotonic(), parts))

    def window(self, begin: float, end: float) -> dict:
        """Median SM clock and throttle reasons of the samples taken in [begin, end]."""
        sm, sm_max, reasons = [], [], set()
        for stamp, parts in list(self.samples):
            if stamp < begin or stamp > end + 0.05:
                continue
            try:
                sm.append(float(parts[0]))
                sm_max.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(self.NAMES, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(sm_max) if sm_max else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }

    def stop(self) -> None:
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()


def pin_to_gpu_numa_node(device_index: int):
    """Restrict this process to the host cores NVML lists as local to the GPU
    (nvmlDeviceGetCpuAffinity) before any pinned memory is allocated: with one rank per GPU, first
    touch then places the staging buffers of the host path on the GPU's own NUMA node instead of
    wherever torchrun started the rank.  Returns the core list as a string, or None when NVML gives
    no usable answer -- fewer than four cores that this process may run on is treated as such: a
    rank squeezed onto one core launches kernels slower than the GPU runs them."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cores = {64 * w + bit for w, word in enumerate(mask) for bit in range(64) if (int(word) >> bit) & 1}
    except Exception:  # pylint: disable=broad-except
        return None
    try:
        allowed = cores & os.sched_getaffinity(0)
        if len(allowed) < 4:
            return None
        os.sched_setaffinity(0, allowed)
    except (ValueError, OSError, AttributeError):
        return None
    ordered = sorted(allowed)
    return f"{len(ordered)} cores, {ordered[0]}-{ordered[-1]}"


def measured_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path, encoding="utf-8") as file:
            data = json.load(file)
        return {"hbm_gbs": data.get("hbm_gbs", 6650.0), "source": "MEASURED_PEAKS.json (measured)"}
    return {"hbm_gbs": 6650.0, "source": "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"}


# ---- DRAM traffic from the committed ncu captures, valid only for the build that was profiled ----------
def kernel_source_hash(files=None) -> str:
    """Hash of the CUDA sources a kernel is built from (all of csrc/ when `files` is None): a
    committed ncu capture speaks for this build only if it matches."""
    digest = hashlib.sha256()
    base = os.path.join(ROOT, "cvpack_b200", "csrc")
    names = sorted(files) if files else sorted(n for n in os.listdir(base) if n.endswith((".cu", ".cuh")))
    for name in names:
        with open(os.path.join(base, name), "rb") as file:
            digest.update(name.encode() + b"\0" + file.read())
    return digest.hexdigest()[:16]


def ncu_traffic(kernel: str, frames_per_launch: float):
    """(bytes per launch or None, provenance).  profiles/traffic.json holds, per kernel, the
    dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture, the frames of that
    launch, the source files of the kernel and their hash in the build that was profiled
    (scripts/make_traffic.py writes it from the committed summaries)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path) or not frames_per_launch:
        return None, "no committed ncu capture"
    with open(path, encoding="utf-8") as file:
        table = json.load(file)
    entry = table.get(kernel)
    if entry is None:
        return None, "no committed ncu capture for this kernel"
    now = kernel_source_hash(entry.get("files"))
    if entry.get("source_hash") != now:
        return None, (f"{entry['source']} profiled build {entry.get('source_hash')}, this build is "
                      f"{now}: traffic not carried over")
    return entry["bytes"] / entry["frames"] * frames_per_launch, entry["source"]


def ncu_pipes(kernel: str):
    """Pipe utilisation of the committed capture of `kernel` (None unless it is of this build)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    with open(path, encoding="utf-8") as file:
        entry = json.load(file).get(kernel)
    if entry is None or entry.get("source_hash") != kernel_source_hash(entry.get("files")):
        return None
    return entry.get("pipes")


# ---- helpers -------------------------------------------------------------------------------------------
def relative_errors(value, grad, ref_value, ref_grad) -> dict:
    """Worst relative value error and worst gradient error relative to the frame's largest component."""
    value, grad = np.asarray(value, dtype=np.float64), np.asarray(grad, dtype=np.float64)
    rel_value = float(np.max(np.abs(value - ref_value) / np.maximum(np.abs(ref_value), 1e-3)))
    scale = np.maximum(np.abs(ref_grad).reshape(len(ref_grad), -1).max(axis=1), 1e-6)
    rel_grad = float(np.max(np.abs(grad - ref_grad).reshape(len(ref_grad), -1).max(axis=1) / scale))
    outside = bool(np.any((ref_grad == 0) & (grad != 0)))
    return {"max_rel_value": rel_value, "max_rel_grad": rel_grad, "nonzero_outside_cv": outside}


def timed_loop(fn, steps: int, warmup: int, min_seconds: float = 0.0, after_timed=None):
    """CUDA-event timing of `steps` calls (each call timed on its own as well); keeps calling (same
    load, untimed) until `min_seconds` have passed so that the clock sampler sees the load."""
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    events = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    begin = time.monotonic()
    events[0].record()
    for k in range(steps):
        fn()
        events[k + 1].record()
    torch.cuda.synchronize()
    per_step = [events[k].elapsed_time(events[k + 1]) for k in range(steps)]
    total_ms = events[0].elapsed_time(events[steps])
    if after_timed is not None:
        after_timed()  # (kernel timers are read here: the continuation below is not part of the region)
    while time.monotonic() - begin < min_seconds:
        fn()
        torch.cuda.synchronize()
    return total_ms, per_step, (begin, time.monotonic())


# ---- CPU oracle arm -------------------------------------------------------------------------------------
def cpu_contacts(x: np.ndarray, spec: dict, cells: bool, need_grad: bool = True):
    from oracle import c_oracle

    n = spec["atoms"]
    box = np.diag([spec["box"]] * 3)
    start = time.perf_counter()
    value, grad = c_oracle.contacts_batch(
        np.ascontiguousarray(x, dtype=np.float64), range(0, n // 2), range(n // 2, n), (), box,
        r0=spec["r0"], rc=2 * spec["r0"], rs=1.5 * spec["r0"], cells=cells, need_grad=need_grad,
    )
    return time.perf_counter() - start, value, grad


def run_reference(args) -> None:
    """`--impl reference`: the CPU oracle on the headline config, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import c_oracle
    from scripts import workloads

    c_oracle.use_all_threads()
    spec = HEADLINE
    sample = spec["cpu_sample_frames"]
    x, _ = workloads.lattice_frames(sample, spec["atoms"], spec["box"], 1234, torch.device("cpu"))
    x = x.numpy()
    for _ in range(args.warmup):
        cpu_contacts(x[: max(1, sample // 4)], spec, cells=True)
    seconds = [cpu_contacts(x, spec, cells=True)[0] for _ in range(args.steps)]
    total = sum(seconds)
    fps = sample * args.steps / total
    all_pairs_s = cpu_contacts(x[:8], spec, cells=False)[0]
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
        "config": {"workload": workloads.CONFIGS["contacts"], "frames_per_step": sample,
                   "atoms": spec["atoms"]},
        "cpu_baseline": {
            "value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{sample} frames per step x {args.steps} steps, C oracle with a cell list "
            "(neighbour-list analogue of OpenMM's CPU platform; OpenMM itself cannot be installed "
            "here), FP64, OpenMP over frames",
            "all_pairs_frames_per_s": 8 / all_pairs_s,
            "all_pairs_note": "the same frames through the all-pairs loop of OpenMM's Reference platform",
        },
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    RECORD.append(json.dumps(line))


def headline_cpu_baseline(spec: dict):
    """Cell-list C oracle on the sample frames for about 10 s; returns (record, x, value, grad)."""
    import torch

    from oracle import c_oracle
    from scripts import workloads

    c_oracle.use_all_threads()
    sample = spec["cpu_sample_frames"]
    xs, _ = workloads.lattice_frames(sample, spec["atoms"], spec["box"], 1234, torch.device("cpu"))
    xs = xs.numpy()
    seconds, passes = 0.0, 0
    value = grad = None
    while seconds < 10.0 and passes < 4096:
        elapsed, value, grad = cpu_contacts(xs, spec, cells=True)
        seconds += elapsed
        passes += 1
    all_pairs_s, ap_value, _ = cpu_contacts(xs[:8], spec, cells=False, need_grad=False)
    assert np.allclose(ap_value, value[:8], rtol=1e-12)
    record = {
        "value": sample * passes / seconds, "unit": "frames/s", "cores": c_oracle.num_threads(),
        "kind": "port",
        "sample": f"{sample} frames of the same workload x {passes} passes, C oracle with a cell list "
        f"(neighbour-list analogue of OpenMM's CPU platform), FP64, OpenMP over frames, {seconds:.1f} s",
        "all_pairs_frames_per_s": 8 / all_pairs_s,
        "all_pairs_note": "8 of the frames through the all-pairs loop of OpenMM's Reference platform",
    }
    return record, xs, value, grad


# ---- extras (N = 1): the other BASELINE configs ------------------------------------------------------------
def oracle_sample(cv, system, x_sample: np.ndarray, box):
    """NumPy / C oracle on a few frames through the test bridge: (seconds, values, gradients)."""
    from tests import helpers

    box_np = None if box is None else np.diag(box)
    start = time.perf_counter()
    out = [helpers.oracle_evaluate(cv, system, frame, box_np) for frame in x_sample]
    seconds = time.perf_counter() - start
    return seconds, np.array([v for v, _ in out]), np.stack([g for _, g in out])


def extra_record(name, config, cv, system, x, box, sampler, steps, peaks, *, precision="single",
                 kernels=(), roofline_fn=None, cpu_fn=None, cpu_frames=4, with_e2e=False):
    """Time one CV on a device-resident batch and describe it."""
    import torch

    import cvpack_b200 as cvpack
    from scripts import workloads

    device = x.device
    if precision == "double":
        x = x.double()
    ctx = cvpack.BatchContext(system, x, box, precision=precision, device=device)
    frames, n = int(x.shape[0]), int(x.shape[1])
    dtype = torch.float32 if precision == "single" else torch.float64
    value = torch.empty(frames, dtype=dtype, device=device)
    grad = torch.empty((frames, n, 3), dtype=dtype, device=device)
    native = None if hasattr(cv, "_evaluateComposite") else ctx._native_cv(cv)  # pylint: disable=protected-access

    def step():
        ctx.evaluateInto(cv, value, grad)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if native is not None and kernels:
        native.set_option("time_kernels", 1)
        for kernel in kernels:
            native.kernel_time(kernel)
    kernel_ms = {}

    def read_timers():
        if native is not None and kernels:
            kernel_ms.update({kernel: native.kernel_time(kernel) for kernel in kernels})
            native.set_option("time_kernels", 0)

    total_ms, per_step, window = timed_loop(step, steps, 0, min_seconds=0.7, after_timed=read_timers)
    ms = total_ms / steps
    record = {
        "name": name, "config": config, "workload": workloads.CONFIGS.get(name, name),
        "frames": frames, "atoms": n, "dtype": "f32" if precision == "single" else "f64",
        "value": frames / ms * 1e3, "unit": "frames/s", "steps": steps, "ms_per_step": ms,
        "per_step_ms": {"min": min(per_step), "median": statistics.median(per_step), "max": max(per_step)},
        "clocks": sampler.window(*window),
    }
    if roofline_fn is not None:
        record["roofline"] = roofline_fn(frames, n, ms, per_step, kernel_ms, steps, peaks)
    if cpu_fn is not None:
        sample = x[:cpu_frames].double().cpu().numpy()
        seconds, ref_value, ref_grad, cores, what = cpu_fn(cv, system, sample, box)
        record["cpu_baseline"] = {
            "value": cpu_frames / seconds, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{cpu_frames} frames of the same batch, {what}, {seconds:.1f} s",
        }
        tol = TOLERANCE if precision == "single" else 1e-10
        parity = relative_errors(value[:cpu_frames].double().cpu().numpy(),
                                 grad[:cpu_frames].double().cpu().numpy(), ref_value, ref_grad)
        parity.update(frames=cpu_frames, tolerance=tol,
                      ok=parity["max_rel_value"] <= tol and parity["max_rel_grad"] <= tol
                      and not parity["nonzero_outside_cv"])
        record["parity"] = parity
    if with_e2e:
        host_ctx = cvpack.BatchContext(system, x.cpu(), box, precision=precision, device=device, residency="host")
        value_h = torch.empty(frames, dtype=dtype, pin_memory=True)
        grad_h = torch.empty((frames, n, 3), dtype=dtype, pin_memory=True)
        host_ctx.evaluateInto(cv, value_h, grad_h)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            host_ctx.evaluateInto(cv, value_h, grad_h)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / 2
        es = value.element_size()
        record["e2e"] = {"value": frames / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": frames * n * 3 * es,
                         "d2h_bytes_per_step": frames * n * 3 * es + frames * es,
                         "host_gb_per_s_each_way": frames * n * 3 * es / e2e_s / 1e9}
        del host_ctx, value_h, grad_h
    del ctx, value, grad
    torch.cuda.empty_cache()
    return record


def hbm_roofline(bytes_per_atom):
    def build(frames, n, ms, per_step, kernel_ms, steps, peaks):
        used = {k: v for k, v in kernel_ms.items() if v[1] > 0}
        kernel_total = sum(v[0] for v in used.values())
        launches = sum(v[1] for v in used.values())
        bytes_per_step = frames * n * float(bytes_per_atom)
        achieved = steps * bytes_per_step / (kernel_total * 1e-3) / 1e9 if kernel_total else None
        fused = [k for k in used if k.endswith("_fused")]
        traffic, source = ncu_traffic(fused[0] if fused else "", frames * steps / launches if launches else 0)
        fracs = sorted(bytes_per_step / (t * 1e-3) / 1e9 / peaks["hbm_gbs"] for t in per_step)
        return {
            "bound": "hbm", "kernel": "+".join(sorted(used)), "achieved": achieved, "peak": peaks["hbm_gbs"],
            "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"] if achieved else None,
            "traffic": traffic, "traffic_source": source, "peak_source": peaks["source"],
            "launches": launches, "avg_launch_ms": kernel_total / max(launches, 1),
            "kernel_share_of_step": kernel_total / (ms * steps) if kernel_total else None,
            "algorithmic_bytes_per_frame": n * bytes_per_atom,
            # whole steps (kernel + launch gaps), one figure per timed iteration
            "frac_per_step": {"min": fracs[0], "median": statistics.median(fracs), "max": fracs[-1]},
        }
    return build


def run_extras(device, sampler, steps, peaks, select=None) -> list:
    import torch

    from cvpack_b200 import _native
    from oracle import c_oracle
    from scripts import workloads

    c_oracle.use_all_threads()
    records = []

    def wanted(*names):
        return select is None or any(name in select for name in names)

    fp64_peak = _native.measure_fp64_peak(device.index or 0)
    fp32_peak = _native.measure_fp32_peak(device.index or 0)

    def c_rmsd(cv, system, sample, box):
        ref = np.array([cv._args["referencePositions"][a] for a in cv._args["group"]])  # pylint: disable=protected-access
        start = time.perf_counter()
        value, grad = c_oracle.rmsd_batch(sample, np.arange(sample.shape[1]), ref)
        return time.perf_counter() - start, value, grad, c_oracle.num_threads(), "C oracle (Horn/Jacobi, OpenMP over frames)"

    def c_gyration(cv, system, sample, box):
        start = time.perf_counter()
        value, grad = c_oracle.gyration_batch(sample, np.arange(sample.shape[1]))
        return time.perf_counter() - start, value, grad, c_oracle.num_threads(), "C oracle (OpenMP over frames)"

    def numpy_oracle(cv, system, sample, box):
        seconds, value, grad = oracle_sample(cv, system, sample, box)
        return seconds, value, grad, 1, "NumPy FP64 oracle (oracle/cv_oracle.py), one thread"

    # configs[1], the other CV it names: ResidueCoordination between the two 10k-atom groups
    if wanted("residue_coordination"):
        cv, system, x, box, info = workloads.residue_coordination(1024, 2, device)
        pairs = info["residues"] ** 2

        def coordination_roofline(frames, n, ms, per_step, kernel_ms, steps_, peaks_):
            # 10^6 centroid pairs x W with p_in = 1 (no cutoff: every pair counts)
            m, count = kernel_ms.get("sweep", (0.0, 0))
            flop = frames * pairs * 60.0
            achieved = steps_ * flop / (m * 1e-3) / 1e12 if count else None
            return {"bound": "fp32", "kernel": "centroid pair sweep", "achieved": achieved, "peak": fp32_peak,
                    "unit": "TFLOP/s", "frac": achieved / fp32_peak if achieved else None, "traffic": None,
                    "launches": count, "kernel_share_of_step": m / (ms * steps_) if count else None,
                    "algorithmic_gflop_per_frame": flop / frames / 1e9}

        records.append(extra_record("residue_coordination", "configs[1]", cv, system, x, box, sampler,
                                    max(steps, 10), peaks, kernels=("sweep",), roofline_fn=coordination_roofline,
                                    cpu_fn=numpy_oracle, cpu_frames=2))
        del cv, system, x
        torch.cuda.empty_cache()

    # configs[2]: streaming RMSD / Rg on 100k atoms
    for name, builder, kernels, cpu in (
        ("rmsd", workloads.rmsd, ("rmsd_sums", "rmsd_gradient", "rmsd_fused"), c_rmsd),
        ("rg", workloads.gyration, ("gyration_sums", "gyration_gradient", "gyration_fused"), c_gyration),
    ):
        if not wanted(name):
            continue
        cv, system, x, box, _ = builder(2048, 3, device)
        records.append(extra_record(
            name, "configs[2]", cv, system, x, box, sampler, max(steps, 20), peaks, kernels=kernels,
            roofline_fn=hbm_roofline(24), cpu_fn=cpu, cpu_frames=64, with_e2e=name == "rmsd",
        ))
        del cv, system, x
        torch.cuda.empty_cache()

    # configs[3]: path CV with 64 references of 20k atoms
    def path_roofline(frames, n, ms, per_step, kernel_ms, steps_, peaks_):
        used = {k: v for k, v in kernel_ms.items() if v[1] > 0}
        kernel_total = sum(v[0] for v in used.values())
        flop = frames * 2.0 * (2 * 9 * 64 * n)  # covariance + gradient contractions, FMA = 2
        achieved = steps_ * flop / (kernel_total * 1e-3) / 1e12 if kernel_total else None
        # DRAM bytes of one covariance + one gradient launch (both see every frame of the step)
        parts = [ncu_traffic(k, frames) for k in ("rmsd_multi_cov", "rmsd_multi_gradient")]
        traffic = sum(p[0] for p in parts) if all(p[0] is not None for p in parts) else None
        return {
            "bound": "fp64", "kernel": "+".join(sorted(used)), "achieved": achieved, "peak": fp64_peak,
            "unit": "TFLOP/s", "frac": achieved / fp64_peak if achieved else None, "traffic": traffic,
            "traffic_source": parts[0][1] + " (covariance + gradient launch)" if traffic else parts[0][1],
            "peak_source": "DFMA microbenchmark in this run (cvp_measure_fp64_peak)",
            "launches": sum(v[1] for v in used.values()),
            "kernel_share_of_step": kernel_total / (ms * steps_) if kernel_total else None,
            "algorithmic_mflop_per_frame": flop / frames / 1e6,
            "hbm_frac": frames * n * 24.0 / (ms * 1e-3) / 1e9 / peaks_["hbm_gbs"],
        }

    if wanted("path"):
        cv, system, x, box, info = workloads.path_rmsd(1024, 4, device)
        records.append(extra_record(
            "path", "configs[3]", cv, system, x, box, sampler, max(steps, 10), peaks,
            kernels=("rmsd_multi_cov", "rmsd_multi_gradient", "rmsd_sums", "rmsd_gradient"),
            roofline_fn=path_roofline, cpu_fn=numpy_oracle, cpu_frames=3,
        ))
        del cv, system, x, info
        torch.cuda.empty_cache()

    # configs[3]: helix contents (latency bound: frames/s and achieved bytes only)
    cvs, system, x, box, info = workloads.helix(1024 if wanted("helix_rmsd", "helix_torsion") else 2, 6, device)
    for name, kernels in (("helix_rmsd", ("rmsd_blocks",)), ("helix_torsion", ())):
        if not wanted(name):
            continue
        active = info["active_atoms"][name]

        def helix_roofline(frames, n, ms, per_step, kernel_ms, steps_, peaks_, active=active):
            gbs = frames * active * 24.0 / (ms * 1e-3) / 1e9
            return {"bound": "latency", "achieved": gbs, "unit": "GB/s", "peak": peaks_["hbm_gbs"],
                    "frac": gbs / peaks_["hbm_gbs"], "traffic": None,
                    "note": "algorithmic bytes (24 B per atom the CV touches) per second; launch/latency bound"}

        records.append(extra_record(name, "configs[3]", cvs[name], system, x, box, sampler, max(steps, 20),
                                    peaks, kernels=kernels, roofline_fn=helix_roofline, cpu_fn=numpy_oracle,
                                    cpu_frames=8))
    del cvs, system, x
    torch.cuda.empty_cache()

    # configs[4]: one GPU's share of the walker batch
    if not wanted("walkers_contacts", "walkers_composite_rmsd", "contacts_fp64"):
        return records
    cvs, system, x, box, _ = workloads.walkers(1024 if wanted("walkers_contacts", "walkers_composite_rmsd") else 2, 8, device)

    def c_walker_contacts(cv, system_, sample, box_):
        start = time.perf_counter()
        value, grad = c_oracle.contacts_batch(sample, range(0, 2500), range(2500, 5000), (), np.diag(box_),
                                              r0=0.6, rc=1.2, rs=0.9, cells=True)
        return time.perf_counter() - start, value, grad, c_oracle.num_threads(), "C oracle with a cell list (OpenMP over frames)"

    def pair_roofline(n1, n2, box_edge, peak, bound):
        def build(frames, n, ms, per_step, kernel_ms, steps_, peaks_):
            m, count = kernel_ms.get("sweep", (0.0, 0))
            p_in = (4.0 / 3.0) * np.pi * 1.2**3 / box_edge**3
            flop = frames * n1 * n2 * (21.0 + 39.0 * p_in)
            achieved = steps_ * flop / (m * 1e-3) / 1e12 if count else None
            return {"bound": bound, "kernel": "cell_sweep_kernel", "achieved": achieved, "peak": peak,
                    "unit": "TFLOP/s", "frac": achieved / peak if achieved else None, "traffic": None,
                    "launches": count, "kernel_share_of_step": m / (ms * steps_) if count else None,
                    "algorithmic_gflop_per_frame": flop / frames / 1e9}
        return build

    if wanted("walkers_contacts"):
        records.append(extra_record("walkers_contacts", "configs[4]", cvs["contacts"], system, x, box, sampler,
                                    max(steps, 10), peaks, kernels=("sweep",),
                                    roofline_fn=pair_roofline(2500, 2500, 6.2, fp32_peak, "fp32"),
                                    cpu_fn=c_walker_contacts, cpu_frames=64))
    if wanted("walkers_composite_rmsd"):
        records.append(extra_record("walkers_composite_rmsd", "configs[4]", cvs["composite"], system, x, box,
                                    sampler, max(steps, 20), peaks, cpu_fn=numpy_oracle, cpu_frames=32))
    del cvs, system, x
    torch.cuda.empty_cache()
    if not wanted("contacts_fp64"):
        return records

    # configs[1] in FP64 mode (north_star: 1e-10 of the Reference platform)
    cv, system, x, box, _ = workloads.contacts(256, 5, device)

    def c_contacts64(cv_, system_, sample, box_):
        seconds, value, grad = cpu_contacts(sample, HEADLINE, cells=True)
        return seconds, value, grad, c_oracle.num_threads(), "C oracle with a cell list (OpenMP over frames)"

    records.append(extra_record("contacts_fp64", "configs[1], FP64 mode", cv, system, x, box, sampler,
                                max(2, steps // 2), peaks, precision="double", kernels=("sweep",),
                                roofline_fn=pair_roofline(10000, 10000, 5.844, fp64_peak, "fp64"),
                                cpu_fn=c_contacts64, cpu_frames=8))
    return records


# ---- scaling extras (N > 1) -----------------------------------------------------------------------------------
def run_scale_extras(device, rank, world, steps) -> dict:
    import torch
    import torch.distributed as dist

    import cvpack_b200 as cvpack
    from cvpack_b200 import distributed
    from scripts import workloads

    out = {}

    def max_over_ranks(ms: float) -> float:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, reps):
        fn()
        fn()
        dist.barrier()
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(reps):
            fn()
        stop.record()
        dist.barrier()
        torch.cuda.synchronize()
        return max_over_ranks(start.elapsed_time(stop)) / reps

    # (a) strong scaling: the fixed 4096-frame batch of configs[1] split over the ranks
    total = HEADLINE["frames"]
    begin, end = cvpack.shard_frames(total, rank, world)
    cv, system, x, box, _ = workloads.contacts(end - begin, 100 + rank, device)
    ctx = cvpack.BatchContext(system, x, box, device=device)
    value = torch.empty(end - begin, device=device)
    grad = torch.empty((end - begin, HEADLINE["atoms"], 3), device=device)
    gathered = torch.empty(total, device=device)

    def strong_step():
        ctx.evaluateInto(cv, value, grad)
        dist.all_gather_into_tensor(gathered, value)

    ms = timed(strong_step, steps)
    out["strong"] = {
        "workload": workloads.CONFIGS["contacts"].replace("4096 frames/GPU", "4096 frames in total"),
        "frames_total": total, "frames_per_gpu": end - begin, "value": total / ms * 1e3, "unit": "frames/s",
        "ms_per_step": ms, "scaling": "strong",
        "path": {0: "all pairs", 1: "cells, two sweeps", 2: "cells, symmetric sweep"}.get(
            int(ctx._native_cv(cv).stat("last_path"))),  # pylint: disable=protected-access
        "limiter": "one CTA per frame: with B/N frames per GPU the sweep fills 296 CTA slots in "
        f"{math.ceil((end - begin) / 296)} waves; below 148 frames per GPU it falls back to the two-sweep split",
    }
    del ctx, cv, system, x, value, grad
    torch.cuda.empty_cache()

    # (b) configs[4]: 8192 walkers sharded by walker, gradient of the contact CV gathered everywhere
    walkers_total = 8192
    begin, end = cvpack.shard_frames(walkers_total, rank, world)
    local = end - begin
    cvs, system, x, box, _ = workloads.walkers(local, 200 + rank, device)
    n = 23000
    ctx = cvpack.BatchContext(system, x, box, device=device)
    values = {k: torch.empty(local, device=device) for k in cvs}
    grads = {k: torch.empty((local, n, 3), device=device) for k in cvs}
    native = ctx._native_cv(cvs["contacts"])  # pylint: disable=protected-access
    active = torch.as_tensor(native.active_atoms().astype(np.int64), device=device)
    active32 = active.int()

    def compute():
        for key, cv_ in cvs.items():
            ctx.evaluateInto(cv_, values[key], grads[key])

    compute_ms = timed(compute, steps)
    dense_bytes = walkers_total * n * 12
    compact_bytes = walkers_total * int(active.numel()) * 12
    record = {
        "workload": workloads.CONFIGS["walkers"] + f", {walkers_total} walkers sharded over {world} GPUs",
        "walkers_total": walkers_total, "walkers_per_gpu": local, "unit": "frames/s",
        "compute_only": {"value": walkers_total / compute_ms * 1e3, "ms_per_step": compute_ms},
        "gradient_bytes_dense": dense_bytes, "gradient_bytes_compact": compact_bytes,
        "active_atoms": int(active.numel()),
    }

    def gather_record(ms, nbytes):
        extra = max(ms - compute_ms, 1e-6)
        return {"value": walkers_total / ms * 1e3, "ms_per_step": ms, "gather_ms_exposed": ms - compute_ms,
                "received_gb_per_s_per_gpu": nbytes * (world - 1) / world / (extra * 1e-3) / 1e9}

    def nccl_dense():
        compute()
        distributed.all_gather_gradients(grads["contacts"], walkers_total)

    def nccl_compact():
        compute()
        distributed.all_gather_gradients(grads["contacts"], walkers_total, active_atoms=active)

    record["nccl_dense"] = gather_record(timed(nccl_dense, steps), dense_bytes)
    record["nccl_compact"] = gather_record(timed(nccl_compact, steps), compact_bytes)
    torch.cuda.empty_cache()

    if hasattr(distributed, "PeerGather"):
        try:
            peer = distributed.PeerGather(walkers_total, n, torch.float32, device)
            half = local // 2
            side = torch.cuda.Stream(device=device)

            def peer_run(compact: bool, overlap: bool):
                """overlap: the contact CV is evaluated in two slabs; slab 1 travels while slab 2 is
                computed, slab 2 while the second CV (CompositeRMSD) is computed."""
                atoms = active32 if compact else None

                def evaluate(key, lo, hi):
                    ctx._run(cvs[key], True, positions=x[lo:hi], out=(values[key][lo:hi], grads[key][lo:hi]))  # pylint: disable=protected-access

                def push_after(lo, hi):
                    done = torch.cuda.Event()
                    done.record(torch.cuda.current_stream(device))
                    side.wait_event(done)
                    with torch.cuda.stream(side):
                        peer.push(grads["contacts"][lo:hi], begin + lo, atoms)

                def run():
                    if overlap:
                        evaluate("contacts", 0, half)
                        push_after(0, half)
                        evaluate("contacts", half, local)
                        push_after(half, local)
                        evaluate("composite", 0, local)
                    else:
                        evaluate("contacts", 0, local)
                        evaluate("composite", 0, local)
                        push_after(0, local)
                    result = peer.finish(side)
                    torch.cuda.current_stream(device).wait_stream(side)
                    return result
                return run

            for compact in (False, True):
                for overlap in (False, True):
                    key = f"peer_{'compact' if compact else 'dense'}{'_overlapped' if overlap else ''}"
                    record[key] = gather_record(timed(peer_run(compact, overlap), steps),
                                                compact_bytes if compact else dense_bytes)
            # the pushed buffer must equal the NCCL gather
            result = peer_run(False, False)()
            torch.cuda.synchronize()
            reference = distributed.all_gather_gradients(grads["contacts"], walkers_total)
            record["peer_equals_nccl"] = bool(torch.equal(result, reference))
            record["peer_timed_out"] = peer.timed_out()
            del reference
            dist.barrier()
            peer.close()
        except Exception as error:  # pylint: disable=broad-except
            record["peer_error"] = f"{type(error).__name__}: {error}"
    record["limiter"] = (
        "compute is independent per walker; the dense gather moves "
        f"{dense_bytes / 1e9:.2f} GB per CV to every GPU over NVLink, the compact one "
        f"{compact_bytes / 1e9:.2f} GB ({int(active.numel())} of {n} atoms carry a gradient)"
    )
    out["walkers"] = record
    return out


# ---- GPU arm ---------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import cvpack_b200 as cvpack
    from cvpack_b200 import _native
    from scripts import workloads

    spec = HEADLINE
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL's banner goes to a file
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO") and "NCCL_DEBUG_FILE" not in os.environ:
            os.environ["NCCL_DEBUG_FILE"] = os.path.join(os.environ.get("TMPDIR", "/tmp"), "nccl_debug.%h.%p.log")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    device = torch.device(f"cuda:{local_rank}")
    affinity = pin_to_gpu_numa_node(local_rank) if distributed and not args.no_affinity else None

    frames = args.frames or spec["frames"]
    n = spec["atoms"]
    # every rank generates its own shard (different seed): weak scaling, no data-path collective
    cv, system, x, box, _ = workloads.contacts(frames, 100 + rank, device, n, spec["box"], spec["r0"])
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

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step()
    barrier()
    native.set_option("time_kernels", 1)
    native.kernel_time("sweep")
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local_rank)
    if rank == 0:
        sampler.start()
    launches_before = _native.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.monotonic()
    start.record()
    for _ in range(args.steps):
        step()
    stop.record()
    barrier()
    elapsed_ms = start.elapsed_time(stop)
    launches = _native.launch_count() - launches_before
    sweep_ms, sweep_count = native.kernel_time("sweep")
    native.set_option("time_kernels", 0)
    kernel_path = int(native.stat("last_path"))
    if distributed:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    # nvidia-smi answers every 100 ms at best: a timed region shorter than about half a second has
    # no clock sample of its own, so the same steps keep running (untimed, same count on every rank)
    clock_window = "timed region"
    if elapsed_ms < 500.0:
        more = int(math.ceil(1000.0 / max(elapsed_ms / args.steps, 1e-3)))
        for _ in range(more):
            step()
        barrier()
        clock_window = f"timed region + {more} more untimed steps of the same load (region < 0.5 s)"
    t_end = time.monotonic()
    clocks = {}
    if rank == 0:
        clocks = sampler.window(t_begin, t_end)
        clocks["window"] = clock_window
    fps = world * frames * args.steps / (elapsed_ms * 1e-3)

    # ---- parity of the benchmarked path: the CPU sample frames inside a batch of >= 148 frames ----
    parity = cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu, xs, ref_value, ref_grad = headline_cpu_baseline(spec)
        sample = len(xs)
        batch = x[: max(160, sample)].clone()
        batch[64:64 + sample] = torch.as_tensor(xs, device=device)
        pctx = cvpack.BatchContext(system, batch, box, precision="single", device=device)
        pvalue, pgrad = cv.getValueAndGradient(pctx)
        parity = relative_errors(pvalue._value[64:64 + sample].double().cpu().numpy(),  # pylint: disable=protected-access
                                 pgrad[64:64 + sample].double().cpu().numpy(), ref_value, ref_grad)
        ppath = int(pctx._native_cv(cv).stat("last_path"))  # pylint: disable=protected-access
        parity.update(
            frames=sample, batch=int(batch.shape[0]), tolerance=TOLERANCE,
            path={0: "all pairs", 1: "cells, two sweeps", 2: "sym"}.get(ppath),
            same_path_as_timed_region=ppath == kernel_path,
            oracle="C oracle (cell list; equal to the all-pairs loop to 1e-12 on 8 of the frames), FP64, "
            "on the FP32 coordinates promoted to double",
            ok=parity["max_rel_value"] <= TOLERANCE and parity["max_rel_grad"] <= TOLERANCE
            and not parity["nonzero_outside_cv"],
        )
        del pctx, batch, pvalue, pgrad
    if distributed:
        dist.barrier()

    # ---- end to end through the host-buffer API (pinned host -> device -> pinned host) ----
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
    e2e_fps = world * frames * e2e_steps / e2e_s
    h2d = frames * n * 12 + 36
    d2h = frames * 4 + frames * n * 12
    assert np.isfinite(float(value_h.double().sum())) and grad_h.shape == (frames, n, 3)
    del host_ctx, value_h, grad_h, ctx
    torch.cuda.empty_cache()

    # ---- scaling extras (all ranks take part) / extras (single GPU) ----
    scale_extra = extra = None
    if distributed and not args.no_extras:
        scale_extra = run_scale_extras(device, rank, world, max(3, min(args.steps, 5)))
    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    if not distributed and not args.no_extras:
        select = None if args.extras == "all" else set(args.extras.split(","))
        extra = run_extras(device, sampler, args.steps, peaks, select)
    sampler.stop()

    # ---- roofline of the dominant kernel ----
    fp32_peak = _native.measure_fp32_peak(local_rank)
    volume = spec["box"] ** 3
    p_in = (4.0 / 3.0) * np.pi * (2 * spec["r0"]) ** 3 / volume
    # algorithmic work = |g1| x |g2| pair evaluations per frame x W flop, W = 21 per pair
    # test + 39 per in-range pair (both gradient sides)
    flop_per_pair = 21.0 + 39.0 * p_in
    flop_per_step = frames * (n // 2) * (n // 2) * flop_per_pair
    achieved = args.steps * flop_per_step / (sweep_ms * 1e-3) / 1e12 if sweep_count else None
    traffic, traffic_source = ncu_traffic("cell_sweep_sym", frames * args.steps / sweep_count if sweep_count else 0)
    roofline = {
        "bound": "fp32", "kernel": "cell_sweep_kernel", "path": {0: "all pairs", 1: "cells, two sweeps",
                                                                 2: "cells, symmetric sweep"}.get(kernel_path),
        "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": achieved / fp32_peak if achieved else None,
        "traffic": traffic, "traffic_source": traffic_source,
        "executed": ncu_pipes("cell_sweep_sym"),
        "peak_source": "FFMA microbenchmark in this run (cvp_measure_fp32_peak)",
        "launches": sweep_count, "avg_launch_ms": sweep_ms / sweep_count if sweep_count else None,
        "kernel_share_of_step": sweep_ms / elapsed_ms,
        "flop_per_pair": flop_per_pair, "algorithmic_gflop_per_frame": flop_per_step / frames / 1e9,
        "note": "achieved = all-pairs-equivalent algorithmic flop / sweep time (SURVEY 8d): the cell path "
        "skips out-of-range clusters and computes each pair once for both groups, so this counts work "
        "avoided as well as work done and can exceed the FMA peak; what the pipes did is in `executed` "
        "(ncu capture of this build under profiles/)",
    }

    line = {
        "metric": "CV value+gradient frames/s",
        "value": fps,
        "unit": "frames/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": warmup,
        "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": workloads.CONFIGS["contacts"], "frames_per_gpu": frames, "atoms": n,
            "l2": "inputs (>= 0.98 GB per step) are larger than the 126 MB L2",
            "parallelism": f"frames sharded over {world} GPU(s), values all-gathered",
        },
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "host_gb_per_s_per_rank_each_way": h2d * e2e_steps / e2e_s / 1e9,
                "cpu_affinity_rank0": affinity},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "parity": parity,
    }
    if extra is not None:
        line["extra"] = extra
    if scale_extra is not None:
        line["scale_extra"] = scale_extra
    RECORD.append(json.dumps(line))
    if distributed:
        dist.destroy_process_group()
    failed = [r["name"] for r in (extra or []) if r.get("parity") and not r["parity"]["ok"]]
    if parity is not None and not parity["ok"]:
        failed.insert(0, "headline")
    if failed:
        print(f"bench.py: parity above tolerance for {failed}", file=sys.stderr)
        sys.exit(1)


def main() -> None:
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=5)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", choices=["ours", "reference"], default="ours")
    parser.add_argument("--frames", type=int, default=0, help="frames per GPU (default: the config's)")
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--no-extras", action="store_true", help="headline only")
    parser.add_argument("--no-affinity", action="store_true",
                        help="N > 1: do not pin the ranks to the host cores local to their GPU")
    parser.add_argument("--extras", default="all", help="comma-separated subset of the extra records "
                        "(rmsd,rg,path,helix_rmsd,helix_torsion,walkers_contacts,walkers_composite_rmsd,contacts_fp64)")
    args = parser.parse_args()
    # stdout carries exactly ONE line (the JSON record): whatever libraries print there (NCCL's
    # version banner, torchrun notices) is sent to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
        if RECORD:
            print(RECORD[-1], flush=True)


if __name__ == "__main__":
    main()
