#!/usr/bin/env python
"""Write profiles/traffic.json from the committed ncu summaries (scripts/ncu_summary.py output):
    python scripts/make_traffic.py
Per kernel: DRAM bytes (read + write) of the captured launch, the frames that launch processed, the
CUDA sources the kernel is built from and their hash *now* -- so run it right after the capture,
with the sources that were profiled.  bench.py carries the figure into `roofline.traffic` only
while that hash still matches."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

# kernel key in bench.py -> (summary file, kernel substring, frames of the captured launch, sources)
CAPTURES = {
    "cell_sweep_sym": ("r2_ncu_cell_sweep_sym.txt", "cell_sweep_kernel", 296,
                       ["pair_cells.cuh", "pair_common.cuh", "pair_sum.cu", "common.cuh"]),
    "rmsd_fused": ("r2_ncu_rmsd_fused.txt", "stream_fused_kernel", 2048,
                   ["stream_fused.cuh", "rmsd.cu", "common.cuh"]),
    "gyration_fused": ("r2_ncu_rg_fused.txt", "stream_fused_kernel", 2048,
                       ["stream_fused.cuh", "gyration.cu", "common.cuh"]),
    "rmsd_multi_cov": ("r2_ncu_rmsd_multi_mma.txt", "rmsd_multi_cov_ring_kernel", 1024,
                       ["rmsd_multi.cuh", "rmsd.cu", "common.cuh"]),
    "rmsd_multi_gradient": ("r2_ncu_rmsd_multi_mma.txt", "rmsd_multi_gradient_pk_kernel", 1024,
                            ["rmsd_multi.cuh", "rmsd.cu", "common.cuh"]),
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram_bytes(path, kernel):
    total, inside = 0.0, False
    for line in open(path, encoding="utf-8"):
        if line.startswith("=="):
            if inside:
                break
            inside = kernel in line
        elif inside:
            match = re.match(r"\s+dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
            if match:
                total += float(match.group(2)) * UNIT[match.group(3)]
    return total


PIPES = {"sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_slots_busy_pct",
         "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
         "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
         "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
         "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
         "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_active_of_32",
         "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
         "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflict_wavefronts",
         "smsp__cycles_active.avg": "cycles_active",
         "gpu__time_duration.sum": "duration"}


def pipes(path, kernel):
    """What the pipes did in the captured launch (the executed side of the roofline)."""
    out, inside = {}, False
    for line in open(path, encoding="utf-8"):
        if line.startswith("=="):
            if inside:
                break
            inside = kernel in line
        elif inside:
            parts = line.split()
            if len(parts) >= 2 and parts[0] in PIPES:
                out[PIPES[parts[0]]] = float(parts[1])
    if "shared_wavefronts" in out and "cycles_active" in out:
        # one wavefront per cycle and SM
        out["shared_pipe_busy_pct"] = 100.0 * out["shared_wavefronts"] / (out["cycles_active"] * 148)
    return out


table = {}
for key, (name, kernel, frames, files) in CAPTURES.items():
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        continue
    size = dram_bytes(path, kernel)
    if size:
        table[key] = {"bytes": size, "frames": frames, "files": files, "pipes": pipes(path, kernel),
                      "source_hash": bench.kernel_source_hash(files),
                      "source": f"profiles/{name} (ncu --set full, one launch of {frames} frames)"}
with open(os.path.join(ROOT, "profiles", "traffic.json"), "w", encoding="utf-8") as file:
    json.dump(table, file, indent=1)
print(json.dumps(table, indent=1))
