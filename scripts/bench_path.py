#!/usr/bin/env python
"""Time PathInRMSDSpace 64 x 20k (configs[3]) alone, DMMA vs DFMA kernels."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cvpack_b200 as cvpack  # noqa: E402
from scripts import workloads  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
device = torch.device("cuda:0")
cv, system, x, box, info = workloads.path_rmsd(frames, 4, device)
ctx = cvpack.BatchContext(system, x, box, device=device)
native = ctx._native_cv(cv)
value = torch.empty(frames, device=device)
grad = torch.empty((frames, 20000, 3), device=device)
from cvpack_b200 import _native  # noqa: E402
print(json.dumps({"fp64_dfma_peak_tflops": _native.measure_fp64_peak(0), "fp32_ffma_peak_tflops": _native.measure_fp32_peak(0)}))
for mma in (1, 2, 0):
    native.set_option("multi_mma", mma)
    for _ in range(2):
        ctx.evaluateInto(cv, value, grad)
    native.set_option("time_kernels", 1)
    native.kernel_time("rmsd_multi_cov"), native.kernel_time("rmsd_multi_gradient")
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(reps):
        ctx.evaluateInto(cv, value, grad)
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop) / reps
    cov, _ = native.kernel_time("rmsd_multi_cov")
    grd, _ = native.kernel_time("rmsd_multi_gradient")
    native.set_option("time_kernels", 0)
    gflop = frames * 2.0 * 9 * 64 * 20000 / 1e9
    print(json.dumps({"mma": mma, "frames": frames, "frames_per_s": frames / ms * 1e3, "ms": ms,
                      "cov_ms": cov / reps, "grad_ms": grd / reps, "cov_tflops": gflop / (cov / reps),
                      "grad_tflops": gflop / (grd / reps), "checksum": float(value.double().sum())}))
