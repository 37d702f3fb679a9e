#!/usr/bin/env python
"""Kernel breakdown of the configs[4] contact sum (2.5k x 2.5k atoms of 23k-atom walkers)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cvpack_b200 as cvpack  # noqa: E402
from scripts import workloads  # noqa: E402
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
device = torch.device("cuda:0")
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cvs, system, x, box, _ = workloads.walkers(frames, seed, device)
cv = cvs["contacts"]
ctx = cvpack.BatchContext(system, x, box, device=device)
native = ctx._native_cv(cv)
value = torch.empty(frames, device=device)
grad = torch.empty((frames, x.shape[1], 3), device=device)
for _ in range(2): ctx.evaluateInto(cv, value, grad)
native.set_option("time_kernels", 1)
for k in ("sweep", "cell_sort"): native.kernel_time(k)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
reps = 5
for _ in range(reps): ctx.evaluateInto(cv, value, grad)
b.record(); torch.cuda.synchronize()
out = {"frames": frames, "ms": a.elapsed_time(b) / reps, "path": int(native.stat("last_path"))}
for k in ("sweep", "cell_sort"):
    t, c = native.kernel_time(k)
    out[k + "_ms"] = t / reps; out[k + "_launches"] = c / reps
print(json.dumps(out))
