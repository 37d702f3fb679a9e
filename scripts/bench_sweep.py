#!/usr/bin/env python
"""Time the cell sweep of configs[1] alone (device-resident frames) -- for A/B runs of tuning
variants: CVPACK_B200_LIB=build/variants/libcvpack_b200_X.so python scripts/bench_sweep.py [frames]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cvpack_b200 as cvpack  # noqa: E402
from cvpack_b200 import _native  # noqa: E402
from scripts import workloads  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
device = torch.device("cuda:0")
cv, system, x, box, _ = workloads.contacts(frames, 100, device)
ctx = cvpack.BatchContext(system, x, box, device=device)
native = ctx._native_cv(cv)
value = torch.empty(frames, device=device)
grad = torch.empty((frames, 20000, 3), device=device)
for _ in range(2):
    ctx.evaluateInto(cv, value, grad)
native.set_option("time_kernels", 1)
native.kernel_time("sweep")
torch.cuda.synchronize()
start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
start.record()
for _ in range(reps):
    ctx.evaluateInto(cv, value, grad)
stop.record()
torch.cuda.synchronize()
ms = start.elapsed_time(stop) / reps
sweep_ms, launches = native.kernel_time("sweep")
path = int(native.stat("last_path"))
# accuracy of the j side against the two-sweep path (no fixed point) on the first 160 frames
sub = cvpack.BatchContext(system, x[:160], box, device=device)
v1, g1 = cv.getValueAndGradient(sub)
sub_native = sub._native_cv(cv)
sub_native.set_option("sym", 0)
v2, g2 = cv.getValueAndGradient(sub)
err = float(((g1 - g2).abs().amax(dim=(1, 2)) / g2.abs().amax(dim=(1, 2))).max())
print(json.dumps({"lib": os.path.basename(_native.library_path()), "frames": frames, "path": path,
                  "frames_per_s": frames / ms * 1e3, "ms": ms, "sweep_ms_per_launch": sweep_ms / max(launches, 1),
                  "sym_vs_two_sweep_rel_grad": err}))
