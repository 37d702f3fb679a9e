"""One evaluation of the fused streaming kernel for ncu: python scripts/profile_fused.py rmsd|rg [frames] [sub] [lag]
(sub / lag 0 or absent: the library's defaults)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cvpack_b200 as cvpack  # noqa: E402
from scripts import workloads  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "rmsd"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
dev = torch.device("cuda:0")
cv, system, x, box, _ = (workloads.rmsd if which == "rmsd" else workloads.gyration)(frames, 3, dev)
ctx = cvpack.BatchContext(system, x, box, device=dev)
nat = ctx._native_cv(cv)
if len(sys.argv) > 3 and int(sys.argv[3]):
    nat.set_option("fused_slice", int(sys.argv[3]))
if len(sys.argv) > 4 and int(sys.argv[4]):
    nat.set_option("fused_lag", int(sys.argv[4]))
value = torch.empty(frames, device=dev)
grad = torch.empty_like(x)
for _ in range(3):
    ctx.evaluateInto(cv, value, grad)
torch.cuda.synchronize()
print("ok", float(value.sum()))
