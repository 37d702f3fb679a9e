"""Sweep the sub-slice size and lag of the fused streaming kernel (RMSD / Rg, 100k atoms, 2048 frames)."""
import sys, torch
sys.path.insert(0, ".")
import cvpack_b200 as cvpack
from scripts import workloads
which = sys.argv[1] if len(sys.argv) > 1 else "rmsd"
subs = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
lags = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [3]
slots_list = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
hints_list = [int(v) for v in sys.argv[5].split(",")] if len(sys.argv) > 5 else [3]
spec = {"frames": 2048, "atoms": 100000}
dev = torch.device("cuda:0")
cv, system, x, _, _ = (workloads.rmsd if which == "rmsd" else workloads.gyration)(spec["frames"], 3, dev)
ctx = cvpack.BatchContext(system, x, None, device=dev)
nat = ctx._native_cv(cv)
value = torch.empty(spec["frames"], device=dev); grad = torch.empty_like(x)
def run(reps=20):
    for _ in range(3): ctx.evaluateInto(cv, value, grad)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): ctx.evaluateInto(cv, value, grad)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    return ms, spec["frames"] * spec["atoms"] * 24 / ms / 1e6
nat.set_option("fused", 0)
ms, gbs = run()
print(f"two-pass: {ms:.3f} ms  {gbs:.0f} GB/s ({gbs/6450.9:.2f})", flush=True)
v0 = value.clone(); g0 = grad.clone()
nat.set_option("fused", 1)
import itertools
for sub, slots, lag, hints in itertools.product(subs, slots_list, lags, hints_list):
    nat.set_option("fused_slice", sub); nat.set_option("fused_lag", lag)
    nat.set_option("fused_slots", slots); nat.set_option("fused_hints", hints)
    value.zero_(); grad.zero_()
    try:
        ms, gbs = run()
    except Exception as exc:
        print(f"sub {sub:5d} slots {slots:2d} lag {lag:3d} hints {hints}: {exc}"); continue
    dv = (value - v0).abs().max().item() / v0.abs().max().item()
    dg = (grad - g0).abs().max().item() / g0.abs().max().item()
    print(f"sub {sub:5d} slots {slots:2d} lag {lag:3d} hints {hints}: {ms:.3f} ms  {gbs:.0f} GB/s ({gbs/6450.9:.2f})  dv {dv:.1e} dg {dg:.1e}", flush=True)
