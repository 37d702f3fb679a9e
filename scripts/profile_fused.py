"""One evaluation of the fused streaming kernel for ncu: python scripts/profile_fused.py rmsd|rg [frames] [sub] [lag]"""
import sys, torch
sys.path.insert(0, ".")
import bench, cvpack_b200 as cvpack
which = sys.argv[1] if len(sys.argv) > 1 else "rmsd"
spec = bench.WORKLOADS[which]
dev = torch.device("cuda:0")
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
x, ref = bench.make_positions(which, spec, frames, 100, device=dev)
cv, system = bench.build_cv(which, spec, ref)
ctx = cvpack.BatchContext(system, x, None, device=dev)
nat = ctx._native_cv(cv)
nat.set_option("fused_slice", int(sys.argv[3]) if len(sys.argv) > 3 else 0)
nat.set_option("fused_lag", int(sys.argv[4]) if len(sys.argv) > 4 else 3)
nat.set_option("fused_slots", int(sys.argv[5]) if len(sys.argv) > 5 else 0)
value = torch.empty(frames, device=dev); grad = torch.empty_like(x)
for _ in range(3): ctx.evaluateInto(cv, value, grad)
torch.cuda.synchronize()
print("ok", float(value.sum()))
