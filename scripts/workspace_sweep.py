import sys, torch
sys.path.insert(0, "/root/repo")
import bench, cvpack_b200 as cvpack
spec = bench.WORKLOADS["contacts"]
dev = torch.device("cuda:0")
x, ref = bench.make_positions("contacts", spec, spec["frames"], 100, device=dev)
cv, system = bench.build_cv("contacts", spec, ref)
ctx = cvpack.BatchContext(system, x, [spec["box"]] * 3, device=dev)
nat = ctx._native_cv(cv)
value = torch.empty(spec["frames"], device=dev); grad = torch.empty_like(x)
for gib in (1, 2, 3, 4, 6):
    nat.set_workspace_limit(int(gib * 2**30))
    for _ in range(2): ctx.evaluateInto(cv, value, grad)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): ctx.evaluateInto(cv, value, grad)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"workspace {gib} GiB: {ms:.1f} ms/step  {spec['frames']/ms*1e3:.0f} frames/s  checksum {float(value.double().sum()):.6f}", flush=True)
