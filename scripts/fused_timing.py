"""Per-frame timeline of the fused RMSD kernel (diagnostics build: make EXTRA=-DFUSED_TIMING).
python scripts/fused_timing.py [sub] [lag]"""
import ctypes, sys, numpy as np, torch
sys.path.insert(0, ".")
import bench, cvpack_b200 as cvpack
from cvpack_b200 import _native
spec = bench.WORKLOADS["rmsd"]
dev = torch.device("cuda:0")
frames = 2048
x, ref = bench.make_positions("rmsd", spec, frames, 100, device=dev)
cv, system = bench.build_cv("rmsd", spec, ref)
ctx = cvpack.BatchContext(system, x, None, device=dev)
nat = ctx._native_cv(cv)
nat.set_option("fused_slice", int(sys.argv[1]) if len(sys.argv) > 1 else 1024)
nat.set_option("fused_lag", int(sys.argv[2]) if len(sys.argv) > 2 else 3)
value = torch.empty(frames, device=dev); grad = torch.empty_like(x)
for _ in range(3): ctx.evaluateInto(cv, value, grad)
torch.cuda.synchronize()
lib = _native.load()
lib.cvp_debug_set_fused_timing.argtypes = [ctypes.c_void_p]
buf = torch.zeros((frames, 6), dtype=torch.int64, device=dev)
buf[:, 0] = 2**62; buf[:, 4] = 2**62
assert lib.cvp_debug_set_fused_timing(ctypes.c_void_p(buf.data_ptr())) == 0
ctx.evaluateInto(cv, value, grad)
torch.cuda.synchronize()
lib.cvp_debug_set_fused_timing(None)
t = buf.cpu().numpy().astype(np.float64)
t0 = t[:, 0].min()
sel = slice(200, 1800)  # steady state
def stat(name, v):
    v = v[sel] / 1e3
    print(f"{name:46s} median {np.median(v):8.2f} us   p10 {np.percentile(v,10):8.2f}   p90 {np.percentile(v,90):8.2f}")
print("kernel span (first record .. last consumer): %.1f us" % ((t[:, 5].max() - t0) / 1e3))
stat("skew: last - first record of a frame", t[:, 1] - t[:, 0])
stat("gather: records gathered - last record", t[:, 2] - t[:, 1])
stat("solve + publish: published - gathered", t[:, 3] - t[:, 2])
stat("slack: first consumer needs - published", t[:, 4] - t[:, 3])
stat("consumer span: last has - first needs", t[:, 5] - t[:, 4])
stat("chain: published - first record", t[:, 3] - t[:, 0])
stat("lag: first consumer needs - first record", t[:, 4] - t[:, 0])
