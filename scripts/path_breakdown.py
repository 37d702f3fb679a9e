import sys, json, torch, numpy as np
sys.path.insert(0, ".")
import cvpack_b200 as cvpack
sys.argv = ["x"]
from scripts import bench_configs as bc
dev = torch.device("cuda:0")
rng = np.random.default_rng(4)
n, m, frames = 20000, 64, 1024
a = 2.0 * rng.normal(size=(n, 3)); b = a + 0.3 * rng.normal(size=(n, 3))
ms = [a + (b - a) * k / (m - 1) for k in range(m)]
x = torch.tensor(np.stack([ms[20 + (i % 20)] for i in range(8)]), dtype=torch.float32, device=dev).repeat(frames // 8, 1, 1)
x += 0.02 * torch.randn_like(x)
system = bc.unit_system(n)
cv = cvpack.PathInRMSDSpace(cvpack.path.progress, [{i: r for i, r in enumerate(q)} for q in ms], n, 0.05)
cv.addToSystem(system)
ctx = cvpack.BatchContext(system, x, None, device=dev)
nat = ctx._native_cv(cv)
value = torch.empty(frames, device=dev); grad = torch.empty_like(x)
for _ in range(2): ctx.evaluateInto(cv, value, grad)
nat.set_option("time_kernels", 1)
for _ in range(3): ctx.evaluateInto(cv, value, grad)
torch.cuda.synchronize()
for k in ("rmsd_sums", "rmsd_solve", "rmsd_gradient", "rmsd_fused"):
    t, c = nat.kernel_time(k); print(k, t / max(c, 1), "ms x", c)
