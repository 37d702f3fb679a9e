import numpy as np, torch, sys
sys.path.insert(0, ".")
import cvpack_b200 as cvpack
from oracle import c_oracle
from tests.test_gpu_large import lattice_frames, unit_system, plain_nonbonded
rng = np.random.default_rng(33)
n, box = 20000, 5.844
x = lattice_frames(rng, 4, n, box).astype(np.float32)
system = unit_system(n)
nb = plain_nonbonded(n)
for g1 in (range(0, 10000), range(0, 5000), range(5000, 10000)):
    cv = cvpack.NumberOfContacts(g1, range(10000, n), nb, thresholdDistance=0.6)
    cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, x, [box] * 3)
    v1, g1g = cv.getValueAndGradient(ctx)
    ctx._native_cv(cv).set_option("cells", 0)
    v0, g0g = cv.getValueAndGradient(ctx)
    ref, refg = c_oracle.contacts_batch(x.astype(np.float64), g1, range(10000, n), (), np.diag([box] * 3), r0=0.6, rc=1.2, rs=0.9)
    print(len(g1), "cells", v1._value.double().cpu().numpy() - ref, "allpairs", v0._value.double().cpu().numpy() - ref,
          "grad err cells", float((g1g.double().cpu() - torch.from_numpy(refg)).abs().max()), "allpairs", float((g0g.double().cpu() - torch.from_numpy(refg)).abs().max()))
