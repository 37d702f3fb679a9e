"""
GPU parity at sizes beyond what the NumPy oracle finishes quickly -- checked against the C oracle --
and, at BASELINE.json's full sizes, through size-independent properties (translation / rotation
invariance, permutation of frames, linearity of the pair sum in the groups, zero net gradient).
"""

import numpy as np
import pytest
import torch

import cvpack_b200 as cvpack
from oracle import c_oracle
from tests import helpers

pytestmark = pytest.mark.gpu


def lattice_frames(rng, num_frames, num_atoms, box):
    m = int(np.ceil(num_atoms ** (1 / 3)))
    grid = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)
    sites = (grid[rng.permutation(m**3)[:num_atoms]] + 0.5) * (box / m)
    return sites[None] + 0.3 * (box / m) * rng.normal(size=(num_frames, num_atoms, 3))


def unit_system(n):
    system = cvpack.System()
    for _ in range(n):
        system.addParticle(1.0)
    return system


def plain_nonbonded(n, periodic=True):
    nb = cvpack.NonbondedForce()
    for _ in range(n):
        nb.addParticle(0.0, 0.3, 0.0)
    nb.setNonbondedMethod(nb.CutoffPeriodic if periodic else nb.CutoffNonPeriodic)
    return nb


@pytest.mark.parametrize("precision,tol", [("single", 1e-5), ("double", 1e-10)])
@pytest.mark.parametrize("periodic", [True, False])
@pytest.mark.parametrize("cells", [1, 0])
def test_contacts_medium_vs_c_oracle(precision, tol, periodic, cells):
    rng = np.random.default_rng(31)
    n, box = 6000, 3.9
    x = lattice_frames(rng, 6, n, box)
    system = unit_system(n)
    cv = cvpack.NumberOfContacts(range(0, 3200), range(2800, n), plain_nonbonded(n, periodic),
                                 thresholdDistance=0.6)
    cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, x, [box] * 3 if periodic else None, precision=precision)
    ctx._native_cv(cv).set_option("cells", cells)  # pylint: disable=protected-access
    value, grad = cv.getValueAndGradient(ctx)
    stored = ctx.getPositions().double().cpu().numpy()
    ref_value, ref_grad = c_oracle.contacts_batch(
        stored, range(0, 3200), range(2800, n), (), np.diag([box] * 3) if periodic else None,
        r0=0.6, rc=1.2, rs=0.9,
    )
    v = value._value.double().cpu().numpy()
    g = grad.double().cpu().numpy()
    assert np.abs(v - ref_value).max() <= tol * np.abs(ref_value).max()
    assert np.abs(g - ref_grad).max() <= tol * np.abs(ref_grad).max()


@pytest.mark.parametrize("precision,tol", [("single", 1e-5), ("double", 1e-10)])
def test_rmsd_and_rg_medium_vs_c_oracle(precision, tol):
    rng = np.random.default_rng(32)
    n = 40000
    reference = 3.0 * rng.normal(size=(n, 3))
    from scipy.spatial.transform import Rotation

    rot = Rotation.random(4, random_state=1).as_matrix()
    x = np.einsum("bij,nj->bni", rot, reference) + 0.05 * rng.normal(size=(4, n, 3)) + 2.0
    system = helpers.make_system(n, rng)
    rmsd = cvpack.RMSD(reference, range(n), n)
    rg = cvpack.RadiusOfGyration(range(n))
    rg_mass = cvpack.RadiusOfGyrationSq(range(n), weighByMass=True)
    part = cvpack.RMSD(reference, range(4000, 30000), n)  # contiguous, not starting at 0
    for cv in (rmsd, rg, rg_mass, part):
        cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, x, precision=precision)
    stored = ctx.getPositions().double().cpu().numpy()
    masses = system.getMasses()
    cases = [
        (rmsd, c_oracle.rmsd_batch(stored, np.arange(n), reference)),
        (part, c_oracle.rmsd_batch(stored, np.arange(4000, 30000), reference[4000:30000])),
        (rg, c_oracle.gyration_batch(stored, np.arange(n))),
        (rg_mass, c_oracle.gyration_batch(stored, np.arange(n), masses / masses.sum(), squared=True)),
    ]
    for cv, (ref_value, ref_grad) in cases:
        value, grad = cv.getValueAndGradient(ctx)
        v = value._value.double().cpu().numpy()
        g = grad.double().cpu().numpy()
        assert np.abs(v - ref_value).max() <= tol * np.abs(ref_value).max(), cv.getName()
        assert np.abs(g - ref_grad).max() <= tol * np.abs(ref_grad).max(), cv.getName()


def test_contacts_full_size_properties():
    """BASELINE config 2 shape (10k x 10k atoms, L = 5.844 nm), 8 frames."""
    rng = np.random.default_rng(33)
    n, box = 20000, 5.844
    x = lattice_frames(rng, 8, n, box).astype(np.float32)
    system = unit_system(n)
    nb = plain_nonbonded(n)
    whole = cvpack.NumberOfContacts(range(0, 10000), range(10000, n), nb, thresholdDistance=0.6)
    half_a = cvpack.NumberOfContacts(range(0, 5000), range(10000, n), nb, thresholdDistance=0.6)
    half_b = cvpack.NumberOfContacts(range(5000, 10000), range(10000, n), nb, thresholdDistance=0.6)
    for cv in (whole, half_a, half_b):
        cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, x, [box] * 3)
    v, g = whole.getValueAndGradient(ctx)
    va, ga = half_a.getValueAndGradient(ctx)
    vb, gb = half_b.getValueAndGradient(ctx)
    v, va, vb = (t._value.double() for t in (v, va, vb))
    # linearity in the first group
    assert torch.allclose(v, va + vb, rtol=2e-6)
    assert (g - (ga + gb)).abs().max() <= 2e-5 * g.abs().max()
    # a pair sum exerts no net force
    assert g.double().sum(dim=1).abs().max() <= 1e-3 * g.abs().max()
    # translating every atom by a lattice vector, or by anything under PBC, changes nothing
    shifted = x + np.array([box, -2 * box, 0.37], dtype=np.float32)
    ctx.setPositions(shifted)
    v2 = whole.getValue(ctx)._value.double()
    assert torch.allclose(v, v2, rtol=2e-5)
    # frames are independent: reversing the batch reverses the results, bit for bit
    ctx.setPositions(x[::-1].copy())
    v3, g3 = whole.getValueAndGradient(ctx)
    assert torch.equal(v3._value.flip(0).double(), v)
    assert torch.equal(g3.flip(0), g)
    # all-pairs and cell-sorted paths agree
    ctx.setPositions(x)
    ctx._native_cv(whole).set_option("cells", 0)  # pylint: disable=protected-access
    v4, g4 = whole.getValueAndGradient(ctx)
    assert torch.allclose(v4._value.double(), v, rtol=2e-6)
    assert (g4 - g).abs().max() <= 1e-5 * g.abs().max()


def test_rmsd_full_size_properties():
    """BASELINE config 3 shape (100k atoms), 6 frames: rigid motions give RMSD 0, noise gives sigma*sqrt(3)."""
    rng = np.random.default_rng(34)
    n = 100000
    reference = (3.0 * rng.normal(size=(n, 3))).astype(np.float32)
    from scipy.spatial.transform import Rotation

    rot = Rotation.random(6, random_state=2).as_matrix()
    sigma = 0.05
    noise = sigma * rng.normal(size=(6, n, 3))
    noise[0] = 0.0  # frame 0: pure rigid motion
    x = np.einsum("bij,bnj->bni", rot, reference[None] + noise) + rng.normal(size=(6, 1, 3))
    system = unit_system(n)
    rmsd = cvpack.RMSD(reference.astype(np.float64), range(n), n)
    rg = cvpack.RadiusOfGyration(range(n))
    for cv in (rmsd, rg):
        cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, x)
    value, grad = rmsd.getValueAndGradient(ctx)
    v = value._value.double().cpu().numpy()
    assert v[0] < 2e-3  # FP32 coordinates of a rotated structure: ~1e-7 nm * sqrt terms
    expected = np.sqrt((noise[1:] ** 2).sum(axis=(1, 2)) / n)
    assert np.allclose(v[1:], expected, rtol=2e-3)  # optimal alignment removes a tiny part of the noise
    # |grad|^2 summed over atoms = 1/n for an RMSD
    norm = (grad[1:].double() ** 2).sum(dim=(1, 2)).cpu().numpy()
    assert np.allclose(norm, 1.0 / n, rtol=1e-4)
    rgv = rg.getValue(ctx)._value.double().cpu().numpy()
    stored = ctx.getPositions().double().cpu().numpy()
    ref_rg, _ = c_oracle.gyration_batch(stored, np.arange(n), need_grad=False)
    assert np.allclose(rgv, ref_rg, rtol=1e-6)


@pytest.mark.parametrize(
    "n,first,frames,sub,lag",
    [
        (20000, 0, 150, 0, 3),      # planner's choice; 150 frames is not a multiple of a round
        (20000, 0, 150, 256, 1),    # smallest sub-slices, shortest lag
        (6000, 2000, 700, 256, 2),  # contiguous part of the frame (zero gradient elsewhere)
        (3076, 4, 333, 1024, 5),    # ragged last sub-slice and chunk
    ],
)
def test_fused_streaming_kernel_vs_c_oracle(n, first, frames, sub, lag):
    """The one-read persistent kernel (stream_fused.cuh) behind RMSD and RadiusOfGyration: batches
    of at least 64 frames of a contiguous group take it; compared with the C oracle and, for the
    same inputs, with the two-pass kernels."""
    rng = np.random.default_rng(35)
    total = first + n + 8
    reference = 2.0 * rng.normal(size=(total, 3))
    from scipy.spatial.transform import Rotation

    rot = Rotation.random(frames, random_state=3).as_matrix()
    x = np.einsum("bij,nj->bni", rot, reference) + 0.05 * rng.normal(size=(frames, total, 3))
    x += rng.normal(size=(frames, 1, 3))
    system = helpers.make_system(total, rng)
    atoms = np.arange(first, first + n)
    rmsd = cvpack.RMSD(reference, atoms, total)
    rg = cvpack.RadiusOfGyration(atoms)
    rg_mass = cvpack.RadiusOfGyrationSq(atoms, weighByMass=True)
    for cv in (rmsd, rg, rg_mass):
        cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, x)
    stored = ctx.getPositions().double().cpu().numpy()
    masses = system.getMasses()[atoms]
    cases = [
        (rmsd, "rmsd_fused", c_oracle.rmsd_batch(stored, atoms, reference[atoms])),
        (rg, "gyration_fused", c_oracle.gyration_batch(stored, atoms)),
        (rg_mass, "gyration_fused",
         c_oracle.gyration_batch(stored, atoms, masses / masses.sum(), squared=True)),
    ]
    for cv, kernel, (ref_value, ref_grad) in cases:
        native = ctx._native_cv(cv)  # pylint: disable=protected-access
        native.set_option("fused_slice", sub)
        native.set_option("fused_lag", lag)
        native.set_option("time_kernels", 1)
        value, grad = cv.getValueAndGradient(ctx)
        assert native.kernel_time(kernel)[1] == 1, "the fused kernel did not run"
        v = value._value.double().cpu().numpy()
        g = grad.double().cpu().numpy()
        assert np.abs(v - ref_value).max() <= 1e-5 * np.abs(ref_value).max(), cv.getName()
        assert np.abs(g - ref_grad).max() <= 1e-5 * np.abs(ref_grad).max(), cv.getName()
        outside = np.ones(total, dtype=bool)
        outside[atoms] = False
        assert not g[:, outside].any()
        # value only (no gradient items, no lag)
        v_only = cv.getValue(ctx)._value
        assert torch.equal(v_only, value._value)
        assert native.kernel_time(kernel)[1] == 1
        # the same through the two-pass kernels
        native.set_option("fused", 0)
        value2, grad2 = cv.getValueAndGradient(ctx)
        assert native.kernel_time(kernel)[1] == 0
        native.set_option("fused", 1)
        assert torch.allclose(value2._value, value._value, rtol=1e-6)
        assert (grad2 - grad).abs().max() <= 5e-6 * grad.abs().max()
        # deterministic
        value3, grad3 = cv.getValueAndGradient(ctx)
        assert torch.equal(value3._value, value._value) and torch.equal(grad3, grad)


@pytest.mark.parametrize("periodic", [True, False])
def test_symmetric_sweep_vs_c_oracle(periodic):
    """Single symmetric sweep (fixed-point j accumulators in shared memory, pair_cells.cuh): taken
    for FP32 contacts between disjoint groups when one CTA handles a frame.  Compared with the C
    oracle and with the two-sweep path; bitwise reproducible."""
    rng = np.random.default_rng(36)
    n, box = 9000, 4.48
    x = lattice_frames(rng, 5, n, box)
    system = unit_system(n)
    cv = cvpack.NumberOfContacts(range(0, 5000), range(5000, n), plain_nonbonded(n, periodic),
                                 thresholdDistance=0.6)
    cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, x, [box] * 3 if periodic else None)
    native = ctx._native_cv(cv)  # pylint: disable=protected-access
    native.set_option("nsplit", 1)
    native.set_option("time_kernels", 1)
    value, grad = cv.getValueAndGradient(ctx)
    assert native.kernel_time("sweep")[1] == 1, "expected one (symmetric) sweep"
    stored = ctx.getPositions().double().cpu().numpy()
    ref_value, ref_grad = c_oracle.contacts_batch(
        stored, range(0, 5000), range(5000, n), (), np.diag([box] * 3) if periodic else None,
        r0=0.6, rc=1.2, rs=0.9,
    )
    v = value._value.double().cpu().numpy()
    g = grad.double().cpu().numpy()
    assert np.abs(v - ref_value).max() <= 1e-5 * np.abs(ref_value).max()
    assert np.abs(g - ref_grad).max() <= 1e-5 * np.abs(ref_grad).max()
    # a pair sum exerts no net force
    assert np.abs(g.sum(axis=1)).max() <= 1e-3 * np.abs(g).max()
    value_again, grad_again = cv.getValueAndGradient(ctx)
    assert torch.equal(value_again._value, value._value) and torch.equal(grad_again, grad)
    native.kernel_time("sweep")
    native.set_option("sym", 0)
    value2, grad2 = cv.getValueAndGradient(ctx)
    assert native.kernel_time("sweep")[1] == 2
    assert torch.allclose(value2._value, value._value, rtol=1e-6)
    assert (grad2 - grad).abs().max() <= 5e-6 * grad.abs().max()


@pytest.mark.parametrize(
    "n1,n2,n,box,frames",
    # group sizes that leave ragged chunks and tiles of the dealt layout (CellLayout): one atom past
    # a chunk, a prime, a group smaller than a tile; the 6.2 nm box with 2 x 2.5k of 23k atoms is
    # configs[4]: sparse groups, wide tiles, some (tile, cluster) pairs in the per-pair
    # minimum-image segment; 2.45 nm is barely more than twice the cutoff: most pairs are
    [(1537, 3089, 6000, 3.92, 3), (2500, 2500, 23000, 6.2, 2), (29, 1901, 2500, 2.93, 3), (700, 800, 1500, 2.45, 3)],
)
@pytest.mark.parametrize("path", ["sym", "two_sweeps"])
def test_sweep_layouts_vs_c_oracle(n1, n2, n, box, frames, path):
    """Both sweep paths on scattered groups of awkward sizes, against the C oracle."""
    rng = np.random.default_rng(41)
    x = lattice_frames(rng, frames, n, box)
    system = unit_system(n)
    members = rng.permutation(n)
    g1, g2 = sorted(members[:n1].tolist()), sorted(members[n1:n1 + n2].tolist())
    cv = cvpack.NumberOfContacts(g1, g2, plain_nonbonded(n, True), thresholdDistance=0.6)
    cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, x, [box] * 3)
    native = ctx._native_cv(cv)  # pylint: disable=protected-access
    native.set_option("time_kernels", 1)
    if path == "sym":
        native.set_option("nsplit", 1)
    else:
        native.set_option("sym", 0)
    value, grad = cv.getValueAndGradient(ctx)
    assert native.kernel_time("sweep")[1] == (1 if path == "sym" else 2)
    assert int(native.stat("last_path")) == (2 if path == "sym" else 1)
    stored = ctx.getPositions().double().cpu().numpy()
    ref_value, ref_grad = c_oracle.contacts_batch(stored, g1, g2, (), np.diag([box] * 3), r0=0.6, rc=1.2, rs=0.9)
    v = value._value.double().cpu().numpy()
    g = grad.double().cpu().numpy()
    assert np.abs(v - ref_value).max() <= 1e-5 * np.abs(ref_value).max()
    assert np.abs(g - ref_grad).max() <= 1e-5 * np.abs(ref_grad).max()
    assert not g[:, sorted(set(range(n)) - set(g1) - set(g2))].any()
    value_again, grad_again = cv.getValueAndGradient(ctx)
    assert torch.equal(value_again._value, value._value) and torch.equal(grad_again, grad)
    # CVP_FLAG_ACCUMULATE through the raw C ABI (the symmetric sweep adds its gradients in place)
    from cvpack_b200 import _native  # pylint: disable=import-outside-toplevel

    pos = ctx.getPositions()
    box9 = torch.tensor(np.diag([box] * 3).ravel(), dtype=torch.float32, device=pos.device)
    base_value = torch.full_like(value._value, 0.25)
    base_grad = torch.full_like(grad, -0.5)
    acc_value, acc_grad = base_value.clone(), base_grad.clone()
    native.evaluate(_native.CVP_F32, pos.data_ptr(), box9.data_ptr(), _native.CVP_BOX_SHARED, frames, n,
                    acc_value.data_ptr(), acc_grad.data_ptr(), _native.CVP_FLAG_ACCUMULATE,
                    torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.allclose(acc_value, base_value + value._value, rtol=1e-6, atol=1e-6)
    assert torch.allclose(acc_grad, base_grad + grad, rtol=0, atol=1e-6 * float(grad.abs().max()) + 1e-7)


@pytest.mark.parametrize("metric", ["progress", "deviation"])
@pytest.mark.parametrize(
    "n_group,first,frames,milestones,tail",
    # (129 = one atom past a tile; the last two shapes have every row of a tile on 16-byte
    # boundaries and take the bulk-copy rows of the gradient kernel, 296 = 4 tiles + 40 atoms)
    [(300, 8, 70, 5, 5), (1000, 0, 33, 37, 5), (129, 3, 40, 9, 5), (257, 0, 65, 3, 5), (296, 8, 70, 5, 4),
     (128, 4, 33, 9, 0)],
)
def test_path_many_references_one_atom_set(metric, n_group, first, frames, milestones, tail):
    """PathInRMSDSpace whose milestones all list the same contiguous atoms takes the tiled
    double-precision contractions of rmsd_multi.cuh: against the NumPy oracle and against the
    generic per-entry kernels."""
    rng = np.random.default_rng(37)
    n = first + n_group + tail
    system = helpers.make_system(n, rng)
    start, end = rng.normal(size=(n, 3)), rng.normal(size=(n, 3))
    atoms = list(range(first, first + n_group))
    refs = []
    for k in range(milestones):
        conf = start + (end - start) * k / (milestones - 1) + 0.01 * rng.normal(size=(n, 3))
        refs.append({a: conf[a] for a in atoms})
    lam = rng.uniform(0.1, 0.9, size=frames)[:, None, None]
    x = start[None] + (end - start)[None] * lam + 0.02 * rng.normal(size=(frames, n, 3))
    cv = cvpack.PathInRMSDSpace(getattr(cvpack.path, metric), refs, n, 0.3)
    cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, x)
    native = ctx._native_cv(cv)  # pylint: disable=protected-access
    native.set_option("time_kernels", 1)
    helpers.check_against_oracle(cv, system, x, None, "single", 1e-5, 1e-5, context=ctx)
    assert native.kernel_time("rmsd_multi_cov")[1] >= 1, "the multi-reference kernels did not run"
    value, grad = cv.getValueAndGradient(ctx)
    native.set_option("multi", 0)
    value2, grad2 = cv.getValueAndGradient(ctx)
    native.set_option("multi", 1)
    assert torch.allclose(value2._value, value._value, rtol=2e-6, atol=1e-7)
    assert (grad2 - grad).abs().max() <= 2e-6 * grad.abs().max()
    # the FP64 tensor-core (DMMA) kernels are the default; the register-tiled DFMA kernels they
    # replace must agree to the rounding of a different summation order
    assert native.set_option("multi_mma", 0) == 1
    value3, grad3 = cv.getValueAndGradient(ctx)
    assert torch.allclose(value3._value, value._value, rtol=1e-6, atol=1e-7)
    assert (grad3 - grad).abs().max() <= 1e-6 * grad.abs().max()
    # ... and the register-staged DMMA gradient sums in the same order as the TMA-fed default
    native.set_option("multi_mma", 2)
    value5, grad5 = cv.getValueAndGradient(ctx)
    assert torch.equal(value5._value, value._value) and torch.equal(grad5, grad)
    # ... and so does the TMA-fed covariance kernel (option 3: kept, not the default)
    native.set_option("multi_mma", 3)
    value6, grad6 = cv.getValueAndGradient(ctx)
    native.set_option("multi_mma", 1)
    assert torch.allclose(value6._value, value._value, rtol=1e-6, atol=1e-7)
    assert (grad6 - grad).abs().max() <= 1e-6 * grad.abs().max()
    value4, grad4 = cv.getValueAndGradient(ctx)
    assert torch.equal(value4._value, value._value) and torch.equal(grad4, grad), "not deterministic"
    outside = torch.ones(n, dtype=torch.bool)
    outside[atoms] = False
    assert not grad[:, outside].any()
    # CVP_FLAG_ACCUMULATE through the raw C ABI: value and gradient are added to what is there
    from cvpack_b200 import _native  # pylint: disable=import-outside-toplevel

    pos = ctx.getPositions()
    base_value = torch.full_like(value._value, 0.25)
    base_grad = torch.full_like(grad, -0.5)
    acc_value, acc_grad = base_value.clone(), base_grad.clone()
    native.evaluate(_native.CVP_F32, pos.data_ptr(), None, _native.CVP_BOX_NONE, frames, n, acc_value.data_ptr(),
                    acc_grad.data_ptr(), _native.CVP_FLAG_ACCUMULATE, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.allclose(acc_value, base_value + value._value, rtol=1e-6, atol=1e-7)
    expected = base_grad + grad
    expected[:, outside] = -0.5
    assert torch.equal(acc_grad, expected)


@pytest.mark.parametrize("frames", [4, 80])
def test_rmsd_of_nearly_identical_frames(frames):
    """RMSD close to zero on the streaming paths (4 frames: two-pass kernels, 80: fused kernel).  The
    covariance is accumulated against the reference in double: a float copy is no longer centred, and
    sum x (x) y^ then carries c (x) sum y^ -- 1e-9 in the mean square deviation, which is 100 % of an
    RMSD of 3e-5 nm."""
    rng = np.random.default_rng(38)
    n = 4096
    reference = rng.uniform(0.0, 2.5, size=(n, 3))
    scales = np.logspace(-6, -2, frames)[:, None, None]
    x = reference[None] + scales * rng.normal(size=(frames, n, 3))
    x[0] = reference
    system = unit_system(n)
    rmsd = cvpack.RMSD(reference, range(n), n)
    rmsd.addToSystem(system)
    ctx = cvpack.BatchContext(system, x)
    stored = ctx.getPositions().double().cpu().numpy()
    ref_value, ref_grad = c_oracle.rmsd_batch(stored, np.arange(n), reference)
    value, grad = rmsd.getValueAndGradient(ctx)
    v = value._value.double().cpu().numpy()
    # 1e-5 relative, with an absolute floor of 1e-7 nm: the RMSD of the reference frame itself is
    # rounding noise of (Gx + Gy - 2 lambda) ~ 1e-13 in both implementations (3e-8 vs 7e-8 nm)
    assert np.all(np.abs(v - ref_value) <= 1e-5 * ref_value + 1e-7), np.abs(v - ref_value).max()
    g = grad.double().cpu().numpy()
    big = ref_value > 1e-4  # the gradient direction of a vanishing RMSD is ill-conditioned in FP32
    err = np.abs(g[big] - ref_grad[big]).max(axis=(1, 2)) / np.abs(ref_grad[big]).max(axis=(1, 2))
    assert err.max() <= 2e-3


def test_symmetric_sweep_dilute_frames():
    """Frames whose gradients are tiny: every group-1 atom has one partner of group 2 at 0.12-0.25 nm
    (far inside r0 = 0.6 nm, where dS/dr ~ x^5 is 1e-3 .. 0.13) and nothing else within the cutoff.
    The j side of the symmetric sweep is accumulated in fixed point: its quantum must be negligible
    against *these* gradients too -- the error is measured relative to the largest component of each
    frame, like everywhere else -- while the sums of a dense frame in the same batch must survive the
    wrap of the 32-bit accumulators."""
    rng = np.random.default_rng(39)
    m, spacing = 6, 1.45
    sites = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3) * spacing
    n1 = len(sites)
    box, frames, n = m * spacing, 150, 2 * len(sites)
    x = np.empty((frames, n, 3))
    x[:, :n1] = sites[None] + rng.uniform(0, 0.05, size=(frames, n1, 3))
    offset = rng.normal(size=(frames, n1, 3))
    offset *= rng.uniform(0.12, 0.25, size=(frames, n1, 1)) / np.linalg.norm(offset, axis=2, keepdims=True)
    x[:, n1:] = x[:, :n1] + offset
    x[7] = rng.uniform(0.0, 1.5, size=(n, 3))   # one dense frame: a hundred contacts per atom
    system = unit_system(n)
    cv = cvpack.NumberOfContacts(range(0, n1), range(n1, n), plain_nonbonded(n, True), thresholdDistance=0.6)
    cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, x, [box] * 3)
    native = ctx._native_cv(cv)  # pylint: disable=protected-access
    value, grad = cv.getValueAndGradient(ctx)
    assert int(native.stat("last_path")) == 2, "expected the symmetric sweep"
    stored = ctx.getPositions().double().cpu().numpy()
    ref_value, ref_grad = c_oracle.contacts_batch(
        stored, range(0, n1), range(n1, n), (), np.diag([box] * 3), r0=0.6, rc=1.2, rs=0.9, cells=True,
    )
    v, g = value._value.double().cpu().numpy(), grad.double().cpu().numpy()
    largest = np.abs(ref_grad).reshape(frames, -1).max(axis=1)
    dilute = np.delete(np.arange(frames), 7)
    assert 0 < largest[dilute].max() < 0.2, largest[dilute].max()
    assert largest[7] > 10.0
    for b in range(frames):
        assert abs(v[b] - ref_value[b]) <= 1e-5 * abs(ref_value[b]), (b, v[b], ref_value[b])
        err = np.abs(g[b] - ref_grad[b]).max() / largest[b]
        assert err <= 1e-5, f"frame {b}: gradient error {err:.2e} of the largest component {largest[b]:.3g}"
        assert np.array_equal(g[b] != 0, ref_grad[b] != 0)
    # Newton's third law holds to the quantum of the accumulators, frame by frame
    assert np.abs(g[dilute].sum(axis=1)).max() <= 1e-5 * largest[dilute].max() * np.sqrt(n)


def test_path_in_rmsd_space_at_config_size():
    """BASELINE configs[3] at full size -- 64 milestones x 20 000 atoms -- on a few frames against the
    NumPy oracle (progress and deviation; the FP64 tensor-core contractions of rmsd_multi.cuh)."""
    from scripts import workloads

    device = torch.device("cuda:0")
    cv, system, x, _, info = workloads.path_rmsd(3, 4, device)
    deviation = cvpack.PathInRMSDSpace(cvpack.path.deviation, cv._args["milestones"], 20000, info["sigma"])  # pylint: disable=protected-access
    deviation.addToSystem(system)
    ctx = cvpack.BatchContext(system, x)
    for variable in (cv, deviation):
        native = ctx._native_cv(variable)  # pylint: disable=protected-access
        native.set_option("time_kernels", 1)
        helpers.check_against_oracle(variable, system, x.double().cpu().numpy(), None, "single", 1e-5, 1e-5, context=ctx)
        assert native.kernel_time("rmsd_multi_cov")[1] >= 1
    # the progress variable recovers where along the path the frames were generated
    s = cv.getValue(ctx)._value.double().cpu().numpy() * 63
    assert np.abs(s - info["generated_position"].double().cpu().numpy()).max() < 0.6


def test_walker_batch_at_config_shape():
    """BASELINE configs[4] shape: 23 000-atom walkers, NumberOfContacts 2 500 x 2 500 (periodic) and a
    two-body CompositeRMSD (1 500 + 1 000 atoms), against the oracles."""
    from scripts import workloads

    device = torch.device("cuda:0")
    cvs, system, x, box, _ = workloads.walkers(6, 8, device)
    ctx = cvpack.BatchContext(system, x, box)
    stored = ctx.getPositions().double().cpu().numpy()
    value, grad = cvs["contacts"].getValueAndGradient(ctx)
    ref_value, ref_grad = c_oracle.contacts_batch(stored, range(0, 2500), range(2500, 5000), (), np.diag(box),
                                                  r0=0.6, rc=1.2, rs=0.9, cells=True)
    v, g = value._value.double().cpu().numpy(), grad.double().cpu().numpy()
    assert np.abs(v - ref_value).max() <= 1e-5 * np.abs(ref_value).max()
    assert np.abs(g - ref_grad).max() <= 1e-5 * np.abs(ref_grad).max()
    assert not g[:, 5000:].any()
    helpers.check_against_oracle(cvs["composite"], system, stored, np.diag(box), "single", 1e-5, 1e-5, context=ctx)
