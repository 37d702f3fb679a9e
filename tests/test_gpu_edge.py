"""Edge cases of the GPU path: empty and single-frame batches, ragged sizes that miss the fast paths,
forced chunking, per-frame boxes, host residency, infinite effective mass, the raw C ABI."""

import ctypes

import numpy as np
import pytest
import torch

import cvpack_b200 as cvpack
from cvpack_b200 import _native, mmunit
from oracle import c_oracle
from oracle import cv_oracle as oracle
from tests import helpers

pytestmark = pytest.mark.gpu


def test_empty_and_single_frame_batches():
    rng = np.random.default_rng(41)
    n = 50
    system = helpers.make_system(n, rng)
    cv = cvpack.RadiusOfGyration(range(n))
    cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, np.zeros((0, n, 3)))
    value, grad = cv.getValueAndGradient(ctx)
    assert value._value.shape == (0,) and grad.shape == (0, n, 3)
    x = rng.normal(size=(n, 3))
    ctx.setPositions(x)  # [N, 3] -> one frame
    value = cv.getValue(ctx)
    assert value._value.shape == (1,)
    assert float(value._value[0]) == pytest.approx(oracle.radius_of_gyration(x.astype(np.float32).astype(float), range(n))[0], rel=1e-6)


@pytest.mark.parametrize("n,first", [(1001, 0), (1000, 3), (7, 0), (1, 0)])
def test_ragged_sizes_fall_back_to_generic_kernels(n, first):
    """Counts not divisible by 4 / unaligned starts cannot use the float4 streaming path."""
    rng = np.random.default_rng(42)
    total = n + first + 5
    system = helpers.make_system(total, rng)
    x = rng.normal(size=(3, total, 3)) * 2.0
    group = range(first, first + n)
    reference = rng.normal(size=(total, 3)) * 2.0
    cvs = [cvpack.RadiusOfGyrationSq(group)]
    if n >= 3:
        cvs.append(cvpack.RMSD(reference, group, total))
    for cv in cvs:
        helpers.check_against_oracle(cv, system, x, None, "single", 1e-5, 1e-5)


def test_forced_chunking_gives_identical_results():
    rng = np.random.default_rng(43)
    n = 900
    box = np.diag([2.6, 2.6, 2.6])
    system = helpers.make_system(n, rng)
    x = rng.uniform(0, 2.6, size=(23, n, 3))
    nb = helpers.make_nonbonded(n, rng, periodic=True, num_exceptions=30)
    cv = cvpack.NumberOfContacts(range(0, 450), range(300, 900), nb)
    cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, x, box)
    v1, g1 = cv.getValueAndGradient(ctx)
    ctx._native_cv(cv).set_workspace_limit(200_000)  # a handful of frames per chunk
    v2, g2 = cv.getValueAndGradient(ctx)
    assert torch.equal(v1._value, v2._value) and torch.equal(g1, g2)


def test_per_frame_boxes_device_and_host():
    rng = np.random.default_rng(44)
    n, frames = 400, 9
    system = helpers.make_system(n, rng)
    x = rng.uniform(0, 3.0, size=(frames, n, 3))
    edges = rng.uniform(2.4, 3.0, size=(frames, 3))
    nb = helpers.make_nonbonded(n, rng, periodic=True)
    contacts = cvpack.NumberOfContacts(range(0, 200), range(200, 400), nb)
    rg = cvpack.RadiusOfGyration(range(100, 300), pbc=True)
    for cv in (contacts, rg):
        cv.addToSystem(system)
    dev = cvpack.BatchContext(system, x, edges)
    host = cvpack.BatchContext(system, x, edges, residency="host")
    stored = dev.getPositions().double().cpu().numpy()
    for cv in (contacts, rg):
        vd, gd = cv.getValueAndGradient(dev)
        vh, gh = cv.getValueAndGradient(host)
        assert torch.equal(vd._value.cpu(), vh._value) and torch.equal(gd.cpu(), gh)
        for b in range(frames):
            box = np.diag(edges[b].astype(np.float32).astype(np.float64))
            ref_value, ref_grad = helpers.oracle_evaluate(cv, system, stored[b], box)
            assert float(vd._value[b]) == pytest.approx(ref_value, rel=1e-5)
            assert np.abs(gd[b].double().cpu().numpy() - ref_grad).max() <= 1e-5 * max(np.abs(ref_grad).max(), 1e-6)


def test_evaluate_into_and_wrong_buffers():
    rng = np.random.default_rng(45)
    n = 64
    system = helpers.make_system(n, rng)
    x = rng.normal(size=(5, n, 3))
    cv = cvpack.Torsion(0, 1, 2, 3)
    cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, x)
    value = torch.empty(5, dtype=torch.float32, device="cuda")
    grad = torch.empty((5, n, 3), dtype=torch.float32, device="cuda")
    ctx.evaluateInto(cv, value, grad)
    v, g = cv.getValueAndGradient(ctx)
    assert torch.equal(value, v._value) and torch.equal(grad, g)
    with pytest.raises(ValueError):
        ctx.evaluateInto(cv, torch.empty(4, dtype=torch.float32, device="cuda"))
    with pytest.raises(ValueError):
        ctx.evaluateInto(cv, torch.empty(5, dtype=torch.float64, device="cuda"))


def test_effective_mass_is_infinite_without_gradient():
    """A Heaviside step function has zero gradient everywhere -> inf (utils.py:181-182)."""
    rng = np.random.default_rng(46)
    n = 120
    system = helpers.make_system(n, rng)
    x = rng.uniform(0, 1.5, size=(3, n, 3))
    nb = helpers.make_nonbonded(n, rng, periodic=False)
    cv = cvpack.NumberOfContacts(range(60), range(60, 120), nb, stepFunction="step(1-x)", switchFactor=None)
    cv.addToSystem(system)
    ctx = cvpack.BatchContext(system, x)
    emass = cv.getEffectiveMass(ctx)
    assert torch.isinf(emass._value).all()
    assert emass.unit.get_symbol() == "nm**2 Da"


def test_raw_c_abi_on_device_pointers():
    """cvp_rmsd_create + cvp_evaluate + cvp_effective_mass called directly, no Python CV classes."""
    rng = np.random.default_rng(47)
    n, frames = 256, 6
    reference = rng.normal(size=(n, 3))
    x = (reference[None] + 0.1 * rng.normal(size=(frames, n, 3))).astype(np.float32)
    offsets, atoms = _native.as_i32([0, n]), _native.as_i32(np.arange(n))
    ref = _native.as_f64(reference)
    desc = _native.RmsdDesc(num_atoms=n, num_groups=1, group_offsets=_native.ptr_i32(offsets),
                            atoms=_native.ptr_i32(atoms), reference=_native.ptr_f64(ref),
                            combine=_native.CVP_COMBINE_IDENTITY)
    handle = ctypes.c_void_p()
    lib = _native.load()
    assert lib.cvp_rmsd_create(ctypes.byref(desc), ctypes.byref(handle)) == 0
    pos = torch.from_numpy(x).cuda()
    value = torch.empty(frames, device="cuda")
    grad = torch.empty_like(pos)
    status = lib.cvp_evaluate(handle, _native.CVP_F32, pos.data_ptr(), None, 0, frames, n,
                              value.data_ptr(), grad.data_ptr(), 0, None)
    assert status == 0, lib.cvp_last_error()
    torch.cuda.synchronize()
    ref_value, ref_grad = c_oracle.rmsd_batch(x.astype(np.float64), np.arange(n), reference)
    assert np.allclose(value.cpu().numpy(), ref_value, rtol=1e-5)
    assert np.abs(grad.cpu().numpy() - ref_grad).max() <= 1e-5 * np.abs(ref_grad).max()
    masses = torch.full((n,), 2.0, dtype=torch.float64, device="cuda")
    emass = torch.empty(frames, device="cuda")
    assert lib.cvp_effective_mass(_native.CVP_F32, grad.data_ptr(), masses.data_ptr(), frames, n,
                                  emass.data_ptr(), None) == 0
    expected = [oracle.effective_mass(ref_grad[b], np.full(n, 2.0)) for b in range(frames)]
    assert np.allclose(emass.cpu().numpy(), expected, rtol=1e-4)
    # wrong atom count is rejected with a message, not a crash
    assert lib.cvp_evaluate(handle, _native.CVP_F32, pos.data_ptr(), None, 0, frames, n + 1,
                            value.data_ptr(), None, 0, None) == -1
    assert b"atoms per frame" in lib.cvp_last_error()
    assert lib.cvp_destroy(handle) == 0


def test_quantity_inputs_and_units():
    rng = np.random.default_rng(48)
    n = 30
    system = cvpack.System()
    for _ in range(n):
        system.addParticle(12.0 * mmunit.daltons)
    x_nm = rng.normal(size=(2, n, 3))
    cv = cvpack.Distance(0, 1)
    cv.addToSystem(system)
    ctx_nm = cvpack.BatchContext(system, x_nm)
    ctx_angstrom = cvpack.BatchContext(system, (10.0 * x_nm) * mmunit.angstroms)
    a = cv.getValue(ctx_nm)
    b = cv.getValue(ctx_angstrom)
    assert torch.allclose(a._value, b._value, rtol=1e-6)
    assert a.unit.get_symbol() == "nm"
    assert torch.allclose(a.value_in_unit(mmunit.angstroms), 10 * a._value)


# ---- round 2: advisor findings and boundary hygiene ------------------------------------------------
@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("separation", [0.86, 0.93, 0.995, 1.2])
def test_shortest_distance_near_and_beyond_the_cutoff(precision, separation):
    """With beta = 100 the weights e^{-beta r/rc} fall below the FP32 range once r > 0.887 rc and
    1/D above it: channels and quotient-rule coefficients are carried in double.  At 1.2 rc no pair
    is inside the cutoff: the value is rc exactly and the gradient vanishes."""
    rng = np.random.default_rng(61)
    n = 80
    system = helpers.make_system(n, rng)
    x = rng.uniform(0.0, 0.3, size=(4, n, 3))
    x[:, 40:, 0] += 0.3 + separation  # group 2 sits `separation` nm beyond the far face of group 1
    cv = cvpack.ShortestDistance(range(40), range(40, 80), n, pbc=False)
    vt, gt = (1e-5, 1e-5) if precision == "single" else (1e-10, 1e-10)
    ctx = helpers.check_against_oracle(cv, system, x, None, precision, vt, gt)
    value, grad = cv.getValueAndGradient(ctx)
    assert torch.isfinite(value._value).all() and torch.isfinite(grad).all()
    if separation > 1.0:
        assert torch.all(value._value == 1.0) and not grad.any()


def test_coincident_atoms_and_shared_residue_give_S0_not_nan():
    """rsqrt(0)*0 is NaN: two atoms on top of each other count S(0) = 1 with no force, and a residue
    listed in both groups of ResidueCoordination contributes S(0) (the reference evaluates it so)."""
    rng = np.random.default_rng(62)
    n = 64
    box = np.diag([2.0, 2.0, 2.0])
    system = helpers.make_system(n, rng, box)
    x = rng.uniform(0, 2.0, size=(3, n, 3))
    x[:, 40] = x[:, 3]  # atom 40 (group 2) coincides with atom 3 (group 1)
    nb = helpers.make_nonbonded(n, rng, periodic=True)
    for kwargs in ({}, {"stepFunction": "1/(1+x^4)"}):
        cv = cvpack.NumberOfContacts(range(32), range(32, 64), nb, **kwargs)
        helpers.check_against_oracle(cv, system, x, box, "single", 1e-5, 1e-5)
    topology = helpers.make_peptide(8)
    residues = list(topology.residues())
    total = topology.getNumAtoms()
    system = helpers.make_system(total, rng)
    y = rng.uniform(0, 1.5, size=(2, total, 3))
    cv = cvpack.ResidueCoordination(residues[:5], residues[4:], pbc=False, thresholdDistance=0.8)
    ctx = helpers.check_against_oracle(cv, system, y, None, "single", 1e-5, 1e-5)
    assert torch.isfinite(cv.getValue(ctx)._value).all()


@pytest.mark.parametrize("offset_bytes", [4, 8, 12])
def test_misaligned_buffers_take_the_generic_kernels(offset_bytes):
    """A tensor view whose storage offset is not a multiple of 16 bytes must not reach the float4 /
    TMA streaming kernels (sticky misaligned-address fault): same values from the generic path."""
    rng = np.random.default_rng(63)
    n, frames = 4096, 70
    reference = rng.normal(size=(n, 3))
    x = (reference[None] + 0.1 * rng.normal(size=(frames, n, 3))).astype(np.float32)
    system = helpers.make_system(n, rng)
    lib = _native.load()
    shift = offset_bytes // 4
    for cv in (cvpack.RMSD(reference, range(n), n), cvpack.RadiusOfGyration(range(n))):
        cv.addToSystem(system)
        ctx = cvpack.BatchContext(system, x)
        value, grad = cv.getValueAndGradient(ctx)
        native = ctx._native_cv(cv)
        flat_pos = torch.empty(x.size + 4, dtype=torch.float32, device="cuda")
        flat_grad = torch.empty(x.size + 4, dtype=torch.float32, device="cuda")
        pos = flat_pos[shift:shift + x.size].view(frames, n, 3)
        pos.copy_(torch.from_numpy(x))
        out_grad = flat_grad[shift:shift + x.size].view(frames, n, 3)
        out_value = torch.empty(frames, device="cuda")
        assert pos.data_ptr() % 16 == offset_bytes
        status = lib.cvp_evaluate(native.handle, _native.CVP_F32, pos.data_ptr(), None, 0, frames, n,
                                  out_value.data_ptr(), out_grad.data_ptr(), 0, None)
        assert status == 0, lib.cvp_last_error()
        torch.cuda.synchronize()
        assert torch.allclose(out_value, value._value, rtol=2e-6)
        assert (out_grad - grad).abs().max() <= 2e-6 * grad.abs().max()


def test_cutoff_larger_than_half_the_box_is_refused():
    """OpenMM: 'The cutoff distance cannot be greater than half the periodic box size.'"""
    rng = np.random.default_rng(64)
    n = 100
    system = helpers.make_system(n, rng)
    x = rng.uniform(0, 1.0, size=(2, n, 3))
    nb = helpers.make_nonbonded(n, rng, periodic=True)
    cv = cvpack.NumberOfContacts(range(50), range(50, 100), nb)  # cutoff 0.6 nm
    cv.addToSystem(system)
    for box in (np.diag([1.1, 3.0, 3.0]), np.array([[[3.0, 0, 0], [0, 3.0, 0], [0, 0, 3.0]],
                                                    [[3.0, 0, 0], [0, 1.19, 0], [0, 0, 3.0]]])):
        for residency in ("device", "host"):
            ctx = cvpack.BatchContext(system, x, box, residency=residency)
            with pytest.raises(ValueError, match="cannot be greater than half the periodic box size"):
                cv.getValue(ctx)
    ctx = cvpack.BatchContext(system, x, np.diag([1.2, 1.2, 1.2]))  # exactly twice the cutoff is fine
    assert torch.isfinite(cv.getValue(ctx)._value).all()
    # the C ABI's host path checks its (host) boxes itself
    native = ctx._native_cv(cv)
    pos = torch.from_numpy(x.astype(np.float32)).pin_memory()
    small = torch.tensor(np.diag([1.0, 3.0, 3.0]).ravel(), dtype=torch.float32)
    value = torch.empty(2, dtype=torch.float32).pin_memory()
    lib = _native.load()
    status = lib.cvp_evaluate_host(native.handle, _native.CVP_F32, pos.data_ptr(), small.data_ptr(),
                                   _native.CVP_BOX_SHARED, 2, n, value.data_ptr(), None, 0, 0)
    assert status == -1 and b"half the periodic box size" in lib.cvp_last_error()


@pytest.mark.parametrize("precision", ["single", "double"])
def test_effective_mass_without_materialising_the_gradient(precision):
    """getEffectiveMass goes through cvp_evaluate_with_mass[_host]: same numbers as the reduction of an
    explicitly requested gradient, for both residencies, without a [B,N,3] tensor on the Python side."""
    rng = np.random.default_rng(65)
    n = 300
    box = np.diag([2.4, 2.4, 2.4])
    system = helpers.make_system(n, rng)
    x = rng.uniform(0, 2.4, size=(37, n, 3))
    nb = helpers.make_nonbonded(n, rng, periodic=True, num_exceptions=10)
    cvs = [cvpack.NumberOfContacts(range(150), range(150, 300), nb), cvpack.RadiusOfGyration(range(20, 200)),
           cvpack.Torsion(1, 5, 9, 13)]
    for cv in cvs:
        cv.addToSystem(system)
    masses = system.getMasses()
    for residency in ("device", "host"):
        ctx = cvpack.BatchContext(system, x, box, precision=precision, residency=residency)
        for cv in cvs:
            emass = cv.getEffectiveMass(ctx)
            assert emass._value.shape == (37,) and emass._value.is_cuda == (residency == "device")
            grad = cv.getGradient(ctx).double().cpu().numpy()
            expected = [oracle.effective_mass(grad[b], masses) for b in range(37)]
            assert np.allclose(emass._value.double().cpu().numpy(), expected,
                               rtol=1e-5 if precision == "single" else 1e-12)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two devices")
def test_host_path_restores_the_current_device():
    rng = np.random.default_rng(66)
    n = 40
    system = helpers.make_system(n, rng)
    cv = cvpack.RadiusOfGyration(range(n))
    cv.addToSystem(system)
    torch.cuda.set_device(0)
    ctx = cvpack.BatchContext(system, rng.normal(size=(3, n, 3)), device="cuda:1", residency="host")
    cv.getValue(ctx)
    assert torch.cuda.current_device() == 0
    _native.measure_fp32_peak(1)
    assert torch.cuda.current_device() == 0
