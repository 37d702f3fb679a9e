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
