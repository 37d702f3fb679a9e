"""Trajectory ingestion and walker coupling (rank 4): host logic on the CPU,
the streamed evaluation and the in-place bias forces on the GPU against the oracle."""

import io
import struct

import numpy as np
import pytest
import torch

import cvpack_b200 as cvpack
from cvpack_b200 import bias as cvbias
from cvpack_b200.trajectory import DCDFile, NumpyTrajectory, TrajectoryEvaluator, XTCFile, open_trajectory
from tests import helpers, xtc_python


def _frames(rng, frames, atoms, scale=2.0):
    return rng.uniform(-scale, scale, size=(frames, atoms, 3))


# ---- host logic (no GPU) ---------------------------------------------------------------------------
def test_dcd_round_trip_without_box(tmp_path):
    rng = np.random.default_rng(7)
    x = _frames(rng, 5, 13)
    path = tmp_path / "a.dcd"
    DCDFile.write(path, x, firstStep=100, interval=50)
    dcd = DCDFile(path)
    assert (len(dcd), dcd.numAtoms, dcd.hasBox, dcd.firstStep, dcd.interval) == (5, 13, False, 100, 50)
    pos, box = dcd.read(0, 5)
    assert box is None and pos.dtype == np.float32 and pos.shape == (5, 13, 3)
    # stored as float32 Angstrom: exactly the correctly rounded quotient of the stored value
    stored = (x * 10.0).astype(np.float32)
    assert np.array_equal(pos, stored / np.float32(10.0))
    assert np.allclose(pos, x, rtol=0, atol=2e-6)
    part, _ = dcd.read(2, 2)
    assert np.array_equal(part, pos[2:4])
    with pytest.raises(IndexError):
        dcd.read(4, 2)


def test_dcd_threaded_read_equals_single_threaded(tmp_path):
    rng = np.random.default_rng(10)
    x = rng.uniform(-5, 5, size=(17, 70000, 3)).astype(np.float32)  # large enough to split
    path = tmp_path / "big.dcd"
    DCDFile.write(path, x)
    one, _ = DCDFile(path, readThreads=1).read(0, 17)
    many = DCDFile(path, readThreads=5)
    assert many.readThreads == 5
    got, _ = many.read(0, 17)
    assert np.array_equal(one, got)
    got, _ = many.read(3, 11)
    assert np.array_equal(one[3:14], got)


def test_dcd_header_layout_matches_the_format(tmp_path):
    """Byte-level check of the writer against the CHARMM layout OpenMM's DCDReporter emits."""
    path = tmp_path / "b.dcd"
    DCDFile.write(path, np.zeros((2, 3, 3)), boxVectors=[2.0, 3.0, 4.0], timeStep=0.5)
    raw = path.read_bytes()
    assert struct.unpack("<i", raw[:4])[0] == 84 and raw[4:8] == b"CORD"
    icntrl = struct.unpack("<20i", raw[8:88])
    assert icntrl[0] == 2 and icntrl[10] == 1 and icntrl[19] == 24
    assert struct.unpack("<f", raw[8 + 36:8 + 40])[0] == 0.5
    assert struct.unpack("<i", raw[88:92])[0] == 84
    title_len = struct.unpack("<i", raw[92:96])[0]
    assert title_len == 84 and struct.unpack("<i", raw[96:100])[0] == 1
    offset = 96 + title_len + 4
    assert struct.unpack("<3i", raw[offset:offset + 12]) == (4, 3, 4)
    offset += 12
    frame_bytes = (4 + 48 + 4) + 3 * (4 + 12 + 4)
    assert len(raw) == offset + 2 * frame_bytes
    assert struct.unpack("<i", raw[offset:offset + 4])[0] == 48
    cell = struct.unpack("<6d", raw[offset + 4:offset + 52])
    assert cell == (20.0, 90.0, 30.0, 90.0, 90.0, 40.0)  # A gamma B beta alpha C (Angstrom, degrees)


@pytest.mark.parametrize("cosines", [False, True])
def test_dcd_triclinic_box_round_trip(tmp_path, cosines):
    rng = np.random.default_rng(8)
    x = _frames(rng, 3, 4)
    boxes = np.zeros((3, 3, 3))
    for b in range(3):
        boxes[b] = [[3.0 + b, 0, 0], [0.5, 2.5, 0], [-0.4, 0.3, 2.0 + 0.1 * b]]
    path = tmp_path / "c.dcd"
    DCDFile.write(path, x, boxVectors=boxes)
    if cosines:  # CHARMM >= 25 stores the cosines of the angles instead of degrees
        dcd = DCDFile(path)
        raw = bytearray(path.read_bytes())
        frame = 56 + 3 * (8 + 4 * dcd.numAtoms)
        data_offset = len(raw) - 3 * frame
        dcd.close()
        for b in range(3):
            at = data_offset + b * frame + 4
            cell = list(struct.unpack("<6d", raw[at:at + 48]))
            for k in (1, 3, 4):
                cell[k] = np.cos(np.deg2rad(cell[k]))
            raw[at:at + 48] = struct.pack("<6d", *cell)
        del dcd
        path.write_bytes(bytes(raw))
    _, got = DCDFile(path).read(0, 3)
    assert np.allclose(got, boxes, atol=1e-12)
    assert np.all(got[:, 0, 1:] == 0) and np.all(got[:, 1, 2] == 0)  # reduced form, exact zeros
    rect = tmp_path / "d.dcd"
    DCDFile.write(rect, x, boxVectors=np.diag([2.0, 3.0, 4.0]))
    _, got = DCDFile(rect).read(1, 1)
    assert np.array_equal(got[0], np.diag([2.0, 3.0, 4.0]))


def test_dcd_rejects_foreign_files(tmp_path):
    path = tmp_path / "x.dcd"
    path.write_bytes(b"\0" * 200)
    with pytest.raises(ValueError, match="not a little-endian CORD DCD"):
        DCDFile(path)
    big = tmp_path / "y.dcd"  # big-endian marker
    big.write_bytes(struct.pack(">i", 84) + b"CORD" + b"\0" * 200)
    with pytest.raises(ValueError):
        DCDFile(big)


def test_numpy_trajectory_and_npy_memmap(tmp_path):
    rng = np.random.default_rng(9)
    x = _frames(rng, 6, 5).astype(np.float32)
    np.save(tmp_path / "t.npy", x)
    traj = open_trajectory(str(tmp_path / "t.npy"))
    assert isinstance(traj, NumpyTrajectory) and len(traj) == 6 and not traj.hasBox
    out = np.empty((2, 5, 3), dtype=np.float32)
    got, box = traj.read(3, 2, out)
    assert got is out and box is None and np.array_equal(out, x[3:5])
    boxed = NumpyTrajectory(x, boxVectors=[2.0, 2.0, 2.0])
    _, box = boxed.read(0, 6)
    assert box.shape == (6, 3, 3) and np.array_equal(box[5], np.diag([2.0, 2.0, 2.0]))
    with pytest.raises(ValueError):
        NumpyTrajectory(np.zeros((3, 4)))


def test_evaluator_headers_follow_the_reference_writer():
    """`name (unit)` / `emass[name] (mass unit)` (reference reporting/cv_writer.py:92-100)."""
    system = cvpack.System()
    for _ in range(20):
        system.addParticle(1.0)
    phi = cvpack.Torsion(6, 8, 14, 16, name="phi")
    psi = cvpack.Torsion(8, 14, 16, 18, name="psi")
    phi.addToSystem(system)
    psi.addToSystem(system)
    evaluator = TrajectoryEvaluator(system, [phi, psi], value=True, emass=True)
    assert evaluator.getHeaders() == [
        "phi (rad)", "emass[phi] (nm**2 Da/(rad**2))", "psi (rad)", "emass[psi] (nm**2 Da/(rad**2))",
    ]
    with pytest.raises(ValueError, match="At least one of value or effective_mass must be True"):
        TrajectoryEvaluator(system, [phi], value=False, emass=False)
    stray = cvpack.Distance(0, 1)
    with pytest.raises(RuntimeError, match="This force is not present in the given context."):
        TrajectoryEvaluator(system, [stray])
    out = io.StringIO()
    table = {"phi (rad)": np.array([1.5, -2.0]), "emass[phi] (nm**2 Da/(rad**2))": np.array([0.05, np.inf])}
    evaluator.writeCSV(out, table, firstStep=100, interval=100)
    assert out.getvalue().splitlines() == [
        '#"Step","phi (rad)","emass[phi] (nm**2 Da/(rad**2))"', "100,1.5,0.05", "200,-2.0,inf",
    ]


def test_bias_functions_against_finite_differences():
    s = torch.linspace(-3.0, 3.0, 41, dtype=torch.float64)
    h = 1e-6
    harmonic = cvbias.HarmonicBias(12.5, 0.4, period=2 * np.pi)
    hills = cvbias.GaussianHillsBias(0.3, period=2 * np.pi)
    assert hills.getNumHills() == 0 and float(hills.energy(s).abs().max()) == 0.0
    hills.addHills(torch.tensor([-2.9, 0.1, 3.0]), 1.2)
    hills.addHills(torch.tensor([0.5]), torch.tensor([0.7]))
    assert hills.getNumHills() == 4
    for bias in (harmonic, hills):
        fd = (bias.energy(s + h) - bias.energy(s - h)) / (2 * h)
        assert torch.allclose(bias.derivative(s), fd, atol=1e-6)
    # periodic image: a hill at -2.9 is felt just below +pi
    assert float(hills.energy(torch.tensor([3.3], dtype=torch.float64))) > 1.0


# ---- GPU -------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("with_box", [False, True])
@pytest.mark.parametrize("fmt", ["dcd", "xtc"])
def test_streamed_file_matches_whole_batch_and_oracle(tmp_path, with_box, fmt):
    rng = np.random.default_rng(11)
    frames, n = 37, 60  # ragged last chunk: 37 = 3 * 10 + 7
    box = np.array([3.0, 3.2, 3.4])
    x = rng.uniform(0.0, 3.0, size=(frames, n, 3))
    system = helpers.make_system(n, rng, box=np.diag(box) if with_box else None)
    rg = cvpack.RadiusOfGyration(range(10, 40), name="rg")
    nonbonded = helpers.make_nonbonded(n, rng, periodic=with_box, cutoff=1.0)
    contacts = cvpack.NumberOfContacts(range(0, 30), range(30, 60), nonbonded, name="nc")
    torsion = cvpack.Torsion(1, 5, 9, 13, pbc=with_box, name="phi")
    for cv in (rg, contacts, torsion):
        cv.addToSystem(system)
    path = tmp_path / f"traj.{fmt}"
    boxes = None
    if with_box:  # a different (rectangular) box per frame
        boxes = np.stack([np.diag(box * (1.0 + 0.01 * b)) for b in range(frames)])
    file_class = DCDFile if fmt == "dcd" else XTCFile
    file_class.write(path, x, boxVectors=boxes)
    stored, stored_box = file_class(path).read(0, frames)

    evaluator = TrajectoryEvaluator(system, [rg, contacts, torsion], value=True, emass=True, chunkFrames=10)
    table = evaluator.evaluate(str(path))
    assert list(table) == evaluator.getHeaders()

    ctx = cvpack.BatchContext(system, stored, None if stored_box is None else stored_box)
    for cv in (rg, contacts, torsion):
        # (value and gradient in one pass, as the evaluator does when it reports effective masses)
        whole = cv.getValueAndGradient(ctx)[0]._value.double().cpu().numpy()
        emass = cv.getEffectiveMass(ctx)._value.double().cpu().numpy()
        name = cv.getName()
        assert np.array_equal(table[f"{name} ({cv.getUnit().get_symbol()})"], whole)  # bitwise
        assert np.array_equal(table[f"emass[{name}] ({cv.getMassUnit().get_symbol()})"], emass)
        for b in (0, 9, 10, 36):
            ref, _ = helpers.oracle_evaluate(
                cv, system, stored[b].astype(np.float64), None if stored_box is None else stored_box[b]
            )
            assert whole[b] == pytest.approx(ref, rel=1e-5, abs=1e-6)
    sub = evaluator.evaluate(str(path), start=5, stop=23)
    assert np.array_equal(sub["rg (nm)"], table["rg (nm)"][5:23])
    assert len(evaluator.evaluate(str(path), start=30, stop=30)["rg (nm)"]) == 0


@pytest.mark.gpu
def test_streamed_array_value_only_and_atom_mismatch():
    rng = np.random.default_rng(12)
    frames, n = 25, 33
    x = rng.normal(size=(frames, n, 3)).astype(np.float32)
    system = helpers.make_system(n, rng)
    rmsd = cvpack.RMSD(x[0].astype(np.float64), range(n), n, name="rmsd")
    rmsd.addToSystem(system)
    table = TrajectoryEvaluator(system, [rmsd], chunkFrames=4).evaluate(x)
    ctx = cvpack.BatchContext(system, x)
    assert np.array_equal(table["rmsd (nm)"], rmsd.getValue(ctx)._value.double().cpu().numpy())
    assert table["rmsd (nm)"][0] == pytest.approx(0.0, abs=1e-6)
    with pytest.raises(ValueError, match="atoms"):
        TrajectoryEvaluator(system, [rmsd]).evaluate(np.zeros((2, n + 1, 3), dtype=np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["single", "double"])
def test_bias_forces_in_place(precision):
    rng = np.random.default_rng(13)
    walkers, n = 19, 48
    x = rng.uniform(0.0, 2.0, size=(walkers, n, 3))
    system = helpers.make_system(n, rng)
    phi = cvpack.Torsion(0, 7, 11, 20, name="phi")
    rg = cvpack.RadiusOfGyration(range(5, 40), name="rg")
    phi.addToSystem(system)
    rg.addToSystem(system)
    ctx = cvpack.BatchContext(system, x, precision=precision)
    hills = cvbias.GaussianHillsBias(0.4, period=2 * np.pi)
    hills.addHills(torch.tensor([0.3, -1.0, 2.5]), 2.0)
    spring = cvbias.HarmonicBias(50.0, 0.7)
    dtype = torch.float32 if precision == "single" else torch.float64
    base = torch.as_tensor(rng.normal(size=(walkers, n, 3)), dtype=dtype, device=ctx.getDevice())
    forces = base.clone()
    out, values, energy = cvbias.applyBias(ctx, [phi, rg], [hills, spring], forces)
    assert out is forces and values.shape == (2, walkers)
    tol = 1e-5 if precision == "single" else 1e-10
    xs = ctx.getPositions().double().cpu().numpy()
    expected = base.double().cpu().numpy().copy()
    total = np.zeros(walkers)
    for b in range(walkers):
        for cv, bias in ((phi, hills), (rg, spring)):
            s, g = helpers.oracle_evaluate(cv, system, xs[b], None)
            s_t = torch.tensor([s], dtype=torch.float64)
            expected[b] -= float(bias.derivative(s_t)) * g
            total[b] += float(bias.energy(s_t))
    scale = np.abs(expected).max()
    assert np.abs(forces.double().cpu().numpy() - expected).max() <= 20 * tol * scale
    assert np.allclose(energy.double().cpu().numpy(), total, rtol=50 * tol, atol=50 * tol)
    # without a force buffer: allocated, and exactly the bias part
    fresh, _, _ = cvbias.applyBias(ctx, [phi, rg], [hills, spring])
    assert np.abs((fresh - (forces - base)).double().cpu().numpy()).max() <= 20 * tol * scale
    with pytest.raises(ValueError):
        cvbias.accumulateBiasForces(forces, forces[:, :10].contiguous(), torch.zeros(walkers))
    with pytest.raises(ValueError):
        cvbias.accumulateBiasForces(forces.cpu(), forces.cpu(), torch.zeros(walkers))


def test_native_dcd_reader_matches_a_python_reading_of_the_file(tmp_path):
    """cvp_dcd_read (csrc/trajio.cu: mmap + host threads) against numpy on the raw bytes: every
    coordinate is the correctly rounded quotient x / 10, whatever the thread count and the window."""
    rng = np.random.default_rng(9)
    frames, atoms = 37, 1013
    x = rng.normal(size=(frames, atoms, 3)) * 3.0
    path = tmp_path / "big.dcd"
    DCDFile.write(path, x, boxVectors=np.diag([7.0, 8.0, 9.0]), firstStep=100, interval=10)
    raw = path.read_bytes()
    frame_bytes = 56 + 3 * (8 + 4 * atoms)
    offset = len(raw) - frames * frame_bytes
    expected = np.empty((frames, atoms, 3), dtype=np.float32)
    for b in range(frames):
        base = offset + b * frame_bytes + 56
        for k in range(3):
            record = np.frombuffer(raw, dtype="<f4", count=atoms, offset=base + k * (8 + 4 * atoms) + 4)
            expected[b, :, k] = record / np.float32(10.0)
    for threads in (1, 3, 64):
        dcd = DCDFile(path, readThreads=threads)
        assert (dcd.numFrames, dcd.numAtoms, dcd.hasBox, dcd.firstStep, dcd.interval) == (frames, atoms, True, 100, 10)
        got, boxes = dcd.read(0, frames)
        assert np.array_equal(got, expected)
        assert np.array_equal(boxes[5], np.diag([7.0, 8.0, 9.0]))
        window = np.zeros((11, atoms, 3), dtype=np.float32)
        dcd.read(20, 11, positions=window)
        assert np.array_equal(window, expected[20:31])
        with pytest.raises(IndexError):
            dcd.read(30, 8)
        with pytest.raises(ValueError):
            dcd.read(0, 2, positions=np.zeros((2, atoms, 3), dtype=np.float64))
        dcd.close()


def test_npz_archives(tmp_path):
    rng = np.random.default_rng(10)
    x = rng.normal(size=(5, 7, 3)).astype(np.float32)
    box = np.diag([2.0, 2.5, 3.0])
    np.savez(tmp_path / "a.npz", xyz=x, unitcell_vectors=np.broadcast_to(box, (5, 3, 3)))
    np.savez(tmp_path / "b.npz", positions=x)
    np.savez(tmp_path / "c.npz", other=x)
    traj = open_trajectory(tmp_path / "a.npz")
    got, boxes = traj.read(1, 3)
    assert np.array_equal(got, x[1:4]) and np.array_equal(boxes[0], box) and traj.hasBox
    traj = open_trajectory(tmp_path / "b.npz")
    assert traj.numFrames == 5 and not traj.hasBox
    with pytest.raises(ValueError, match="no 'positions'"):
        open_trajectory(tmp_path / "c.npz")


# ---- XTC (GROMACS compressed coordinates) -----------------------------------------------------------
def _water_box(rng, molecules, edge):
    """Three-site molecules (sites within 0.1 nm of the first) scattered in a box: the case the
    format's run coding is made for."""
    centres = rng.uniform(0.0, edge, size=(molecules, 1, 3))
    sites = centres + np.concatenate([np.zeros((molecules, 1, 3)), rng.normal(0, 0.05, (molecules, 2, 3))], axis=1)
    return sites.reshape(-1, 3)


def test_xtc_uncompressed_frame_is_the_documented_byte_layout(tmp_path):
    # <= 9 atoms: magic 1995, natoms, step, time, 9 box floats, natoms again, raw floats; big-endian
    x = np.arange(15, dtype=np.float32).reshape(1, 5, 3) * np.float32(0.125) - np.float32(0.5)
    box = np.array([[3.0, 0, 0], [0.5, 4.0, 0], [0.25, 0.75, 5.0]])
    path = tmp_path / "small.xtc"
    XTCFile.write(path, x, box, steps=[42], times=[8.5])
    expected = struct.pack(">iiif", 1995, 5, 42, 8.5) + struct.pack(">9f", *box.ravel())
    expected += struct.pack(">i", 5) + struct.pack(">15f", *x.ravel())
    assert path.read_bytes() == expected
    xtc = XTCFile(path)
    assert (len(xtc), xtc.numAtoms, xtc.precision) == (1, 5, 0.0)
    pos, boxes, steps, times = xtc.readFrames(0, 1)
    assert np.array_equal(pos, x) and np.array_equal(boxes[0], box)
    assert steps.tolist() == [42] and times.tolist() == [8.5]


def test_xtc_has_box_only_when_the_writer_had_one(tmp_path):
    """Every XTC frame has a box record; a file written without boxes holds zeros there and must be
    streamed as a trajectory without periodic box (it used to reach the context as a zero box)."""
    rng = np.random.default_rng(23)
    x = _water_box(rng, 10, 2.0)[None].repeat(3, axis=0)
    XTCFile.write(tmp_path / "nobox.xtc", x)
    XTCFile.write(tmp_path / "box.xtc", x, [2.0, 2.0, 2.0])
    assert XTCFile(tmp_path / "nobox.xtc").hasBox is False
    assert XTCFile(tmp_path / "nobox.xtc").read(0, 3)[1] is None
    assert XTCFile(tmp_path / "box.xtc").hasBox is True


def test_xtc_compressed_header_fields(tmp_path):
    rng = np.random.default_rng(17)
    x = _water_box(rng, 40, 3.0)[None]
    path = tmp_path / "w.xtc"
    XTCFile.write(path, x, [3.0, 3.0, 3.0], precision=1000.0)
    blob = path.read_bytes()
    natoms, precision = struct.unpack_from(">if", blob, 52)
    minint = np.array(struct.unpack_from(">3i", blob, 60))
    maxint = np.array(struct.unpack_from(">3i", blob, 72))
    smallidx, nbytes = struct.unpack_from(">ii", blob, 84)
    assert natoms == 120 and precision == 1000.0
    ints = np.trunc(x[0].astype(np.float32) * np.float32(1000.0) + np.copysign(np.float32(0.5), x[0])).astype(int)
    assert np.array_equal(minint, ints.min(axis=0)) and np.array_equal(maxint, ints.max(axis=0))
    assert 9 <= smallidx < 73 and len(blob) == 92 + (nbytes + 3) // 4 * 4
    # far fewer bits than 3 x 12 per atom: the run coding engaged
    assert nbytes < 120 * 36 // 8


@pytest.mark.parametrize("case", ["water", "gas", "huge", "tiny-spread", "ten"])
def test_xtc_round_trip_and_independent_decoder(tmp_path, case):
    rng = np.random.default_rng(23)
    precision = 1000.0
    if case == "water":
        x = np.stack([_water_box(rng, 300, 4.0) for _ in range(4)])
    elif case == "gas":  # no two consecutive atoms close: every atom takes the full-size code
        x = rng.uniform(-6.0, 6.0, size=(3, 257, 3))
    elif case == "huge":  # extent above 2^24 quanta: the per-component bit fields
        x, precision = rng.uniform(-10.0, 10.0, size=(2, 100, 3)), 1.0e6
    elif case == "tiny-spread":  # everything within a few quanta: smallest radix, long runs
        x = 1.0 + rng.integers(0, 4, size=(2, 64, 3)) * 1e-3
    else:  # the smallest compressed frame
        x = rng.uniform(0.0, 2.0, size=(3, 10, 3))
    frames = x.shape[0]
    path = tmp_path / "t.xtc"
    boxes = np.array([np.diag([4.0, 4.5, 5.0]) + 0.01 * b for b in range(frames)])
    XTCFile.write(path, x, boxes, steps=np.arange(frames) * 500, times=np.arange(frames) * 1.0, precision=precision)
    xtc = open_trajectory(path)
    assert isinstance(xtc, XTCFile) and (len(xtc), xtc.numAtoms) == (frames, x.shape[1])
    assert xtc.precision == np.float32(precision)
    pos, box, steps, times = xtc.readFrames(0, frames)
    # rounding to the grid: half a quantum plus float32 rounding of the product
    assert np.abs(pos - x).max() <= 0.5 / precision + 2e-6 * np.abs(x).max()
    assert np.allclose(box, boxes.astype(np.float32)) and steps.tolist() == (np.arange(frames) * 500).tolist()
    # the independent reading of the same bytes
    reference = xtc_python.read_frames(path)
    assert len(reference) == frames
    for b, (step, time, ref_box, xyz) in enumerate(reference):
        assert step == steps[b] and time == times[b]
        assert np.array_equal(ref_box, box[b].astype(np.float32))
        assert np.array_equal(xyz, pos[b]), f"frame {b} differs from the reference decoder"
    # decoded values sit exactly on the grid: a second round trip is lossless
    again = tmp_path / "again.xtc"
    XTCFile.write(again, pos, box, steps=steps, times=times, precision=precision)
    second = XTCFile(again).readFrames(0, frames)[0]
    if case == "huge":  # 1e7 quanta leave float32 one bit: the product can round to the neighbour
        assert np.abs(second - pos).max() <= 1.5 / precision
    else:
        assert np.array_equal(second, pos)
    # random access and threads
    one = XTCFile(path, readThreads=1)
    assert np.array_equal(one.read(frames - 1, 1)[0][0], pos[-1])
    with pytest.raises(IndexError):
        one.read(frames, 1)


def test_xtc_append_truncation_and_foreign_files(tmp_path):
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 3, size=(4, 50, 3))
    path = tmp_path / "a.xtc"
    XTCFile.write(path, x[:2])
    XTCFile.write(path, x[2:], append=True)
    whole = XTCFile(path)
    assert len(whole) == 4
    pos, box = whole.read(0, 4)
    assert box is None and np.abs(pos - x).max() <= 5.1e-4
    # a writer killed mid-frame leaves a partial last frame: ignored
    blob = path.read_bytes()
    (tmp_path / "cut.xtc").write_bytes(blob[:-7])
    assert len(XTCFile(tmp_path / "cut.xtc")) == 3
    (tmp_path / "bad.xtc").write_bytes(b"\0" * 200)
    with pytest.raises(Exception, match="not an XTC file"):
        XTCFile(tmp_path / "bad.xtc")
    other = tmp_path / "other.xtc"
    XTCFile.write(other, x[:1, :20])
    (tmp_path / "mixed.xtc").write_bytes(blob + other.read_bytes())
    with pytest.raises(Exception, match="number of atoms changes"):
        XTCFile(tmp_path / "mixed.xtc")
    with pytest.raises(Exception, match="too large"):
        XTCFile.write(tmp_path / "big.xtc", np.full((1, 12, 3), 3.0e6), precision=1000.0)
