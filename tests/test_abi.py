"""
The C ABI: the shared library loads, exports every symbol include/cvpack_b200.h declares, and its
host-side spec builders validate their input (no kernel is launched: no GPU needed).
"""

import ctypes
import os
import re

import numpy as np
import pytest

from cvpack_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "cvpack_b200.h"), encoding="utf-8") as file:
        header = file.read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    return sorted(set(re.findall(r"\b(cvp_[a-z0-9_]+)\s*\(", header)))


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    symbols = declared_symbols()
    assert len(symbols) >= 18
    for name in symbols:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
        assert name in _native.SYMBOLS, f"{name} has no ctypes signature"
    assert lib.cvp_version() >= 100


def test_descriptor_layouts_match_header():
    """Field order of the ctypes structures mirrors the C structs (names checked against the header)."""
    with open(os.path.join(ROOT, "include", "cvpack_b200.h"), encoding="utf-8") as file:
        header = re.sub(r"/\*.*?\*/", "", file.read(), flags=re.S)
    for struct, cls in (
        ("cvp_pair_sum_desc", _native.PairSumDesc),
        ("cvp_centroid_pair_sum_desc", _native.CentroidPairSumDesc),
        ("cvp_rmsd_desc", _native.RmsdDesc),
        ("cvp_gyration_desc", _native.GyrationDesc),
        ("cvp_bonded_desc", _native.BondedDesc),
    ):
        end = header.index("} " + struct + ";")
        body = header[header.rindex("typedef struct {", 0, end) + len("typedef struct {") : end]
        names = []
        for statement in body.split(";"):
            statement = statement.strip()
            if not statement:
                continue
            for part in statement.split(","):
                names.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
        assert names == [field[0] for field in cls._fields_], struct


def test_spec_builders_validate_on_the_host():
    group = _native.as_i32([0, 1, 2])
    weights = _native.as_f64([1 / 3] * 3)
    good = _native.GyrationDesc(num_atoms=5, n=3, group=_native.ptr_i32(group),
                                weights=_native.ptr_f64(weights), periodic=0, squared=0)
    cv = _native.NativeCV.create("cvp_gyration_create", good, (group, weights))
    assert cv.active_atoms().tolist() == [0, 1, 2]
    bad_group = _native.as_i32([0, 1, 7])
    bad = _native.GyrationDesc(num_atoms=5, n=3, group=_native.ptr_i32(bad_group),
                               weights=_native.ptr_f64(weights), periodic=0, squared=0)
    with pytest.raises(ValueError, match="out of range"):
        _native.NativeCV.create("cvp_gyration_create", bad)
    atoms = _native.as_i32([[0, 1, 2, 3], [1, 2, 3, 4]])
    desc = _native.BondedDesc(num_atoms=5, num_terms=2, atoms_per_term=4, atoms=_native.ptr_i32(atoms),
                              function=_native.CVP_TERM_IDENTITY, coefficient=1.0)
    assert _native.NativeCV.create("cvp_bonded_create", desc, (atoms,)).active_atoms().tolist() == [0, 1, 2, 3, 4]
    desc.atoms_per_term = 5
    with pytest.raises(ValueError):
        _native.NativeCV.create("cvp_bonded_create", desc, (atoms,))


def test_pair_spec_sets_and_active_atoms():
    g1, g2 = _native.as_i32([5, 1, 1, 3]), _native.as_i32([3, 8, 9])
    desc = _native.PairSumDesc(kind=_native.CVP_PAIR_CONTACTS, num_atoms=12, n1=4,
                               group1=_native.ptr_i32(g1), n2=3, group2=_native.ptr_i32(g2),
                               cutoff=0.6, switch_distance=0.45, scale=1.0,
                               step_kind=_native.CVP_STEP_RATIONAL, step_p0=6.0, r0=0.3)
    cv = _native.NativeCV.create("cvp_pair_sum_create", desc, (g1, g2))
    assert cv.active_atoms().tolist() == [1, 3, 5, 8, 9]  # sets: duplicates dropped, sorted
    desc.r0 = 0.0
    with pytest.raises(ValueError, match="threshold"):
        _native.NativeCV.create("cvp_pair_sum_create", desc, (g1, g2))


def test_rmsd_spec_rejects_overlapping_bodies():
    offsets = _native.as_i32([0, 4])
    atoms = _native.as_i32([0, 1, 2, 1])
    ref = _native.as_f64(np.zeros((4, 3)))
    desc = _native.RmsdDesc(num_atoms=6, num_groups=1, group_offsets=_native.ptr_i32(offsets),
                            atoms=_native.ptr_i32(atoms), reference=_native.ptr_f64(ref),
                            combine=_native.CVP_COMBINE_IDENTITY)
    with pytest.raises(ValueError, match="repeated atom"):
        _native.NativeCV.create("cvp_rmsd_create", desc, (offsets, atoms, ref))


def test_evaluate_without_device_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    group = _native.as_i32([0, 1, 2])
    weights = _native.as_f64([1 / 3] * 3)
    desc = _native.GyrationDesc(num_atoms=3, n=3, group=_native.ptr_i32(group),
                                weights=_native.ptr_f64(weights), periodic=0, squared=0)
    cv = _native.NativeCV.create("cvp_gyration_create", desc, (group, weights))
    pos = np.zeros((1, 3, 3), dtype=np.float32)
    out = np.zeros(1, dtype=np.float32)
    with pytest.raises(_native.NativeLibraryError):
        cv.evaluate(_native.CVP_F32, pos.ctypes.data, None, 0, 1, 3, out.ctypes.data, None, 0, 0)
    sm = ctypes.c_int()
    assert _native.load().cvp_device_info(0, ctypes.byref(sm), None, None) != 0


@pytest.mark.parametrize(
    "frames,frame_bytes",
    [(4096, 240000), (2048, 1200000), (1, 12), (7, 10**9), (1184, 240000), (100000, 1200), (593, 240000)],
)
def test_host_chunk_schedule(frames, frame_bytes):
    """cvp_host_chunks: the chunks of cvp_evaluate_host cover the batch in order, none above 256 MiB
    of coordinates (or one frame), at least four when there are four frames; with long chunks the
    first and the last are short -- whole 148-frame waves -- and the ones in between are equal."""
    lib = _native.load()
    count = ctypes.c_int32()
    assert lib.cvp_host_chunks(frames, frame_bytes, None, 0, ctypes.byref(count)) == 0
    sizes = (ctypes.c_int64 * count.value)()
    assert lib.cvp_host_chunks(frames, frame_bytes, sizes, count.value, ctypes.byref(count)) == 0
    sizes = list(sizes)
    assert sum(sizes) == frames and all(s >= 1 for s in sizes)
    assert all(s == 1 or s * frame_bytes <= 256 << 20 for s in sizes)
    assert len(sizes) >= min(frames, 4)
    if len(sizes) >= 3 and sizes[0] < max(sizes):  # tapered
        assert sizes[0] == sizes[-1] and sizes[0] % 148 == 0
        middle = sizes[1:-1]
        assert max(middle) - min(middle) <= 1 and sizes[0] <= min(middle)


def test_host_chunk_schedule_of_the_headline_batch():
    lib = _native.load()
    count = ctypes.c_int32()
    sizes = (ctypes.c_int64 * 16)()
    assert lib.cvp_host_chunks(4096, 240000, sizes, 16, ctypes.byref(count)) == 0
    assert list(sizes)[: count.value] == [296, 876, 876, 876, 876, 296]
