"""ctypes wrapper of oracle/liboracle.so (the C restatement) -- TEST INFRASTRUCTURE ONLY."""

from __future__ import annotations

import ctypes as C
import os
import typing as t

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liboracle.so")
_lib: t.Optional[C.CDLL] = None
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


def load() -> C.CDLL:
    """Load liboracle.so (built by `make oracle`)."""
    global _lib  # pylint: disable=global-statement
    if _lib is None:
        _lib = C.CDLL(_PATH)
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def num_threads() -> int:
    """OpenMP threads the oracle uses."""
    return int(load().oracle_num_threads())


def use_all_threads() -> int:
    """Use every host thread this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        count = len(os.sched_getaffinity(0))
    except AttributeError:
        count = os.cpu_count() or 1
    load().oracle_set_num_threads(C.c_int(count))
    return num_threads()


def contacts_batch(x, group1, group2, exclusions=(), box=None, r0=0.3, rc=0.6, rs=0.45,
                   step_kind=0, step_n=6, reference=1.0, need_grad=True, cells=False):
    """NumberOfContacts of every frame of ``x`` [B, N, 3]; returns (value [B], grad [B, N, 3]).

    ``cells=True`` uses the cell-list variant (the neighbour-list analogue of OpenMM's CPU platform)
    and raises ``ValueError`` if the geometry is not supported (triclinic, box < 3 cutoffs)."""
    x = _f64(x)
    num_frames, num_atoms, _ = x.shape
    g1, g2 = _i32(sorted(set(map(int, group1)))), _i32(sorted(set(map(int, group2))))
    offsets = partners = None
    exclusions = list(exclusions)
    if exclusions:
        lists = [[] for _ in range(num_atoms)]
        for i, j in exclusions:
            lists[int(i)].append(int(j))
            lists[int(j)].append(int(i))
        offsets = _i32(np.concatenate([[0], np.cumsum([len(v) for v in lists])]))
        partners = _i32([p for v in lists for p in v] or [0])
    per_frame = 0
    if box is not None:
        box = _f64(box)
        per_frame = int(box.ndim == 3)
        box = box.reshape(-1)
    value = np.empty(num_frames)
    grad = np.empty_like(x) if need_grad else None
    function = load().oracle_contacts_cells_batch if cells else load().oracle_contacts_batch
    function.restype = C.c_int if cells else None
    status = function(
        _p(x, _f64p), C.c_int64(num_frames), C.c_int(num_atoms), _p(g1, _i32p), C.c_int(len(g1)),
        _p(g2, _i32p), C.c_int(len(g2)), _p(offsets, _i32p), _p(partners, _i32p), _p(box, _f64p),
        C.c_int(per_frame), C.c_double(r0), C.c_double(rc), C.c_double(-1.0 if rs is None else rs),
        C.c_int(step_kind), C.c_int(step_n), C.c_double(reference), _p(value, _f64p),
        _p(grad, _f64p),
    )
    if cells and status != 0:
        raise ValueError("cell-list oracle: unsupported geometry (triclinic box or box < 3 cutoffs)")
    return value, grad


def rmsd_batch(x, group, reference, need_grad=True):
    """RMSD of every frame; ``reference`` rows pair with ``group``."""
    x = _f64(x)
    num_frames, num_atoms, _ = x.shape
    group = _i32(group)
    reference = _f64(reference)
    value = np.empty(num_frames)
    grad = np.empty_like(x) if need_grad else None
    load().oracle_rmsd_batch(
        _p(x, _f64p), C.c_int64(num_frames), C.c_int(num_atoms), _p(group, _i32p),
        C.c_int(len(group)), _p(reference, _f64p), _p(value, _f64p), _p(grad, _f64p),
    )
    return value, grad


def gyration_batch(x, group, weights=None, squared=False, need_grad=True):
    """Radius of gyration of every frame (non-periodic)."""
    x = _f64(x)
    num_frames, num_atoms, _ = x.shape
    group = _i32(group)
    weights = _f64(np.full(len(group), 1.0 / len(group)) if weights is None else weights)
    value = np.empty(num_frames)
    grad = np.empty_like(x) if need_grad else None
    load().oracle_gyration_batch(
        _p(x, _f64p), C.c_int64(num_frames), C.c_int(num_atoms), _p(group, _i32p),
        C.c_int(len(group)), _p(weights, _f64p), C.c_int(int(squared)), _p(value, _f64p),
        _p(grad, _f64p),
    )
    return value, grad
