"""
ctypes binding of ``libcvpack_b200.so`` (the C ABI of ``include/cvpack_b200.h``).

The CUDA library is the only evaluation path: if it cannot be loaded, or there is no sm_100
device, evaluation raises -- there is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import os
import typing as t

import numpy as np

_LIB_NAME = "libcvpack_b200.so"
# (CVPACK_B200_LIB: another build of the same library, e.g. a tuning variant -- never a fallback)
_LIB_PATH = os.environ.get("CVPACK_B200_LIB") or os.path.join(
    os.path.dirname(os.path.abspath(__file__)), _LIB_NAME
)

CVP_F32, CVP_F64 = 0, 1
CVP_BOX_NONE, CVP_BOX_SHARED, CVP_BOX_PER_FRAME = 0, 1, 2
CVP_FLAG_TRICLINIC, CVP_FLAG_ACCUMULATE = 1, 2
CVP_STEP_RATIONAL, CVP_STEP_HEAVISIDE, CVP_STEP_RATIONAL_PAIR, CVP_STEP_PLUMED = 0, 1, 2, 3
CVP_PAIR_CONTACTS, CVP_PAIR_ATTRACTION, CVP_PAIR_SOFTMIN = 0, 1, 2
CVP_COMBINE_IDENTITY, CVP_COMBINE_STEP_SUM = 0, 1
CVP_COMBINE_PATH_PROGRESS, CVP_COMBINE_PATH_DEVIATION = 2, 3
CVP_TERM_IDENTITY, CVP_TERM_BOXCAR, CVP_TERM_HALF_COSINE = 0, 1, 2

CVP_EXPR_MAX_OPS, CVP_EXPR_MAX_INPUTS, CVP_EXPR_MAX_TEMPS = 192, 16, 16
CVP_EXPR_CONST, CVP_EXPR_INPUT, CVP_EXPR_LOAD, CVP_EXPR_STORE = 0, 1, 2, 3
(CVP_EXPR_ADD, CVP_EXPR_SUB, CVP_EXPR_MUL, CVP_EXPR_DIV, CVP_EXPR_POW, CVP_EXPR_MIN, CVP_EXPR_MAX,
 CVP_EXPR_ATAN2) = range(10, 18)
CVP_EXPR_SELECT = 20
(CVP_EXPR_NEG, CVP_EXPR_SQRT, CVP_EXPR_EXP, CVP_EXPR_LOG, CVP_EXPR_SIN, CVP_EXPR_COS, CVP_EXPR_TAN,
 CVP_EXPR_SEC, CVP_EXPR_CSC, CVP_EXPR_COT, CVP_EXPR_ASIN, CVP_EXPR_ACOS, CVP_EXPR_ATAN, CVP_EXPR_SINH,
 CVP_EXPR_COSH, CVP_EXPR_TANH, CVP_EXPR_ERF, CVP_EXPR_ERFC, CVP_EXPR_STEP, CVP_EXPR_DELTA,
 CVP_EXPR_SQUARE, CVP_EXPR_CUBE, CVP_EXPR_RECIP, CVP_EXPR_ABS, CVP_EXPR_FLOOR, CVP_EXPR_CEIL) = range(30, 56)

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class PairSumDesc(C.Structure):
    """``cvp_pair_sum_desc``"""

    _fields_ = [
        ("kind", C.c_int32),
        ("num_atoms", C.c_int32),
        ("n1", C.c_int32),
        ("group1", _i32p),
        ("n2", C.c_int32),
        ("group2", _i32p),
        ("num_exclusions", C.c_int32),
        ("exclusions", _i32p),
        ("periodic", C.c_int32),
        ("cutoff", C.c_double),
        ("switch_distance", C.c_double),
        ("scale", C.c_double),
        ("step_kind", C.c_int32),
        ("step_p0", C.c_double),
        ("step_p1", C.c_double),
        ("r0", C.c_double),
        ("charge", _f64p),
        ("sigma", _f64p),
        ("epsilon", _f64p),
        ("sign", _f64p),
        ("coulomb_constant", C.c_double),
        ("beta", C.c_double),
    ]


class CentroidPairSumDesc(C.Structure):
    """``cvp_centroid_pair_sum_desc``"""

    _fields_ = [
        ("num_atoms", C.c_int32),
        ("n1", C.c_int32),
        ("n2", C.c_int32),
        ("group_offsets", _i32p),
        ("group_atoms", _i32p),
        ("group_weights", _f64p),
        ("periodic", C.c_int32),
        ("step_kind", C.c_int32),
        ("step_p0", C.c_double),
        ("step_p1", C.c_double),
        ("r0", C.c_double),
        ("scale", C.c_double),
    ]


class RmsdDesc(C.Structure):
    """``cvp_rmsd_desc``"""

    _fields_ = [
        ("num_atoms", C.c_int32),
        ("num_groups", C.c_int32),
        ("group_offsets", _i32p),
        ("atoms", _i32p),
        ("body_ids", _i32p),
        ("reference", _f64p),
        ("combine", C.c_int32),
        ("step_kind", C.c_int32),
        ("step_p0", C.c_double),
        ("step_p1", C.c_double),
        ("r0", C.c_double),
        ("scale", C.c_double),
        ("sigma", C.c_double),
    ]


class GyrationDesc(C.Structure):
    """``cvp_gyration_desc``"""

    _fields_ = [
        ("num_atoms", C.c_int32),
        ("n", C.c_int32),
        ("group", _i32p),
        ("weights", _f64p),
        ("periodic", C.c_int32),
        ("squared", C.c_int32),
    ]


class BondedDesc(C.Structure):
    """``cvp_bonded_desc``"""

    _fields_ = [
        ("num_atoms", C.c_int32),
        ("num_terms", C.c_int32),
        ("atoms_per_term", C.c_int32),
        ("atoms", _i32p),
        ("function", C.c_int32),
        ("reference", _f64p),
        ("tolerance", C.c_double),
        ("half_exponent", C.c_int32),
        ("coefficient", C.c_double),
        ("periodic", C.c_int32),
    ]


class PathCVDesc(C.Structure):
    """``cvp_path_cv_desc``"""

    _fields_ = [
        ("num_variables", C.c_int32),
        ("num_milestones", C.c_int32),
        ("milestones", _f64p),
        ("periods", _f64p),
        ("scales", _f64p),
        ("sigma", C.c_double),
        ("metric", C.c_int32),
    ]


class ExprOp(C.Structure):
    """``cvp_expr_op``"""

    _fields_ = [("code", C.c_int32), ("arg", C.c_int32), ("value", C.c_double)]


class ExprProgram(C.Structure):
    """``cvp_expr_program``"""

    _fields_ = [
        ("num_ops", C.c_int32),
        ("num_variables", C.c_int32),
        ("num_parameters", C.c_int32),
        ("reserved", C.c_int32),
        ("parameters", C.c_double * CVP_EXPR_MAX_INPUTS),
        ("ops", ExprOp * CVP_EXPR_MAX_OPS),
    ]


# every symbol include/cvpack_b200.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
_eval_args = [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int64, C.c_int64, _vp, _vp, C.c_int]
SYMBOLS: t.Dict[str, t.Tuple[t.Any, t.List[t.Any]]] = {
    "cvp_pair_sum_create": (C.c_int, [C.POINTER(PairSumDesc), C.POINTER(_vp)]),
    "cvp_centroid_pair_sum_create": (C.c_int, [C.POINTER(CentroidPairSumDesc), C.POINTER(_vp)]),
    "cvp_rmsd_create": (C.c_int, [C.POINTER(RmsdDesc), C.POINTER(_vp)]),
    "cvp_gyration_create": (C.c_int, [C.POINTER(GyrationDesc), C.POINTER(_vp)]),
    "cvp_bonded_create": (C.c_int, [C.POINTER(BondedDesc), C.POINTER(_vp)]),
    "cvp_destroy": (C.c_int, [_vp]),
    "cvp_active_atoms": (C.c_int, [_vp, _i32p, _i32p]),
    "cvp_periodic_cutoff": (C.c_int, [_vp, _f64p]),
    "cvp_evaluate": (C.c_int, _eval_args + [_vp]),
    "cvp_evaluate_host": (C.c_int, _eval_args + [C.c_int]),
    "cvp_effective_mass": (C.c_int, [C.c_int, _vp, _vp, C.c_int64, C.c_int64, _vp, _vp]),
    "cvp_evaluate_with_mass": (
        C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int64, C.c_int64, _vp, _vp, _vp, C.c_int, _vp]
    ),
    "cvp_evaluate_with_mass_host": (
        C.c_int,
        [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp, C.c_int, C.c_int],
    ),
    "cvp_path_cv_combine": (C.c_int, [C.POINTER(PathCVDesc), C.c_int, _vp, C.c_int64, _vp, _vp, _vp]),
    "cvp_expr_eval": (C.c_int, [C.POINTER(ExprProgram), C.c_int, _vp, C.c_int64, _vp, _vp, _vp]),
    "cvp_axpy_frames": (C.c_int, [C.c_int, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int, _vp]),
    "cvp_peer_buffer_create": (C.c_int, [C.c_int64, C.POINTER(_vp)]),
    "cvp_peer_buffer_handle": (C.c_int, [_vp, _vp]),
    "cvp_peer_buffer_open": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "cvp_peer_buffer_local": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_int64)]),
    "cvp_peer_push": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, C.c_int64, C.c_int64, _vp, C.c_int32, _vp]),
    "cvp_peer_push_rows": (
        C.c_int, [_vp, C.c_int, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp]
    ),
    "cvp_peer_signal": (C.c_int, [_vp, _vp]),
    "cvp_peer_wait": (C.c_int, [_vp, _vp]),
    "cvp_peer_status": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "cvp_peer_buffer_destroy": (C.c_int, [_vp]),
    "cvp_dcd_open": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "cvp_dcd_info": (
        C.c_int,
        [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int32),
         C.POINTER(C.c_int32)],
    ),
    "cvp_dcd_read": (C.c_int, [_vp, C.c_int64, C.c_int64, _vp, _vp, C.c_int]),
    "cvp_dcd_close": (C.c_int, [_vp]),
    "cvp_xtc_open": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "cvp_xtc_info": (C.c_int, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_float)]),
    "cvp_xtc_read": (C.c_int, [_vp, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp, C.c_int]),
    "cvp_xtc_close": (C.c_int, [_vp]),
    "cvp_xtc_write": (
        C.c_int, [C.c_char_p, C.c_int, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp, C.c_float]
    ),
    "cvp_last_error": (C.c_char_p, []),
    "cvp_version": (C.c_int, []),
    "cvp_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cvp_launch_count": (C.c_int64, []),
    "cvp_host_chunks": (C.c_int, [C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.c_int32, C.POINTER(C.c_int32)]),
    "cvp_set_workspace_limit": (C.c_int, [_vp, C.c_int64]),
    "cvp_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "cvp_get_stat": (C.c_int, [_vp, C.c_char_p, _f64p]),
    "cvp_measure_fp32_peak": (C.c_int, [C.c_int, _f64p]),
    "cvp_measure_fp64_peak": (C.c_int, [C.c_int, _f64p]),
}

_lib: t.Optional[C.CDLL] = None


class NativeLibraryError(RuntimeError):
    """The CUDA library is missing or reported an error."""


def library_path() -> str:
    """Absolute path at which the shared library is expected."""
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (once) and declare the signatures of all entry points."""
    global _lib  # pylint: disable=global-statement
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise NativeLibraryError(
                f"{_LIB_NAME} has not been built (expected at {_LIB_PATH}); run `make` or "
                "`python -c 'import __graft_entry__ as g; g.build()'` in the repository root. "
                "There is no CPU fallback."
            )
        lib = C.CDLL(_LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            function = getattr(lib, name)
            function.restype = restype
            function.argtypes = argtypes
        _lib = lib
    return _lib


def check(status: int) -> None:
    """Turn a negative status into an exception carrying the library's message."""
    if status != 0:
        message = load().cvp_last_error().decode()
        if status == -1:
            raise ValueError(message)
        raise NativeLibraryError(f"libcvpack_b200 error {status}: {message}")


def as_i32(values: t.Any) -> np.ndarray:
    """Contiguous int32 array; raises if a value does not fit (index handling is exact)."""
    array = np.asarray(values)
    if array.size and not np.issubdtype(array.dtype, np.integer):
        if not np.all(np.equal(np.mod(array, 1), 0)):
            raise TypeError("atom indices must be integers")
    result = np.ascontiguousarray(array, dtype=np.int64)
    if result.size and (result.min() < -(2**31) or result.max() >= 2**31):
        raise OverflowError("atom index does not fit in 32 bits")
    return np.ascontiguousarray(result, dtype=np.int32)


def as_f64(values: t.Any) -> np.ndarray:
    """Contiguous float64 array."""
    return np.ascontiguousarray(np.asarray(values, dtype=np.float64))


def ptr_i32(array: t.Optional[np.ndarray]) -> t.Any:
    """ctypes pointer to an int32 array (or NULL)."""
    return None if array is None else array.ctypes.data_as(_i32p)


def ptr_f64(array: t.Optional[np.ndarray]) -> t.Any:
    """ctypes pointer to a float64 array (or NULL)."""
    return None if array is None else array.ctypes.data_as(_f64p)


class NativeCV:
    """Owner of one ``cvp_cv*`` kernel spec."""

    def __init__(self, handle: int, keepalive: t.Sequence[t.Any] = ()) -> None:
        self._handle = handle
        self._keepalive = list(keepalive)

    @classmethod
    def create(cls, function: str, desc: C.Structure, keepalive: t.Sequence[t.Any] = ()) -> "NativeCV":
        """Call one of the ``cvp_*_create`` functions."""
        handle = _vp()
        check(getattr(load(), function)(C.byref(desc), C.byref(handle)))
        return cls(handle.value, keepalive)

    @property
    def handle(self) -> int:
        """The raw ``cvp_cv*``."""
        return self._handle

    def active_atoms(self) -> np.ndarray:
        """Sorted indices of the atoms that can carry a non-zero gradient."""
        count = C.c_int32()
        check(load().cvp_active_atoms(self._handle, C.byref(count), None))
        atoms = np.empty(count.value, dtype=np.int32)
        check(load().cvp_active_atoms(self._handle, C.byref(count), ptr_i32(atoms)))
        return atoms

    def periodic_cutoff(self) -> float:
        """Cutoff of a minimum-image sum (0 if the spec has none)."""
        out = C.c_double()
        check(load().cvp_periodic_cutoff(self._handle, C.byref(out)))
        return out.value

    def set_workspace_limit(self, nbytes: int) -> None:
        """Cap the device workspace this spec may use per evaluate call."""
        check(load().cvp_set_workspace_limit(self._handle, int(nbytes)))

    def set_option(self, name: str, value: int) -> int:
        """Set a per-family tuning switch; returns the previous value (or -1 if unknown)."""
        return load().cvp_set_option(self._handle, name.encode(), int(value))

    def stat(self, name: str) -> float:
        """A named statistic of the spec (``cvp_get_stat``), e.g. ``"last_path"`` of a pair sum."""
        out = (C.c_double * 2)()
        check(load().cvp_get_stat(self._handle, name.encode(), out))
        return float(out[0])

    def kernel_time(self, kernel: str) -> t.Tuple[float, int]:
        """(device milliseconds, launches) of a named kernel since the last query."""
        out = (C.c_double * 2)()
        check(load().cvp_get_stat(self._handle, f"{kernel}_ms".encode(), out))
        return float(out[0]), int(out[1])

    def evaluate(  # pylint: disable=too-many-arguments
        self,
        dtype: int,
        positions: int,
        box: t.Optional[int],
        box_mode: int,
        num_frames: int,
        num_atoms: int,
        value: int,
        grad: t.Optional[int],
        flags: int,
        stream: int,
    ) -> None:
        """``cvp_evaluate`` on device pointers."""
        check(
            load().cvp_evaluate(
                self._handle, dtype, positions, box, box_mode, num_frames, num_atoms, value, grad,
                flags, stream,
            )
        )

    def evaluate_host(  # pylint: disable=too-many-arguments
        self,
        dtype: int,
        positions: int,
        box: t.Optional[int],
        box_mode: int,
        num_frames: int,
        num_atoms: int,
        value: int,
        grad: t.Optional[int],
        flags: int,
        device: int,
    ) -> None:
        """``cvp_evaluate_host`` on host pointers."""
        check(
            load().cvp_evaluate_host(
                self._handle, dtype, positions, box, box_mode, num_frames, num_atoms, value, grad,
                flags, device,
            )
        )

    def evaluate_with_mass(  # pylint: disable=too-many-arguments
        self, dtype: int, positions: int, box: t.Optional[int], box_mode: int, num_frames: int,
        num_atoms: int, masses: int, value: int, emass: int, flags: int, stream: int,
    ) -> None:
        """``cvp_evaluate_with_mass`` on device pointers: the gradient never leaves the library."""
        check(
            load().cvp_evaluate_with_mass(
                self._handle, dtype, positions, box, box_mode, num_frames, num_atoms, masses, value,
                emass, flags, stream,
            )
        )

    def evaluate_with_mass_host(  # pylint: disable=too-many-arguments
        self, dtype: int, positions: int, box: t.Optional[int], box_mode: int, num_frames: int,
        num_atoms: int, masses: int, value: int, emass: int, grad: t.Optional[int], flags: int,
        device: int,
    ) -> None:
        """``cvp_evaluate_with_mass_host`` on host pointers: only ``[B]`` results travel back."""
        check(
            load().cvp_evaluate_with_mass_host(
                self._handle, dtype, positions, box, box_mode, num_frames, num_atoms, masses, value,
                emass, grad, flags, device,
            )
        )

    def __del__(self) -> None:
        handle, self._handle = getattr(self, "_handle", None), None
        if handle and _lib is not None:
            _lib.cvp_destroy(handle)


def effective_mass(dtype: int, grad: int, masses: int, num_frames: int, num_atoms: int, out: int, stream: int) -> None:
    """``cvp_effective_mass`` on device pointers."""
    check(load().cvp_effective_mass(dtype, grad, masses, num_frames, num_atoms, out, stream))


def path_cv_combine(  # pylint: disable=too-many-arguments
    milestones: t.Any, periods: t.Any, scales: t.Any, sigma: float, metric: int, dtype: int,
    values: int, num_frames: int, out_value: int, out_coef: t.Optional[int], stream: int,
) -> None:
    """``cvp_path_cv_combine`` on device pointers (``values`` / ``out_coef``: ``[nv][B]``)."""
    matrix = as_f64(milestones)
    per, sca = as_f64(periods), as_f64(scales)
    desc = PathCVDesc(
        num_variables=matrix.shape[1], num_milestones=matrix.shape[0], milestones=ptr_f64(matrix),
        periods=ptr_f64(per), scales=ptr_f64(sca), sigma=float(sigma), metric=int(metric),
    )
    check(load().cvp_path_cv_combine(C.byref(desc), dtype, values, num_frames, out_value, out_coef, stream))


def expr_eval(  # pylint: disable=too-many-arguments
    program: ExprProgram, dtype: int, values: t.Optional[int], num_frames: int, out_value: int,
    out_partial: t.Optional[int], stream: int,
) -> None:
    """``cvp_expr_eval`` on device pointers (``values``: ``[nv][B]``, ``out_partial``: ``[nv+np][B]``)."""
    check(load().cvp_expr_eval(C.byref(program), dtype, values, num_frames, out_value, out_partial, stream))


def axpy_frames(  # pylint: disable=too-many-arguments
    dtype: int, coef: int, x: int, y: int, num_frames: int, num_atoms: int, accumulate: bool, stream: int
) -> None:
    """``cvp_axpy_frames`` on device pointers."""
    check(load().cvp_axpy_frames(dtype, coef, x, y, num_frames, num_atoms, int(accumulate), stream))


def device_info(device: int = 0) -> t.Tuple[int, int, int]:
    """(SM count, compute capability major, minor) of a device."""
    sm, major, minor = C.c_int(), C.c_int(), C.c_int()
    check(load().cvp_device_info(device, C.byref(sm), C.byref(major), C.byref(minor)))
    return sm.value, major.value, minor.value


def launch_count() -> int:
    """Kernels launched by the library since the process started."""
    return int(load().cvp_launch_count())


def measure_fp64_peak(device: int = 0) -> float:
    """Measured FP64 FMA throughput of a device in TFLOP/s."""
    out = C.c_double()
    check(load().cvp_measure_fp64_peak(device, C.byref(out)))
    return out.value


def measure_fp32_peak(device: int = 0) -> float:
    """Measured FP32 FMA throughput of a device in TFLOP/s."""
    out = C.c_double()
    check(load().cvp_measure_fp32_peak(device, C.byref(out)))
    return out.value
