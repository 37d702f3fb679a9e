"""
Trajectory ingestion: the step *before* the hot path.

cvpack's callers evaluate a CV once per reported step of a running simulation
(``cvpack/reporting/cv_writer.py:102-109``: ``getValue(context)`` / ``getEffectiveMass(context)``
per row of the ``StateDataReporter`` table, ``state_data_reporter.py:127-162``).  The batched
analogue reads the frames of a stored trajectory, streams them through pinned host staging buffers
to the GPU in chunks and writes the same table, one row per frame:

* :class:`DCDFile` reads and writes CHARMM/NAMD DCD files (the format OpenMM's ``DCDReporter``
  produces: ``CORD`` header, optional 48-byte unit-cell record, X/Y/Z ``float32`` records in
  Angstrom per frame).  Frames have a fixed size, so chunks are random-access views of a memory map.
* :class:`XTCFile` reads and writes GROMACS XTC files (compressed integer coordinates in nm;
  variable-size frames, indexed once on opening and decoded by host threads).
* :class:`NumpyTrajectory` wraps an ``[B, N, 3]`` array (nm) in memory or memory-mapped from ``.npy``.
* :class:`TrajectoryEvaluator` runs the chunks through a two-deep pipeline: a reader thread fills
  pinned buffer ``k+1`` (de-interleaving X/Y/Z, Angstrom -> nm) while the copy stream uploads
  buffer ``k`` and the compute stream evaluates chunk ``k-1``; only the ``[B]`` values (and
  effective masses) return to the host.  ``writeCSV`` prints the reference's column headers
  (``name (unit)``, ``emass[name] (mass unit)``, ``cv_writer.py:92-100``).

Everything numerical happens in the CUDA library through :class:`~cvpack_b200.batch.BatchContext`;
this module is host-side plumbing.
"""

from __future__ import annotations

import ctypes as C
import math
import os
import queue
import struct
import threading
import typing as t

import numpy as np
import torch

from . import _native
from . import mm
from .batch import BatchContext

_ANGSTROM_PER_NM = 10.0


def _box_vectors_from_lengths_angles(
    lengths_nm: np.ndarray, angles_rad: np.ndarray
) -> np.ndarray:
    """``[..., 3]`` edge lengths and (alpha, beta, gamma) -> ``[..., 3, 3]`` reduced-form rows."""
    a, b, c = lengths_nm[..., 0], lengths_nm[..., 1], lengths_nm[..., 2]
    alpha, beta, gamma = angles_rad[..., 0], angles_rad[..., 1], angles_rad[..., 2]
    vectors = np.zeros(lengths_nm.shape[:-1] + (3, 3), dtype=np.float64)
    vectors[..., 0, 0] = a
    vectors[..., 1, 0] = b * np.cos(gamma)
    vectors[..., 1, 1] = b * np.sin(gamma)
    cx = c * np.cos(beta)
    cy = c * (np.cos(alpha) - np.cos(beta) * np.cos(gamma)) / np.sin(gamma)
    vectors[..., 2, 0] = cx
    vectors[..., 2, 1] = cy
    vectors[..., 2, 2] = np.sqrt(np.maximum(c * c - cx * cx - cy * cy, 0.0))
    # exact zeros for right angles (cos(pi/2) is 6e-17 in floating point)
    vectors[np.abs(vectors) < 1e-12 * np.max(np.abs(vectors), initial=1.0)] = 0.0
    return vectors


def _lengths_angles_from_box_vectors(vectors_nm: np.ndarray) -> t.Tuple[np.ndarray, np.ndarray]:
    a, b, c = vectors_nm[..., 0, :], vectors_nm[..., 1, :], vectors_nm[..., 2, :]
    la, lb, lc = (np.linalg.norm(v, axis=-1) for v in (a, b, c))
    alpha = np.arccos(np.clip(np.sum(b * c, axis=-1) / (lb * lc), -1.0, 1.0))
    beta = np.arccos(np.clip(np.sum(a * c, axis=-1) / (la * lc), -1.0, 1.0))
    gamma = np.arccos(np.clip(np.sum(a * b, axis=-1) / (la * lb), -1.0, 1.0))
    return np.stack([la, lb, lc], axis=-1), np.stack([alpha, beta, gamma], axis=-1)


class DCDFile:
    """
    Random-access reader of a DCD trajectory (little-endian, 32-bit record markers).

    Attributes
    ----------
    numFrames, numAtoms
        Sizes found in the file (the frame count is derived from the file size when the header
        says 0 or disagrees, as writers that were interrupted leave it stale).
    hasBox
        Whether every frame carries a unit-cell record.
    """

    def __init__(self, path: t.Union[str, os.PathLike], readThreads: t.Optional[int] = None) -> None:
        self.path = os.fspath(path)
        if readThreads is None:
            try:
                readThreads = len(os.sched_getaffinity(0))
            except AttributeError:
                readThreads = os.cpu_count() or 1
        self.readThreads = max(1, int(readThreads))
        # the file is parsed, mapped and de-interleaved by the native library (csrc/trajio.cu)
        lib = _native.load()
        handle = C.c_void_p()
        _native.check(lib.cvp_dcd_open(self.path.encode(), C.byref(handle)))
        self._handle: t.Optional[int] = handle.value
        frames, atoms = C.c_int64(), C.c_int64()
        has_box, first, interval = C.c_int(), C.c_int32(), C.c_int32()
        _native.check(
            lib.cvp_dcd_info(self._handle, C.byref(frames), C.byref(atoms), C.byref(has_box),
                             C.byref(first), C.byref(interval))
        )
        self.numFrames, self.numAtoms = int(frames.value), int(atoms.value)
        self.hasBox = bool(has_box.value)
        self.firstStep, self.interval = int(first.value), int(interval.value)

    def close(self) -> None:
        """Unmap the file."""
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            _native.load().cvp_dcd_close(handle)

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass

    def __len__(self) -> int:
        return self.numFrames

    def read(
        self,
        start: int,
        count: int,
        positions: t.Optional[np.ndarray] = None,
        boxVectors: t.Optional[np.ndarray] = None,
    ) -> t.Tuple[np.ndarray, t.Optional[np.ndarray]]:
        """
        Frames ``start .. start+count-1`` as ``[count, N, 3]`` ``float32`` **nm** (into
        ``positions`` if given: any writable C-contiguous array of that shape, e.g. a view of pinned
        memory) and their box vectors ``[count, 3, 3]`` ``float64`` nm (``None`` without unit
        cells).  The X/Y/Z records are de-interleaved by ``cvp_dcd_read`` on ``readThreads`` host
        threads, straight into the destination.
        """
        if start < 0 or count < 0 or start + count > self.numFrames:
            raise IndexError(f"frames {start}..{start + count - 1} outside 0..{self.numFrames - 1}")
        if positions is None:
            positions = np.empty((count, self.numAtoms, 3), dtype=np.float32)
        if positions.shape != (count, self.numAtoms, 3) or positions.dtype != np.float32:
            raise ValueError("positions must be a float32 array of shape [count, numAtoms, 3]")
        if not positions.flags.c_contiguous or not positions.flags.writeable:
            raise ValueError("positions must be C-contiguous and writable")
        cell = np.empty((count, 6), dtype=np.float64) if self.hasBox else None
        _native.check(
            _native.load().cvp_dcd_read(
                self._handle, int(start), int(count), positions.ctypes.data,
                None if cell is None else cell.ctypes.data, self.readThreads,
            )
        )
        boxes = None
        if cell is not None:
            lengths = cell[:, [0, 2, 5]] / _ANGSTROM_PER_NM  # A, gamma, B, beta, alpha, C
            raw = cell[:, [4, 3, 1]]  # alpha, beta, gamma
            # CHARMM >= 25 stores cosines, everybody else degrees
            cosines = np.all(np.abs(raw) <= 1.0, axis=1, keepdims=True)
            angles = np.where(cosines, np.arccos(np.clip(raw, -1.0, 1.0)), np.deg2rad(raw))
            boxes = _box_vectors_from_lengths_angles(lengths, angles)
            if boxVectors is not None:
                boxVectors[...] = boxes
                boxes = boxVectors
        return positions, boxes

    @staticmethod
    def write(
        path: t.Union[str, os.PathLike],
        positions: t.Any,
        boxVectors: t.Any = None,
        firstStep: int = 0,
        interval: int = 1,
        timeStep: float = 0.0,
    ) -> None:
        """
        Write ``positions`` (``[B, N, 3]`` nm) and optional box vectors (``[3, 3]`` or
        ``[B, 3, 3]`` nm) the way OpenMM's ``DCDReporter`` does (lengths in Angstrom, angles in
        degrees).
        """
        pos = np.asarray(positions, dtype=np.float64)
        if pos.ndim == 2:
            pos = pos[None]
        if pos.ndim != 3 or pos.shape[2] != 3:
            raise ValueError("positions must have shape [B, N, 3]")
        frames, atoms = pos.shape[0], pos.shape[1]
        cells = None
        if boxVectors is not None:
            box = np.asarray(boxVectors, dtype=np.float64)
            if box.shape == (3,):
                box = np.diag(box)
            if box.ndim == 2:
                box = np.broadcast_to(box, (frames, 3, 3))
            lengths, angles = _lengths_angles_from_box_vectors(box)
            cells = np.empty((frames, 6), dtype=np.float64)
            cells[:, [0, 2, 5]] = lengths * _ANGSTROM_PER_NM
            cells[:, [4, 3, 1]] = np.rad2deg(angles)
        icntrl = [0] * 20
        icntrl[0], icntrl[1], icntrl[2], icntrl[3] = frames, firstStep, interval, firstStep + interval * max(frames - 1, 0)
        icntrl[10] = 1 if cells is not None else 0
        icntrl[19] = 24
        header = struct.pack("<i4s9if10ii", 84, b"CORD", *icntrl[:9], float(timeStep), *icntrl[10:], 84)
        title = b"Created by cvpack_b200".ljust(80)
        with open(os.fspath(path), "wb") as file:
            file.write(header)
            file.write(struct.pack("<ii", 84, 1) + title + struct.pack("<i", 84))
            file.write(struct.pack("<3i", 4, atoms, 4))
            marker = struct.pack("<i", 4 * atoms)
            for b in range(frames):
                if cells is not None:
                    file.write(struct.pack("<i", 48) + cells[b].astype("<f8").tobytes() + struct.pack("<i", 48))
                frame = (pos[b] * _ANGSTROM_PER_NM).astype("<f4")
                for k in range(3):
                    file.write(marker + np.ascontiguousarray(frame[:, k]).tobytes() + marker)


class XTCFile:
    """
    Random-access reader (and writer) of a GROMACS XTC trajectory: compressed coordinates in nm,
    a 3x3 box, step and time per frame.  Frames have variable size, so opening the file walks the
    frame headers once (``cvp_xtc_open`` keeps the offset index); chunks are then decoded by
    ``readThreads`` host threads straight into the destination (``csrc/xtc.cu``).

    Attributes
    ----------
    numFrames, numAtoms
        Sizes found in the file (a truncated last frame is ignored).
    precision
        The writer's rounding (integers per nm, usually 1000); 0 for files of at most 9 atoms,
        which are stored uncompressed.
    hasBox
        Every XTC frame carries box vectors, all zeros when the writer had none: true when the box
        of the first frame is not all zeros.
    """

    def __init__(self, path: t.Union[str, os.PathLike], readThreads: t.Optional[int] = None) -> None:
        self.path = os.fspath(path)
        if readThreads is None:
            try:
                readThreads = len(os.sched_getaffinity(0))
            except AttributeError:
                readThreads = os.cpu_count() or 1
        self.readThreads = max(1, int(readThreads))
        lib = _native.load()
        handle = C.c_void_p()
        _native.check(lib.cvp_xtc_open(self.path.encode(), C.byref(handle)))
        self._handle: t.Optional[int] = handle.value
        frames, atoms, precision = C.c_int64(), C.c_int64(), C.c_float()
        _native.check(lib.cvp_xtc_info(self._handle, C.byref(frames), C.byref(atoms), C.byref(precision)))
        self.numFrames, self.numAtoms = int(frames.value), int(atoms.value)
        self.precision = float(precision.value)
        self.hasBox = bool(self.numFrames and self.readFrames(0, 1)[1].any())

    def close(self) -> None:
        """Unmap the file."""
        handle, self._handle = getattr(self, "_handle", None), None
        if handle:
            _native.load().cvp_xtc_close(handle)

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass

    def __len__(self) -> int:
        return self.numFrames

    def read(
        self,
        start: int,
        count: int,
        positions: t.Optional[np.ndarray] = None,
        boxVectors: t.Optional[np.ndarray] = None,
    ) -> t.Tuple[np.ndarray, t.Optional[np.ndarray]]:
        """Same contract as :meth:`DCDFile.read`; a file whose boxes are all zero returns ``None``."""
        positions, boxes, _, _ = self.readFrames(start, count, positions)
        if not boxes.any():
            return positions, None
        if boxVectors is not None:
            boxVectors[...] = boxes
            boxes = boxVectors
        return positions, boxes

    def readFrames(
        self, start: int, count: int, positions: t.Optional[np.ndarray] = None
    ) -> t.Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """Positions ``[count, N, 3]`` nm, boxes ``[count, 3, 3]`` nm, steps ``[count]``, times ``[count]`` (ps)."""
        if start < 0 or count < 0 or start + count > self.numFrames:
            raise IndexError(f"frames {start}..{start + count - 1} outside 0..{self.numFrames - 1}")
        if positions is None:
            positions = np.empty((count, self.numAtoms, 3), dtype=np.float32)
        if positions.shape != (count, self.numAtoms, 3) or positions.dtype != np.float32:
            raise ValueError("positions must be a float32 array of shape [count, numAtoms, 3]")
        if not positions.flags.c_contiguous or not positions.flags.writeable:
            raise ValueError("positions must be C-contiguous and writable")
        boxes = np.zeros((count, 3, 3), dtype=np.float64)
        steps = np.zeros(count, dtype=np.int32)
        times = np.zeros(count, dtype=np.float32)
        _native.check(
            _native.load().cvp_xtc_read(
                self._handle, int(start), int(count), positions.ctypes.data, boxes.ctypes.data,
                steps.ctypes.data, times.ctypes.data, self.readThreads,
            )
        )
        return positions, boxes, steps, times

    @staticmethod
    def write(  # pylint: disable=too-many-arguments
        path: t.Union[str, os.PathLike],
        positions: t.Any,
        boxVectors: t.Any = None,
        steps: t.Any = None,
        times: t.Any = None,
        precision: float = 1000.0,
        append: bool = False,
    ) -> None:
        """
        Write ``positions`` (``[B, N, 3]`` nm) with optional box vectors (``[3]``, ``[3, 3]`` or
        ``[B, 3, 3]`` nm), steps and times (ps) as XTC frames rounded to ``1/precision`` nm.
        """
        pos = np.ascontiguousarray(positions, dtype=np.float32)
        if pos.ndim == 2:
            pos = pos[None]
        if pos.ndim != 3 or pos.shape[2] != 3:
            raise ValueError("positions must have shape [B, N, 3]")
        frames, atoms = pos.shape[0], pos.shape[1]
        box = None
        if boxVectors is not None:
            box = np.asarray(boxVectors, dtype=np.float64)
            if box.shape == (3,):
                box = np.diag(box)
            if box.ndim == 2:
                box = np.broadcast_to(box, (frames, 3, 3))
            if box.shape != (frames, 3, 3):
                raise ValueError("boxVectors must have shape [3], [3, 3] or [B, 3, 3]")
            box = np.ascontiguousarray(box)
        step_array = None if steps is None else np.ascontiguousarray(steps, dtype=np.int32)
        time_array = None if times is None else np.ascontiguousarray(times, dtype=np.float32)
        for name, array in (("steps", step_array), ("times", time_array)):
            if array is not None and array.shape != (frames,):
                raise ValueError(f"{name} must have shape [B]")
        _native.check(
            _native.load().cvp_xtc_write(
                os.fspath(path).encode(), int(bool(append)), frames, atoms, pos.ctypes.data,
                None if box is None else box.ctypes.data,
                None if step_array is None else step_array.ctypes.data,
                None if time_array is None else time_array.ctypes.data,
                float(precision),
            )
        )


class NumpyTrajectory:
    """Frames held in (or memory-mapped from) an ``[B, N, 3]`` array in nm, optional boxes."""

    def __init__(self, positions: t.Any, boxVectors: t.Any = None) -> None:
        if isinstance(positions, (str, os.PathLike)):
            path = os.fspath(positions)
            if path.lower().endswith(".npz"):
                # an archive of arrays: coordinates under "positions" / "xyz" / "coordinates" (nm),
                # optional box under "box" / "boxVectors" / "unitcell_vectors"
                archive = np.load(path)
                key = next((k for k in ("positions", "xyz", "coordinates") if k in archive.files), None)
                if key is None:
                    raise ValueError(f"{path}: no 'positions', 'xyz' or 'coordinates' array in the archive")
                positions = archive[key]
                if boxVectors is None:
                    box_key = next(
                        (k for k in ("box", "boxVectors", "unitcell_vectors") if k in archive.files), None
                    )
                    boxVectors = None if box_key is None else archive[box_key]
            else:
                positions = np.load(path, mmap_mode="r")
        if isinstance(positions, torch.Tensor):
            positions = positions.detach().cpu().numpy()
        if positions.ndim != 3 or positions.shape[2] != 3:
            raise ValueError("positions must have shape [B, N, 3]")
        self._positions = positions
        self.numFrames, self.numAtoms = int(positions.shape[0]), int(positions.shape[1])
        self._box = None
        if boxVectors is not None:
            box = np.asarray(boxVectors, dtype=np.float64)
            if box.shape == (3,):
                box = np.diag(box)
            if box.ndim == 2:
                box = np.broadcast_to(box, (self.numFrames, 3, 3))
            if box.shape != (self.numFrames, 3, 3):
                raise ValueError("boxVectors must have shape [3], [3, 3] or [B, 3, 3]")
            self._box = box
        self.hasBox = self._box is not None

    def __len__(self) -> int:
        return self.numFrames

    def read(
        self,
        start: int,
        count: int,
        positions: t.Optional[np.ndarray] = None,
        boxVectors: t.Optional[np.ndarray] = None,
    ) -> t.Tuple[np.ndarray, t.Optional[np.ndarray]]:
        """Same contract as :meth:`DCDFile.read`."""
        if start < 0 or count < 0 or start + count > self.numFrames:
            raise IndexError(f"frames {start}..{start + count - 1} outside 0..{self.numFrames - 1}")
        if positions is None:
            positions = np.empty((count, self.numAtoms, 3), dtype=np.float32)
        np.copyto(positions, self._positions[start:start + count], casting="same_kind")
        boxes = None
        if self._box is not None:
            boxes = np.array(self._box[start:start + count])
            if boxVectors is not None:
                boxVectors[...] = boxes
                boxes = boxVectors
        return positions, boxes


def open_trajectory(source: t.Any, boxVectors: t.Any = None) -> t.Any:
    """A :class:`DCDFile` / :class:`XTCFile` for ``*.dcd`` / ``*.xtc`` paths, else a :class:`NumpyTrajectory` (arrays, ``.npy``, ``.npz``)."""
    if hasattr(source, "read") and hasattr(source, "numFrames"):
        return source
    if isinstance(source, (str, os.PathLike)) and os.fspath(source).lower().endswith(".dcd"):
        return DCDFile(source)
    if isinstance(source, (str, os.PathLike)) and os.fspath(source).lower().endswith(".xtc"):
        return XTCFile(source)
    return NumpyTrajectory(source, boxVectors)


class TrajectoryEvaluator:
    """
    Values (and effective masses) of collective variables over a stored trajectory.

    Parameters
    ----------
    system
        The system the variables were added to (``cv.addToSystem(system)``).
    variables
        The collective variables to report, in column order.
    value, emass
        Which columns to produce per variable (``cv_writer.py:80-91``; at least one).
    chunkFrames
        Frames per pipeline stage.  Default: as many as fit 256 MiB of pinned staging.
    device, precision
        As for :class:`BatchContext` (single precision only streams ``float32`` coordinates;
        ``"double"`` converts on the device).
    boxVectors
        Box used for every frame when the trajectory carries none.
    """

    def __init__(  # pylint: disable=too-many-arguments
        self,
        system: mm.System,
        variables: t.Sequence[t.Any],
        value: bool = True,
        emass: bool = False,
        chunkFrames: t.Optional[int] = None,
        device: t.Union[str, int, torch.device, None] = None,
        precision: str = "single",
        boxVectors: t.Any = None,
    ) -> None:
        if not (value or emass):
            raise ValueError("At least one of value or effective_mass must be True")
        forces = system.getForces()
        for variable in variables:
            if not any(force is variable for force in forces):
                raise RuntimeError("This force is not present in the given context.")
        self._system = system
        self._variables = list(variables)
        self._value, self._emass = value, emass
        self._chunk = chunkFrames
        self._device = device
        self._precision = precision
        self._box = boxVectors
        self._staging: t.Optional[t.Tuple[t.Any, t.List[torch.Tensor], t.List[torch.Tensor]]] = None

    # ---- table layout (reference: reporting/cv_writer.py:92-100) --------------------------------
    def getHeaders(self) -> t.List[str]:
        """Column headers in the reference's format."""
        headers = []
        for cv in self._variables:
            if self._value:
                headers.append(f"{cv.getName()} ({cv.getUnit().get_symbol()})")
            if self._emass:
                headers.append(f"emass[{cv.getName()}] ({cv.getMassUnit().get_symbol()})")
        return headers

    def chunkFrames(self, num_atoms: int) -> int:
        """Frames per pipeline stage for a trajectory of ``num_atoms`` atoms."""
        if self._chunk is not None:
            return max(1, int(self._chunk))
        return max(1, (256 << 20) // (12 * max(num_atoms, 1)))

    def evaluate(
        self, source: t.Any, start: int = 0, stop: t.Optional[int] = None
    ) -> t.Dict[str, np.ndarray]:
        """
        Evaluate frames ``start .. stop-1`` of ``source`` (a path, array or trajectory object).
        Returns ``{header: float64 array [frames]}`` in the order of :meth:`getHeaders`, in the
        units named by the headers.
        """
        _native.load()
        if not torch.cuda.is_available():
            raise _native.NativeLibraryError(
                "no CUDA device is available; cvpack_b200 has no CPU fallback"
            )
        traj = open_trajectory(source, None)
        num_atoms = traj.numAtoms
        if num_atoms != self._system.getNumParticles():
            raise ValueError(
                f"the trajectory has {num_atoms} atoms, the system {self._system.getNumParticles()}"
            )
        stop = len(traj) if stop is None else min(int(stop), len(traj))
        start = max(0, int(start))
        total = max(0, stop - start)
        headers = self.getHeaders()
        if total == 0:
            return {h: np.zeros(0) for h in headers}
        chunk = min(self.chunkFrames(num_atoms), total)
        ctx = BatchContext(self._system, None, None, precision=self._precision, device=self._device)
        device = ctx.getDevice()
        dtype = torch.float32 if self._precision == "single" else torch.float64
        per_frame_box = bool(traj.hasBox)
        shared_box = None
        if not per_frame_box:
            shared_box = self._box if self._box is not None else self._system.getDefaultPeriodicBoxVectors()

        depth = 2
        # pinned staging is expensive to allocate (page locking): kept across calls
        key = (chunk, num_atoms)
        if self._staging is None or self._staging[0] != key:
            self._staging = (
                key,
                [torch.empty((chunk, num_atoms, 3), dtype=torch.float32, pin_memory=True) for _ in range(depth)],
                [torch.empty((chunk, 3, 3), dtype=torch.float64, pin_memory=True) for _ in range(depth)],
            )
        _, host_pos, host_box = self._staging
        with torch.cuda.device(device):
            dev_pos = [torch.empty((chunk, num_atoms, 3), dtype=torch.float32, device=device) for _ in range(depth)]
            dev_box = [torch.empty((chunk, 3, 3), dtype=torch.float64, device=device) for _ in range(depth)]
            results = torch.zeros((len(headers), total), dtype=dtype, device=device)
            gradient = (
                torch.empty((chunk, num_atoms, 3), dtype=dtype, device=device) if self._emass else None
            )
            scratch = torch.empty(chunk, dtype=dtype, device=device)
            masses = (
                torch.as_tensor(self._system.getMasses(), dtype=torch.float64, device=device)
                if self._emass
                else None
            )
            copy_stream = torch.cuda.Stream(device=device)
            compute_stream = torch.cuda.current_stream(device)
            uploaded = [torch.cuda.Event() for _ in range(depth)]   # H2D of buffer k finished
            consumed = [torch.cuda.Event() for _ in range(depth)]   # kernels on buffer k finished

        chunks = [(s, min(chunk, start + total - s)) for s in range(start, start + total, chunk)]
        filled: "queue.Queue[t.Tuple[int, t.Optional[BaseException]]]" = queue.Queue()
        free = [threading.Semaphore(1) for _ in range(depth)]  # pinned buffer k may be refilled

        def reader() -> None:
            try:
                for index, (first, count) in enumerate(chunks):
                    k = index % depth
                    free[k].acquire()
                    traj.read(
                        first, count, host_pos[k].numpy()[:count],
                        host_box[k].numpy()[:count] if per_frame_box else None,
                    )
                    filled.put((index, None))
            except BaseException as exc:  # pylint: disable=broad-except
                filled.put((-1, exc))

        thread = threading.Thread(target=reader, name="cvpack-trajectory-reader", daemon=True)
        thread.start()
        pending: t.List[t.Tuple[int, torch.cuda.Event]] = []
        cvp_dtype = _native.CVP_F32 if self._precision == "single" else _native.CVP_F64
        try:
            with torch.cuda.device(device):
                for index, (first, count) in enumerate(chunks):
                    got, error = filled.get()
                    if error is not None:
                        raise error
                    assert got == index
                    k = index % depth
                    with torch.cuda.stream(copy_stream):
                        copy_stream.wait_event(consumed[k])  # (no-op before its first record)
                        dev_pos[k][:count].copy_(host_pos[k][:count], non_blocking=True)
                        if per_frame_box:
                            dev_box[k][:count].copy_(host_box[k][:count], non_blocking=True)
                        uploaded[k].record(copy_stream)
                    pending.append((k, uploaded[k]))
                    # the pinned buffer is free again once its upload has finished
                    while len(pending) > 1:
                        done_k, event = pending.pop(0)
                        event.synchronize()
                        free[done_k].release()
                    compute_stream.wait_event(uploaded[k])
                    pos = dev_pos[k][:count] if dtype == torch.float32 else dev_pos[k][:count].to(dtype)
                    ctx.setPositions(pos)
                    if per_frame_box:
                        ctx.setPeriodicBoxVectors(dev_box[k][:count])
                    elif shared_box is not None and index == 0:
                        ctx.setPeriodicBoxVectors(shared_box)
                    lo = first - start
                    row = 0
                    for cv in self._variables:
                        out_value = results[row, lo:lo + count] if self._value else scratch[:count]
                        grad = gradient[:count] if gradient is not None else None
                        ctx.evaluateInto(cv, out_value, grad)
                        row += int(self._value)
                        if self._emass:
                            _native.effective_mass(
                                cvp_dtype, grad.data_ptr(), masses.data_ptr(), count, num_atoms,
                                results[row, lo:lo + count].data_ptr(), compute_stream.cuda_stream,
                            )
                            row += 1
                    consumed[k].record(compute_stream)
                for done_k, event in pending:
                    event.synchronize()
                    free[done_k].release()
                table = results.double().cpu().numpy()
        finally:
            for semaphore in free:  # let a blocked reader run to its end
                semaphore.release()
            thread.join(timeout=60)
        return {header: table[i] for i, header in enumerate(headers)}

    def writeCSV(
        self,
        file: t.Union[str, os.PathLike, t.TextIO],
        table: t.Dict[str, np.ndarray],
        firstStep: int = 0,
        interval: int = 1,
        separator: str = ",",
    ) -> None:
        """
        Write a table returned by :meth:`evaluate` the way the reference's ``StateDataReporter``
        does with ``step=True`` (``#"Step","phi (rad)",...`` then one row per frame).
        """
        headers = list(table.keys())
        columns = [np.asarray(table[h]) for h in headers]
        rows = len(columns[0]) if columns else 0
        own = isinstance(file, (str, os.PathLike))
        stream = open(os.fspath(file), "w", encoding="utf-8") if own else file
        try:
            stream.write("#" + separator.join(f'"{h}"' for h in ["Step"] + headers) + "\n")
            for i in range(rows):
                fields = [str(firstStep + i * interval)]
                for column in columns:
                    v = float(column[i])
                    fields.append("inf" if math.isinf(v) and v > 0 else repr(v))
                stream.write(separator.join(fields) + "\n")
        finally:
            if own:
                stream.close()
