"""
The batch container that takes the place of ``openmm.Context`` on the evaluation path.

The reference evaluates one frame per ``context.getState(getEnergy/getForces, groups=1<<g)``
(``/root/reference/cvpack/utils.py:140``).  A :class:`BatchContext` holds ``B`` frames at once --
positions ``[B, N, 3]``, periodic box vectors (one box or one per frame) and the system (masses,
registered collective variables) -- and every ``getValue`` / ``getGradient`` /
``getEffectiveMass`` call on a collective variable evaluates all of them with the CUDA library.

PyTorch tensors are only the container: device memory, streams, dtype.  Two residencies:

``device`` (default)
    positions live in HBM; results are CUDA tensors (``cvp_evaluate``).
``host``
    positions live in (pinned) host memory; each evaluation streams chunks of frames through
    host->device copy, kernels and device->host copy (``cvp_evaluate_host``); results are pinned
    CPU tensors.  This is the end-to-end path the benchmark's ``e2e`` number measures.
"""

from __future__ import annotations

import typing as t

import numpy as np
import torch

from . import _native, mm, mmunit
from .units import value_in_md_units

_DTYPES = {"single": (torch.float32, _native.CVP_F32), "double": (torch.float64, _native.CVP_F64)}


def _as_tensor(data: t.Any) -> torch.Tensor:
    data = value_in_md_units(data)
    if isinstance(data, torch.Tensor):
        return data
    if isinstance(data, (list, tuple)) and data and mmunit.is_quantity(data[0]):
        data = [value_in_md_units(item) for item in data]
    return torch.as_tensor(np.asarray(data))


class BatchContext:
    """
    A batch of configurations of one :class:`~cvpack_b200.mm.System`.

    Parameters
    ----------
    system
        The system (masses; the CVs added to it are the ones that can be evaluated).
    positions
        Coordinates, ``[B, N, 3]`` or ``[N, 3]`` (one frame), nm or a length quantity; numpy array
        or torch tensor.
    boxVectors
        ``None``, a ``[3, 3]`` matrix of row vectors (one box for all frames), ``[3]`` edge lengths
        of a rectangular box, or the batched forms ``[B, 3]`` / ``[B, 3, 3]``.  Defaults to the
        system's default box.
    precision
        ``"single"`` (FP32 coordinates and results; sums that need it are carried in FP64) or
        ``"double"``.
    device
        CUDA device (``"cuda:0"``, index or ``torch.device``).
    residency
        ``"device"`` or ``"host"`` (see module docstring).
    """

    def __init__(  # pylint: disable=too-many-arguments
        self,
        system: mm.System,
        positions: t.Any = None,
        boxVectors: t.Any = None,
        precision: str = "single",
        device: t.Union[str, int, torch.device, None] = None,
        residency: str = "device",
    ) -> None:
        if precision not in _DTYPES:
            raise ValueError("precision must be 'single' or 'double'")
        if residency not in ("device", "host"):
            raise ValueError("residency must be 'device' or 'host'")
        self._system = system
        self._precision = precision
        self._torch_dtype, self._cvp_dtype = _DTYPES[precision]
        if device is None:
            device = "cuda:0"
        if isinstance(device, int):
            device = f"cuda:{device}"
        self._device = torch.device(device)
        if self._device.type != "cuda":
            raise ValueError("BatchContext evaluates on a CUDA device; there is no CPU path")
        self._residency = residency
        self._positions: t.Optional[torch.Tensor] = None
        self._box: t.Optional[torch.Tensor] = None
        self._box_mode = _native.CVP_BOX_NONE
        self._triclinic = False
        self._half_min_edge = float("inf")
        self._compiled: t.Dict[int, t.Tuple[t.Any, _native.NativeCV]] = {}
        self._masses_dev: t.Optional[torch.Tensor] = None
        self._last: t.Dict[int, t.Tuple[torch.Tensor, torch.Tensor]] = {}
        self._parameters: t.Dict[str, float] = {}
        if positions is not None:
            self.setPositions(positions)
        if boxVectors is None and system.getDefaultPeriodicBoxVectors() is not None:
            boxVectors = system.getDefaultPeriodicBoxVectors()
        if boxVectors is not None:
            self.setPeriodicBoxVectors(boxVectors)

    # ---- state ----------------------------------------------------------------------------------
    def getSystem(self) -> mm.System:
        """The system of this context."""
        return self._system

    def getNumFrames(self) -> int:
        """Number of frames in the batch."""
        return 0 if self._positions is None else int(self._positions.shape[0])

    def getPrecision(self) -> str:
        """``"single"`` or ``"double"``."""
        return self._precision

    def getDevice(self) -> torch.device:
        """The CUDA device results are computed on."""
        return self._device

    def _place(self, tensor: torch.Tensor) -> torch.Tensor:
        tensor = tensor.to(self._torch_dtype)
        if self._residency == "device":
            return tensor.to(self._device).contiguous()
        tensor = tensor.cpu().contiguous()
        if not tensor.is_pinned() and torch.cuda.is_available():
            tensor = tensor.pin_memory()
        return tensor

    def setPositions(self, positions: t.Any) -> None:
        """Set the coordinates of all frames (``[B, N, 3]`` or ``[N, 3]``)."""
        tensor = _as_tensor(positions)
        if tensor.dim() == 2:
            tensor = tensor.unsqueeze(0)
        num_atoms = self._system.getNumParticles()
        if tensor.dim() != 3 or tensor.shape[1] != num_atoms or tensor.shape[2] != 3:
            raise ValueError(
                f"positions must have shape [B, {num_atoms}, 3]; got {list(tensor.shape)}"
            )
        self._positions = self._place(tensor)
        self._last.clear()

    def getPositions(self) -> torch.Tensor:
        """The ``[B, N, 3]`` coordinate tensor held by this context."""
        if self._positions is None:
            raise RuntimeError("positions have not been set")
        return self._positions

    def setPeriodicBoxVectors(self, a: t.Any, b: t.Any = None, c: t.Any = None) -> None:
        """
        Set the periodic box: three vectors ``a, b, c`` (like ``Context.setPeriodicBoxVectors``), a
        ``[3, 3]`` matrix, ``[3]`` edge lengths, or the batched forms ``[B, 3]`` / ``[B, 3, 3]``.
        Boxes must be in OpenMM's reduced form (a along x, b in the xy plane).
        """
        if b is not None:
            tensor = torch.stack([_as_tensor(v).to(torch.float64).flatten() for v in (a, b, c)])
        else:
            tensor = _as_tensor(a).to(torch.float64)
        batched = False
        if tensor.dim() == 1:
            tensor = torch.diag(tensor)
        elif tensor.dim() == 2 and tensor.shape != (3, 3):
            tensor = torch.diag_embed(tensor)
            batched = True
        elif tensor.dim() == 3:
            batched = True
        if tensor.shape[-2:] != (3, 3):
            raise ValueError(f"cannot interpret box vectors of shape {list(tensor.shape)}")
        flat = tensor.reshape(-1, 3, 3).cpu()
        upper = torch.stack([flat[:, 0, 1], flat[:, 0, 2], flat[:, 1, 2]])
        if bool((upper != 0).any()):
            raise ValueError("box vectors must be in reduced form: a=(ax,0,0), b=(bx,by,0)")
        if bool((torch.diagonal(flat, dim1=1, dim2=2) <= 0).any()):
            raise ValueError("box vectors must have positive diagonal elements")
        lower = torch.stack([flat[:, 1, 0], flat[:, 2, 0], flat[:, 2, 1]])
        self._triclinic = bool((lower != 0).any())
        self._box_mode = _native.CVP_BOX_PER_FRAME if batched else _native.CVP_BOX_SHARED
        self._box_host = flat.clone()
        self._half_min_edge = 0.5 * float(torch.diagonal(flat, dim1=1, dim2=2).min())
        self._box = self._place(tensor.reshape(-1, 9) if batched else tensor.reshape(9))
        self._last.clear()

    def setParameter(self, name: str, value: t.Any) -> None:
        """Set a named global parameter (like ``openmm.Context.setParameter``); it overrides the default
        of every :class:`MetaCollectiveVariable` parameter of that name."""
        self._parameters[name] = float(value_in_md_units(value))

    def getParameter(self, name: str, default: t.Any = None) -> float:
        """Value of a named global parameter (``default`` if it was never set)."""
        if name in self._parameters:
            return self._parameters[name]
        if default is None:
            raise KeyError(f"Called getParameter() with invalid parameter name: {name}")
        return float(default)

    def getPeriodicBoxVectors(self) -> t.Optional[torch.Tensor]:
        """Box vectors (``[9]`` or ``[B, 9]`` row-major) or None."""
        return self._box

    def reinitialize(self, preserveState: bool = True) -> None:
        """
        Drop compiled kernel specs so that CVs added to (or changed in) the system since they were
        first evaluated are picked up -- the analogue of ``Context.reinitialize`` the reference's
        tests call after ``addToSystem`` (``tests/test_cvpack.py:80``).
        """
        self._compiled.clear()
        self._last.clear()
        self._masses_dev = None
        if not preserveState:
            self._positions = None
            self._box = None
            self._box_mode = _native.CVP_BOX_NONE

    # ---- evaluation -----------------------------------------------------------------------------
    def _native_cv(self, cv: t.Any) -> _native.NativeCV:
        key = id(cv)
        entry = self._compiled.get(key)
        if entry is None or entry[0] is not cv:
            native = cv._createNative(self._system)
            self._compiled[key] = (cv, native)
            return native
        return entry[1]

    def _require_cuda(self) -> None:
        _native.load()
        if not torch.cuda.is_available():
            raise _native.NativeLibraryError(
                "no CUDA device is available; cvpack_b200 has no CPU fallback"
            )

    def _run(
        self,
        cv: t.Any,
        need_gradient: bool,
        positions: t.Optional[torch.Tensor] = None,
        out: t.Optional[t.Tuple[torch.Tensor, t.Optional[torch.Tensor]]] = None,
    ) -> t.Tuple[torch.Tensor, t.Optional[torch.Tensor]]:
        self._require_cuda()
        if hasattr(cv, "_evaluateComposite"):
            # a variable defined on other variables (PathInCVSpace): no single kernel spec
            if out is not None:
                pos = self.getPositions() if positions is None else positions
                self._check_out(
                    out, int(pos.shape[0]), int(pos.shape[1]), need_gradient,
                    self._residency == "host" and positions is None,
                )
            return cv._evaluateComposite(self, need_gradient, positions, out)
        pos = self.getPositions() if positions is None else positions
        num_frames, num_atoms = int(pos.shape[0]), int(pos.shape[1])
        native = self._native_cv(cv)
        self._check_cutoff(native)
        flags = _native.CVP_FLAG_TRICLINIC if self._triclinic else 0
        box_ptr = None if self._box is None else self._box.data_ptr()
        box_mode = self._box_mode if self._box is not None else _native.CVP_BOX_NONE
        if box_mode == _native.CVP_BOX_PER_FRAME and int(self._box.shape[0]) != num_frames:
            if positions is None:
                raise ValueError("the number of boxes differs from the number of frames")
            box_ptr = self._box[:num_frames].contiguous().data_ptr()
        on_host = self._residency == "host" and positions is None
        if out is not None:
            self._check_out(out, num_frames, num_atoms, need_gradient, on_host)
        if on_host:
            if out is not None:
                value, grad = out
            else:
                value = torch.empty(num_frames, dtype=self._torch_dtype, pin_memory=True)
                grad = (
                    torch.empty((num_frames, num_atoms, 3), dtype=self._torch_dtype, pin_memory=True)
                    if need_gradient
                    else None
                )
            native.evaluate_host(
                self._cvp_dtype, pos.data_ptr(), box_ptr, box_mode, num_frames, num_atoms,
                value.data_ptr(), None if grad is None else grad.data_ptr(), flags,
                self._device.index or 0,
            )
            return value, grad
        with torch.cuda.device(self._device):
            pos = pos.to(self._device)
            box = self._box
            if box is not None and not box.is_cuda:
                box = box.to(self._device)
                box_ptr = box.data_ptr()
            if out is not None:
                value, grad = out
            else:
                value = torch.empty(num_frames, dtype=self._torch_dtype, device=self._device)
                grad = (
                    torch.empty(
                        (num_frames, num_atoms, 3), dtype=self._torch_dtype, device=self._device
                    )
                    if need_gradient
                    else None
                )
            stream = torch.cuda.current_stream(self._device).cuda_stream
            native.evaluate(
                self._cvp_dtype, pos.data_ptr(), box_ptr, box_mode, num_frames, num_atoms,
                value.data_ptr(), None if grad is None else grad.data_ptr(), flags, stream,
            )
        return value, grad

    def _check_cutoff(self, native: _native.NativeCV) -> None:
        """OpenMM refuses a periodic cutoff larger than half the smallest box edge; so do we (the
        kernels would otherwise return a single-image sum silently)."""
        if self._box is None:
            return
        cutoff = native.periodic_cutoff()
        if cutoff > self._half_min_edge:
            raise ValueError(
                "The cutoff distance cannot be greater than half the periodic box size "
                f"(cutoff {cutoff:g} nm, smallest box edge {2 * self._half_min_edge:g} nm)."
            )

    def _check_out(
        self,
        out: t.Tuple[torch.Tensor, t.Optional[torch.Tensor]],
        num_frames: int,
        num_atoms: int,
        need_gradient: bool,
        on_host: bool,
    ) -> None:
        value, grad = out
        tensors = [(value, (num_frames,))]
        if need_gradient:
            if grad is None:
                raise ValueError("a gradient buffer is required")
            tensors.append((grad, (num_frames, num_atoms, 3)))
        for tensor, shape in tensors:
            if tuple(tensor.shape) != shape or tensor.dtype != self._torch_dtype:
                raise ValueError(f"output buffer must have shape {shape} and dtype {self._torch_dtype}")
            if not tensor.is_contiguous() or tensor.is_cuda == on_host:
                raise ValueError("output buffers must be contiguous and live where the positions do")

    def _evaluate(
        self, cv: t.Any, need_gradient: bool
    ) -> t.Tuple[torch.Tensor, t.Optional[torch.Tensor]]:
        return self._run(cv, need_gradient)

    def evaluateInto(
        self, cv: t.Any, value: torch.Tensor, gradient: t.Optional[torch.Tensor] = None
    ) -> None:
        """
        Evaluate a CV of this context's system into caller-owned buffers (``value`` ``[B]`` and,
        optionally, ``gradient`` ``[B, N, 3]``; CUDA tensors for device residency, pinned CPU
        tensors for host residency).  Lets a loop over trajectory chunks reuse its buffers.
        """
        if not any(force is cv for force in self._system.getForces()):
            raise RuntimeError("This force is not present in the given context.")
        self._run(cv, gradient is not None, out=(value, gradient))

    def _evaluateDetached(self, cv: t.Any) -> float:
        """Value of a CV (not necessarily in the system) at the first frame."""
        first = self.getPositions()[:1]
        if self._residency == "host":
            first = first.to(self._device)
        value, _ = self._run(cv, False, positions=first.contiguous())
        self._compiled.pop(id(cv), None)
        return float(value[0])

    def _effectiveMass(self, cv: t.Any) -> torch.Tensor:
        """
        ``1/sum_i |g_i|^2/m_i`` per frame (``utils.py:151-184``).  The ``[B, N, 3]`` gradient is an
        intermediate: it lives in a library-owned scratch buffer on the device and only ``[B]``
        numbers come back -- on a host-resident context nothing but the coordinates crosses PCIe.
        """
        self._require_cuda()
        if hasattr(cv, "_evaluateComposite"):
            # variables defined on other variables: gradient assembled by the composite, then reduced
            _, grad = self._run(cv, True)
            return self._massFromGradient(grad)
        pos = self.getPositions()
        num_frames, num_atoms = int(pos.shape[0]), int(pos.shape[1])
        native = self._native_cv(cv)
        self._check_cutoff(native)
        flags = _native.CVP_FLAG_TRICLINIC if self._triclinic else 0
        box_ptr = None if self._box is None else self._box.data_ptr()
        box_mode = self._box_mode if self._box is not None else _native.CVP_BOX_NONE
        if self._residency == "host":
            masses = np.ascontiguousarray(self._system.getMasses(), dtype=np.float64)
            value = torch.empty(num_frames, dtype=self._torch_dtype, pin_memory=True)
            emass = torch.empty(num_frames, dtype=self._torch_dtype, pin_memory=True)
            native.evaluate_with_mass_host(
                self._cvp_dtype, pos.data_ptr(), box_ptr, box_mode, num_frames, num_atoms,
                masses.ctypes.data, value.data_ptr(), emass.data_ptr(), None, flags,
                self._device.index or 0,
            )
            return emass
        with torch.cuda.device(self._device):
            masses_dev = self._massesOnDevice()
            value = torch.empty(num_frames, dtype=self._torch_dtype, device=self._device)
            emass = torch.empty(num_frames, dtype=self._torch_dtype, device=self._device)
            stream = torch.cuda.current_stream(self._device).cuda_stream
            native.evaluate_with_mass(
                self._cvp_dtype, pos.data_ptr(), box_ptr, box_mode, num_frames, num_atoms,
                masses_dev.data_ptr(), value.data_ptr(), emass.data_ptr(), flags, stream,
            )
        return emass

    def _massesOnDevice(self) -> torch.Tensor:
        if self._masses_dev is None:
            self._masses_dev = torch.as_tensor(
                self._system.getMasses(), dtype=torch.float64, device=self._device
            )
        return self._masses_dev

    def _massFromGradient(self, grad: torch.Tensor) -> torch.Tensor:
        with torch.cuda.device(self._device):
            masses_dev = self._massesOnDevice()
            grad_dev = grad.to(self._device, non_blocking=True)
            num_frames, num_atoms = int(grad_dev.shape[0]), int(grad_dev.shape[1])
            out = torch.empty(num_frames, dtype=self._torch_dtype, device=self._device)
            stream = torch.cuda.current_stream(self._device).cuda_stream
            _native.effective_mass(
                self._cvp_dtype, grad_dev.data_ptr(), masses_dev.data_ptr(), num_frames,
                num_atoms, out.data_ptr(), stream,
            )
        return out.cpu() if self._residency == "host" else out


def shard_frames(num_frames: int, rank: int, world_size: int) -> t.Tuple[int, int]:
    """
    Contiguous split of the batch dimension across ranks: rank ``r`` owns frames
    ``[begin, end)``; the first ``num_frames % world_size`` ranks get one extra frame.
    """
    if not 0 <= rank < world_size:
        raise ValueError("rank must be in [0, world_size)")
    base, extra = divmod(num_frames, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)
