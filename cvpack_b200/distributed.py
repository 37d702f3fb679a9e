"""
Multi-GPU plumbing: one process per GPU, frames sharded by rank, results gathered over
``torch.distributed`` (NCCL on GPUs, gloo on CPU for tests).

Frames (or metadynamics walkers) are independent -- every reference evaluation is a single-context
call (``/root/reference/cvpack/utils.py:140``) -- so the batch dimension is split contiguously
across ranks and the data path needs no collective at all.  The only exchange is the optional
all-gather of results: values ``[B]`` (tiny) and, when asked for, gradients, either dense
``[B, N, 3]`` or restricted to the atoms that can carry a non-zero gradient (``compact=True``),
which is what makes the gather cheap for CVs that touch a small part of a large system.
"""

from __future__ import annotations

import ctypes as C
import typing as t

import torch
import torch.distributed as dist

from . import _native
from .batch import shard_frames


def _world(group: t.Optional[t.Any]) -> t.Tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def local_shard(num_frames: int, group: t.Optional[t.Any] = None) -> t.Tuple[int, int]:
    """``[begin, end)`` of the frames owned by the calling rank."""
    rank, world = _world(group)
    return shard_frames(num_frames, rank, world)


def all_gather_frames(
    local: torch.Tensor, num_frames: int, group: t.Optional[t.Any] = None
) -> torch.Tensor:
    """
    Concatenate per-rank results along the frame dimension.  ``local`` holds the calling rank's
    frames ``[b_local, ...]`` (shards from :func:`local_shard`, possibly of unequal length);
    returns ``[num_frames, ...]`` on every rank.
    """
    rank, world = _world(group)
    if world == 1:
        return local
    counts = [shard_frames(num_frames, r, world) for r in range(world)]
    sizes = [end - begin for begin, end in counts]
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} frames, expected {sizes[rank]}")
    tail = tuple(local.shape[1:])
    if len(set(sizes)) == 1:
        out = torch.empty((num_frames,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # unequal shards: pad to the largest, gather, trim
    largest = max(sizes)
    padded = torch.zeros((largest,) + tail, dtype=local.dtype, device=local.device)
    padded[: sizes[rank]] = local
    out = torch.empty((world * largest,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    pieces = [out[r * largest : r * largest + sizes[r]] for r in range(world)]
    return torch.cat(pieces, dim=0)


def all_gather_gradients(
    local_gradient: torch.Tensor,
    num_frames: int,
    active_atoms: t.Optional[torch.Tensor] = None,
    group: t.Optional[t.Any] = None,
) -> torch.Tensor:
    """
    Gather gradients ``[b_local, N, 3]`` from all ranks.  With ``active_atoms`` (sorted indices of
    the atoms that can carry a non-zero gradient, ``cvp_active_atoms``) only ``[b, n_active, 3]``
    crosses the links and the dense ``[num_frames, N, 3]`` tensor is rebuilt locally.
    """
    if active_atoms is None:
        return all_gather_frames(local_gradient, num_frames, group)
    index = active_atoms.to(local_gradient.device, dtype=torch.long)
    compact = local_gradient.index_select(1, index)
    gathered = all_gather_frames(compact, num_frames, group)
    dense = torch.zeros(
        (num_frames,) + tuple(local_gradient.shape[1:]),
        dtype=local_gradient.dtype,
        device=local_gradient.device,
    )
    dense.index_copy_(1, index, gathered)
    return dense


class _DevicePointer:  # pylint: disable=too-few-public-methods
    """Adapter that lets torch alias raw device memory (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape: t.Tuple[int, ...], typestr: str) -> None:
        self.__cuda_array_interface__ = {
            "shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None,
        }


class PeerGather:
    """
    All-gather of gradients ``[b_local, N, 3] -> [num_frames, N, 3]`` through the library's own
    peer-memory collective (``csrc/peer.cu``): every rank owns a gather buffer, maps the others'
    through CUDA IPC, and :meth:`push` writes a slab of local frames into *every* rank's buffer with
    one kernel (remote stores over NVLink).  With ``active_atoms`` only those atoms travel and land
    in place -- compaction, transfer and re-expansion are one pass.  Pushes are ordered on the
    current stream, so chunks can be pushed while the next chunk is evaluated; :meth:`finish` is a
    stream-ordered barrier (no host synchronisation).

    One process per GPU on one node; ``group`` is only used to exchange the 64-byte IPC handles.
    ``depth`` buffers are used in turn (2 lets a step be consumed while the next one is gathered).
    """

    def __init__(  # pylint: disable=too-many-arguments
        self, num_frames: int, num_atoms: int, dtype: torch.dtype, device: torch.device,
        group: t.Optional[t.Any] = None, depth: int = 1,
    ) -> None:
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("gradients are float32 or float64")
        self._rank, self._world = _world(group)
        self._shape = (int(num_frames), int(num_atoms), 3)
        self._dtype = dtype
        self._cvp_dtype = _native.CVP_F32 if dtype == torch.float32 else _native.CVP_F64
        self._device = torch.device(device)
        self._buffers: t.List[int] = []
        self._tensors: t.List[torch.Tensor] = []
        self._turn = 0
        self._run_cache: t.Dict[t.Tuple[int, int], t.Optional[t.List[t.Tuple[int, int]]]] = {}
        lib = _native.load()
        element = 4 if dtype == torch.float32 else 8
        nbytes = self._shape[0] * self._shape[1] * 3 * element
        with torch.cuda.device(self._device):
            for _ in range(max(1, depth)):
                handle = C.c_void_p()
                _native.check(lib.cvp_peer_buffer_create(nbytes, C.byref(handle)))
                self._buffers.append(handle.value)
                ipc = C.create_string_buffer(64)
                _native.check(lib.cvp_peer_buffer_handle(handle.value, ipc))
                handles: t.List[t.Any] = [None] * self._world
                if self._world > 1:
                    dist.all_gather_object(handles, ipc.raw, group=group)
                else:
                    handles = [ipc.raw]
                packed = C.create_string_buffer(b"".join(handles), 64 * self._world)
                _native.check(lib.cvp_peer_buffer_open(handle.value, self._world, self._rank, packed))
                ptr, size = C.c_void_p(), C.c_int64()
                _native.check(lib.cvp_peer_buffer_local(handle.value, C.byref(ptr), C.byref(size)))
                view = torch.as_tensor(
                    _DevicePointer(ptr.value, self._shape, "<f4" if dtype == torch.float32 else "<f8"),
                    device=self._device,
                )
                self._tensors.append(view)
        if self._world > 1:
            dist.barrier(group=group)  # every rank has mapped every buffer before anyone pushes

    def tensor(self) -> torch.Tensor:
        """The gather buffer of the current turn as a ``[num_frames, N, 3]`` tensor (aliases it)."""
        return self._tensors[self._turn]

    def push(self, gradient: torch.Tensor, first_frame: int, active_atoms: t.Optional[torch.Tensor] = None) -> None:
        """Write ``gradient`` ``[b, N, 3]`` (frames ``first_frame ...`` of the batch) into every rank's
        buffer, on the current stream.  ``active_atoms``: int32 CUDA tensor of the atoms to move."""
        if gradient.dtype != self._dtype or not gradient.is_cuda or not gradient.is_contiguous():
            raise ValueError("gradient must be a contiguous CUDA tensor of the buffer's dtype")
        if tuple(gradient.shape[1:]) != self._shape[1:]:
            raise ValueError(f"gradient must have shape [b, {self._shape[1]}, 3]")
        stream = torch.cuda.current_stream(self._device).cuda_stream
        active_ptr, count = None, 0
        if active_atoms is not None:
            if active_atoms.dtype != torch.int32 or not active_atoms.is_cuda:
                raise ValueError("active_atoms must be an int32 CUDA tensor")
            runs = self._runs(active_atoms)
            if runs is not None:
                # the active atoms are a few contiguous runs: wide row copies instead of a scatter
                for first_atom, atom_count in runs:
                    _native.check(
                        _native.load().cvp_peer_push_rows(
                            self._buffers[self._turn], self._cvp_dtype, gradient.data_ptr(),
                            int(first_frame), int(gradient.shape[0]), self._shape[1], first_atom,
                            atom_count, stream,
                        )
                    )
                return
            active_ptr, count = active_atoms.data_ptr(), int(active_atoms.numel())
        _native.check(
            _native.load().cvp_peer_push(
                self._buffers[self._turn], self._cvp_dtype, gradient.data_ptr(), int(first_frame),
                int(gradient.shape[0]), self._shape[1], active_ptr, count, stream,
            )
        )

    def _runs(self, active_atoms: torch.Tensor) -> t.Optional[t.List[t.Tuple[int, int]]]:
        """Contiguous runs (first atom, count) of a sorted atom list, or None if there are many."""
        key = (active_atoms.data_ptr(), int(active_atoms.numel()))
        if key not in self._run_cache:
            atoms = active_atoms.cpu().numpy()
            breaks = [0] + [int(k) for k in (atoms[1:] != atoms[:-1] + 1).nonzero()[0] + 1] + [len(atoms)]
            runs = [(int(atoms[a]), b - a) for a, b in zip(breaks[:-1], breaks[1:]) if b > a]
            sorted_unique = bool((atoms[1:] > atoms[:-1]).all()) if len(atoms) > 1 else True
            self._run_cache[key] = runs if sorted_unique and len(runs) <= 8 else None
        return self._run_cache[key]

    def finish(self, stream: t.Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        """Stream-ordered barrier: after it (on that stream) the buffer holds every rank's frames.
        Returns the gathered tensor and moves on to the next buffer of the ring."""
        cuda_stream = (stream or torch.cuda.current_stream(self._device)).cuda_stream
        lib = _native.load()
        _native.check(lib.cvp_peer_signal(self._buffers[self._turn], cuda_stream))
        _native.check(lib.cvp_peer_wait(self._buffers[self._turn], cuda_stream))
        gathered = self._tensors[self._turn]
        self._turn = (self._turn + 1) % len(self._buffers)
        return gathered

    def timed_out(self) -> bool:
        """Whether a wait gave up (a rank never signalled); synchronises the device."""
        flag = C.c_int()
        result = False
        for handle in self._buffers:
            _native.check(_native.load().cvp_peer_status(handle, C.byref(flag)))
            result = result or bool(flag.value)
        return result

    def close(self) -> None:
        """Unmap the peers' buffers and free this rank's (all ranks, after a barrier)."""
        torch.cuda.synchronize(self._device)
        self._tensors.clear()
        lib = _native.load()
        for handle in self._buffers:
            lib.cvp_peer_buffer_destroy(handle)
        self._buffers = []
