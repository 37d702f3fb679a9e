"""
Evaluation of collective variables that are functions of other collective variables
(:class:`PathInCVSpace`, :class:`MetaCollectiveVariable`).

The reference nests the inner variables in an ``openmm.CustomCVForce``
(``/root/reference/cvpack/base_path_cv.py:61-62``, ``meta_collective_variable.py:127-130``): OpenMM
evaluates every inner force once per frame and applies the chain rule.  Here every inner variable
runs its own batched CUDA kernels **once** -- value and gradient in the same call, into one slab of a
``[nv, B, N, 3]`` buffer -- the combination kernel yields ``d value / d cv_j`` per frame, and
``cvp_axpy_frames`` accumulates ``sum_j coef_j[b] * grad_j[b]``.  When the slab would exceed
``MAX_INNER_GRADIENT_BYTES`` the inner variables are evaluated a second time, one after the other,
into a single reused buffer (values first, gradients after the combination): memory instead of time.
"""

from __future__ import annotations

import typing as t

import torch

from . import _native

MAX_INNER_GRADIENT_BYTES = 2 << 30

Combine = t.Callable[[torch.Tensor, torch.Tensor, t.Optional[torch.Tensor], int], None]


def evaluate_composite(  # pylint: disable=too-many-arguments,too-many-locals
    context: t.Any,
    variables: t.Sequence[t.Any],
    combine: Combine,
    num_partials: int,
    need_gradient: bool,
    positions: t.Optional[torch.Tensor],
    out: t.Optional[t.Tuple[torch.Tensor, t.Optional[torch.Tensor]]],
) -> t.Tuple[torch.Tensor, t.Optional[torch.Tensor], torch.Tensor, t.Optional[torch.Tensor]]:
    """
    Returns ``(value [B], gradient [B, N, 3] or None, inner values [nv, B], partials or None)``.

    ``combine(values, out_value, out_partials, stream)`` launches the kernel that turns the inner
    values ``[nv, B]`` into the composite value ``[B]`` and ``d value / d input`` ``[num_partials, B]``
    (the first ``nv`` rows belong to the inner variables).
    """
    device = context.getDevice()
    dtype = context._torch_dtype  # pylint: disable=protected-access
    cvp_dtype = context._cvp_dtype  # pylint: disable=protected-access
    pos = context.getPositions() if positions is None else positions
    with torch.cuda.device(device):
        pos = pos.to(device).contiguous()
        num_frames, num_atoms = int(pos.shape[0]), int(pos.shape[1])
        nv = len(variables)
        stream = torch.cuda.current_stream(device).cuda_stream
        values = torch.empty((nv, num_frames), dtype=dtype, device=device)
        frame_bytes = num_atoms * 3 * values.element_size()
        keep_all = need_gradient and nv * num_frames * frame_bytes <= MAX_INNER_GRADIENT_BYTES
        inner_grads = None
        if keep_all:
            inner_grads = torch.empty((nv, num_frames, num_atoms, 3), dtype=dtype, device=device)
        for j, variable in enumerate(variables):
            context._run(  # pylint: disable=protected-access
                variable, keep_all, positions=pos,
                out=(values[j], inner_grads[j] if keep_all else None),
            )
        value = torch.empty(num_frames, dtype=dtype, device=device)
        partials = (
            torch.empty((num_partials, num_frames), dtype=dtype, device=device) if need_gradient else None
        )
        combine(values, value, partials, stream)
        grad = None
        if need_gradient:
            grad = torch.empty((num_frames, num_atoms, 3), dtype=dtype, device=device)
            scratch_value = None if keep_all else torch.empty(num_frames, dtype=dtype, device=device)
            scratch = None if keep_all else torch.empty((num_frames, num_atoms, 3), dtype=dtype, device=device)
            for j, variable in enumerate(variables):
                if keep_all:
                    inner = inner_grads[j]
                else:
                    context._run(variable, True, positions=pos, out=(scratch_value, scratch))  # pylint: disable=protected-access
                    inner = scratch
                _native.axpy_frames(
                    cvp_dtype, partials[j].data_ptr(), inner.data_ptr(), grad.data_ptr(), num_frames,
                    num_atoms, j > 0, stream,
                )
            if nv == 0:
                grad.zero_()
    if out is not None:
        out[0].copy_(value)
        if need_gradient:
            out[1].copy_(grad)
        return out[0], out[1], values, partials
    if context._residency == "host" and positions is None:  # pylint: disable=protected-access
        value = value.cpu().pin_memory()
        grad = None if grad is None else grad.cpu().pin_memory()
    return value, grad, values, partials
