"""
Walker coupling: the step *after* the hot path.

In the reference a collective variable reaches the integrator through OpenMM: the user wraps the
CV in ``openmm.app.BiasVariable`` and ``Metadynamics`` adds a ``CustomCVForce`` whose force on atom
``i`` is ``-dV/ds * ds/dr_i`` (README ``:32-34``; the chain rule of ``CustomCVForce``,
``meta_collective_variable.py:114-141`` for functions of several CVs).  With a batch of walkers the
same chain rule is one fused pass per variable over the ``[B, N, 3]`` gradient:

    ``forces[b] -= dV/ds_k[b] * grad s_k[b]``

done in place on the device by ``cvp_axpy_frames`` (``include/cvpack_b200.h``), so that neither the
gradients nor the forces leave the GPU or get gathered across ranks: each rank owns its walkers.

:class:`HarmonicBias` and :class:`GaussianHillsBias` supply ``V`` and ``dV/ds`` for ``[B]`` values
(tiny elementwise work on the batch container); :func:`applyBias` is the coupling itself.
"""

from __future__ import annotations

import typing as t

import torch

from . import _native
from .batch import BatchContext

_CVP_DTYPE = {torch.float32: _native.CVP_F32, torch.float64: _native.CVP_F64}


def accumulateBiasForces(
    forces: torch.Tensor, gradient: torch.Tensor, dbias: torch.Tensor, accumulate: bool = True
) -> torch.Tensor:
    """
    ``forces[b] = (accumulate ? forces[b] : 0) - dbias[b] * gradient[b]`` in place, on the device.

    ``forces`` and ``gradient`` are ``[B, N, 3]`` CUDA tensors of the same dtype, ``dbias`` is
    ``dV/ds`` per walker, ``[B]``.
    """
    if not (forces.is_cuda and gradient.is_cuda):
        raise ValueError("forces and gradient must be CUDA tensors; there is no CPU path")
    if forces.shape != gradient.shape or forces.dim() != 3 or forces.shape[2] != 3:
        raise ValueError("forces and gradient must both have shape [B, N, 3]")
    if forces.dtype != gradient.dtype or forces.dtype not in _CVP_DTYPE:
        raise ValueError("forces and gradient must share a float32 or float64 dtype")
    if not (forces.is_contiguous() and gradient.is_contiguous()):
        raise ValueError("forces and gradient must be contiguous")
    num_frames, num_atoms = int(forces.shape[0]), int(forces.shape[1])
    if tuple(dbias.shape) != (num_frames,):
        raise ValueError(f"dbias must have shape [{num_frames}]")
    with torch.cuda.device(forces.device):
        coef = (-dbias).to(device=forces.device, dtype=forces.dtype).contiguous()
        stream = torch.cuda.current_stream(forces.device).cuda_stream
        _native.axpy_frames(
            _CVP_DTYPE[forces.dtype], coef.data_ptr(), gradient.data_ptr(), forces.data_ptr(),
            num_frames, num_atoms, accumulate, stream,
        )
    return forces


class HarmonicBias:
    """``V(s) = k/2 (s - s0)^2``, with the minimum-image difference for a periodic variable."""

    def __init__(self, forceConstant: float, center: t.Any, period: t.Optional[float] = None):
        self.forceConstant = float(forceConstant)
        self.center = center
        self.period = None if period is None else float(period)

    def _delta(self, values: torch.Tensor) -> torch.Tensor:
        center = torch.as_tensor(self.center, dtype=values.dtype, device=values.device)
        delta = values - center
        if self.period is not None:
            delta = delta - self.period * torch.round(delta / self.period)
        return delta

    def energy(self, values: torch.Tensor) -> torch.Tensor:
        """``V`` per walker."""
        return 0.5 * self.forceConstant * self._delta(values) ** 2

    def derivative(self, values: torch.Tensor) -> torch.Tensor:
        """``dV/ds`` per walker."""
        return self.forceConstant * self._delta(values)


class GaussianHillsBias:
    """
    Metadynamics bias of one variable shared by all walkers: ``V(s) = sum_h w_h
    exp(-(s - s_h)^2 / (2 sigma^2))`` (multiple-walker metadynamics: every walker deposits into and
    feels the same set of hills).
    """

    def __init__(self, sigma: float, period: t.Optional[float] = None) -> None:
        if sigma <= 0:
            raise ValueError("sigma must be positive")
        self.sigma = float(sigma)
        self.period = None if period is None else float(period)
        self._centers: t.Optional[torch.Tensor] = None
        self._heights: t.Optional[torch.Tensor] = None

    def getNumHills(self) -> int:
        """Hills deposited so far."""
        return 0 if self._centers is None else int(self._centers.numel())

    def addHills(self, centers: torch.Tensor, heights: t.Any) -> None:
        """Deposit one hill per entry of ``centers`` (e.g. the walkers' current values)."""
        centers = centers.detach().to(torch.float64).flatten()
        heights = torch.as_tensor(heights, dtype=torch.float64, device=centers.device).expand_as(centers)
        if self._centers is None:
            self._centers, self._heights = centers.clone(), heights.clone()
        else:
            self._centers = torch.cat([self._centers.to(centers.device), centers])
            self._heights = torch.cat([self._heights.to(centers.device), heights])

    def _terms(self, values: torch.Tensor) -> t.Tuple[torch.Tensor, torch.Tensor]:
        assert self._centers is not None and self._heights is not None
        s = values.detach().to(torch.float64)
        delta = s[:, None] - self._centers.to(s.device)[None, :]
        if self.period is not None:
            delta = delta - self.period * torch.round(delta / self.period)
        gauss = self._heights.to(s.device)[None, :] * torch.exp(-0.5 * (delta / self.sigma) ** 2)
        return delta, gauss

    def energy(self, values: torch.Tensor) -> torch.Tensor:
        """``V`` per walker."""
        if self._centers is None:
            return torch.zeros_like(values)
        _, gauss = self._terms(values)
        return gauss.sum(dim=1).to(values.dtype)

    def derivative(self, values: torch.Tensor) -> torch.Tensor:
        """``dV/ds`` per walker."""
        if self._centers is None:
            return torch.zeros_like(values)
        delta, gauss = self._terms(values)
        return (-(delta / self.sigma**2) * gauss).sum(dim=1).to(values.dtype)


def applyBias(
    context: BatchContext,
    variables: t.Sequence[t.Any],
    biases: t.Sequence[t.Any],
    forces: t.Optional[torch.Tensor] = None,
) -> t.Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """
    Evaluate ``variables`` on the walkers of ``context`` and add the bias forces
    ``-dV_k/ds_k * grad s_k`` to ``forces`` (``[B, N, 3]`` on the context's device; allocated and
    zero-filled when ``None``).  ``biases[k]`` provides ``energy(values)`` and
    ``derivative(values)`` for ``[B]`` tensors in the unit of ``variables[k]``.

    Returns ``(forces, values [K, B], bias energy [B])``; nothing leaves the device.
    """
    if len(variables) != len(biases):
        raise ValueError("one bias per variable is required")
    num_frames = context.getNumFrames()
    num_atoms = context.getSystem().getNumParticles()
    device = context.getDevice()
    dtype = torch.float32 if context.getPrecision() == "single" else torch.float64
    with torch.cuda.device(device):
        accumulate = forces is not None
        if forces is None:
            forces = torch.empty((num_frames, num_atoms, 3), dtype=dtype, device=device)
        values = torch.empty((len(variables), num_frames), dtype=dtype, device=device)
        gradient = torch.empty((num_frames, num_atoms, 3), dtype=dtype, device=device)
        energy = torch.zeros(num_frames, dtype=dtype, device=device)
        if not variables and not accumulate:
            forces.zero_()
        for k, (cv, bias) in enumerate(zip(variables, biases)):
            context.evaluateInto(cv, values[k], gradient)
            energy += bias.energy(values[k])
            accumulateBiasForces(forces, gradient, bias.derivative(values[k]), accumulate or k > 0)
    return forces, values, energy
