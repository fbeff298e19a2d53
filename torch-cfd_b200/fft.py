"""Batched 2-D real transforms on the libtcfd FFT core, with ``torch.fft``'s semantics and layouts
(``rfft2`` / ``irfft2`` over the last two axes, "backward" normalisation), for the callers either side of the
fused step: post-processing of recorded trajectories (fno/data_gen/data_gen_Kolmogorov2d.py:178-188,
data_gen_McWilliams2d.py:154-165), forcing spectra (torch_cfd/equations.py:429-437) and initial conditions.
CUDA tensors only (square power-of-two grids 32..2048); no torch.fft, no cuFFT, no CPU fallback."""
from __future__ import annotations

import torch

from . import _lib

_PLANS = {}


def _plan(x: torch.Tensor, n: int, real_dtype: torch.dtype) -> "_lib.FFT2Plan":
    if x.device.type != "cuda":
        raise RuntimeError("torch-cfd_b200 runs its transforms on CUDA devices only (no CPU fallback); "
                           f"got a tensor on {x.device}")
    dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
    key = (dev, n, real_dtype)
    plan = _PLANS.get(key)
    if plan is None:
        with torch.cuda.device(dev):
            plan = _lib.FFT2Plan(_lib.load_library(), n, real_dtype)
        _PLANS[key] = plan
    return plan


def irfft2(x_hat: torch.Tensor) -> torch.Tensor:
    """== torch.fft.irfft2(x_hat): (*, n, n//2+1) complex64/128 -> (*, n, n) float32/64."""
    n = x_hat.shape[-2]
    real = torch.float32 if x_hat.dtype == torch.complex64 else torch.float64
    with torch.cuda.device(x_hat.device):
        return _plan(x_hat, n, real).irfft2(x_hat)


def rfft2(x: torch.Tensor) -> torch.Tensor:
    """== torch.fft.rfft2(x): (*, n, n) float32/64 -> (*, n, n//2+1) complex."""
    with torch.cuda.device(x.device):
        return _plan(x, x.shape[-1], x.dtype).rfft2(x)


def interpolate_bilinear(x: torch.Tensor, size: int, dtype: torch.dtype = None) -> torch.Tensor:
    """== F.interpolate(x.to(dtype), size=(size, size), mode="bilinear") for (*, n, n) fields."""
    if x.device.type != "cuda":
        raise RuntimeError(f"torch-cfd_b200: CUDA tensors only (no CPU fallback); got a tensor on {x.device}")
    with torch.cuda.device(x.device):
        return _lib.resample_bilinear(_lib.load_library(), x, int(size), x.dtype if dtype is None else dtype)
