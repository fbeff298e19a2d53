"""Spectral helper functions of hot path A (reference: torch_cfd/spectral.py:29-115).

These are the table builders and the point-wise spectral operators of the reference's public
API.  Inside the fused CUDA step the same arithmetic runs in the kernels
(csrc/ns2d_kernels.cuh); the functions here are used once per equation to build the tables (so
that their rounding is the reference's) and by callers that want a single operator.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .grids import Grid


def fft_mesh_2d(n: int, diam: float, device=None):
    k = [torch.fft.fftfreq(n, d=diam / n) for _ in range(2)]
    kx, ky = torch.meshgrid(k, indexing="ij")
    return kx.to(device), ky.to(device)


def fft_expand_dims(fft_mesh, batch_size: int):
    kx, ky = fft_mesh
    return tuple(z[None, :, :, None].expand(batch_size, *z.shape, 1) for z in (kx, ky))


def spectral_laplacian_2d(fft_mesh, device=None):
    """-4 pi^2 |k|^2 with the mean mode patched to 1 so that it can be inverted."""
    kx, ky = fft_mesh
    lap = -4 * (torch.pi**2) * (abs(kx) ** 2 + abs(ky) ** 2)
    lap[..., 0, 0] = 1
    return lap.to(device)


def spectral_curl_2d(vhat, rfft_mesh):
    uhat, vhat = vhat
    kx, ky = rfft_mesh
    return 2j * torch.pi * (vhat * kx - uhat * ky)


def spectral_div_2d(vhat, rfft_mesh):
    uhat, vhat = vhat
    kx, ky = rfft_mesh
    return 2j * torch.pi * (uhat * kx + vhat * ky)


def spectral_grad_2d(vhat, rfft_mesh):
    kx, ky = rfft_mesh
    return 2j * torch.pi * kx * vhat, 2j * torch.pi * ky * vhat


def spectral_rot_2d(vhat, rfft_mesh):
    gx, gy = spectral_grad_2d(vhat, rfft_mesh)
    return gy, -gx


def brick_wall_bounds(n: int) -> Tuple[int, int, int]:
    """(rows kept at the low end, rows kept at the high end, columns kept) of the 2/3 rule.
    The high block is one row taller when int(2n/3) is odd: the reference writes
    ``-int(2/3*n) // 2`` and unary minus binds before ``//``."""
    m = int(2 / 3 * n)
    return m // 2, -((-m) // 2), int(2 / 3 * (n // 2 + 1))


def brick_wall_filter_2d(grid: Grid):
    """2/3-rule de-aliasing mask on the rfft2 half spectrum (n, n//2+1)."""
    n, _ = grid.shape
    r_lo, r_hi, c = brick_wall_bounds(n)
    filter_ = torch.zeros((n, n // 2 + 1))
    filter_[:r_lo, :c] = 1
    filter_[n - r_hi:, :c] = 1
    return filter_


def vorticity_to_velocity(grid: Grid, w_hat: torch.Tensor,
                          rfft_mesh: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """Velocity and stream function spectra from the vorticity spectrum:
    psi = -w / lap', (u, v) = (d psi/dy, -d psi/dx)."""
    kx, ky = rfft_mesh if rfft_mesh is not None else grid.rfft_mesh()
    kx, ky = kx.to(w_hat.device), ky.to(w_hat.device)
    assert kx.shape[-2:] == w_hat.shape[-2:]
    lap = spectral_laplacian_2d((kx, ky))
    psi_hat = -1 / lap * w_hat
    u_hat, v_hat = spectral_rot_2d(psi_hat, (kx, ky))
    return (u_hat, v_hat), psi_hat
