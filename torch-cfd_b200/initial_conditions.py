"""Initial-condition generators of the spectral data-generation scripts on the libtcfd transforms:

* ``vorticity_field``      McWilliams-spectrum vorticity (torch_cfd/initial_conditions.py:170-199; used by
                           fno/data_gen/data_gen_McWilliams2d.py:119-126 and by the bench configs),
* ``GRF2d.sample``         Gaussian random field (fno/data_gen/grf.py:13-115),
* ``spectral_poisson_apply``  ``irfftn(multiplier * rfftn(rhs))`` of the fast-diagonalisation pressure solve
                           (torch_cfd/pressure.py:357-364).

All three are "real FFT -> pointwise -> real inverse FFT".  The reference writes them with full complex transforms
of real data and takes ``.real``; a real field's spectrum is Hermitian and the filters depend on ``|k|`` only, so
the half-spectrum transforms of ``torch-cfd_b200/fft.py`` give the same result at half the work.  The white noise
comes from the reference's generator (CPU, seeded) so that a sample does not depend on the device; everything after
it runs on the CUDA device."""
from __future__ import annotations

import math
from typing import Callable, Optional

import torch

from . import fft as _fft
from .forcings import Field
from .grids import Grid


def McWilliams_density(k, mode: float, tau: float = 1.0):
    """|psi|^2 ~ k^-1 (tau^2 + (k / k0)^4)^-1 (McWilliams 1984; torch_cfd/initial_conditions.py:68-78)."""
    return (k * (tau ** 2 + (k / mode) ** 4)) ** (-1)


def _log_normal_density(k, mode: float, variance=0.25):
    mean = math.log(mode) + variance
    logk = torch.log(k)
    return torch.exp(-((mean - logk) ** 2) / 2 / variance - logk)


def _angular_frequency_magnitude(grid: Grid, half: bool = False, device=None) -> torch.Tensor:
    """|2 pi k| on the full (n, n) mesh, or on the rfft2 half mesh (n, n//2+1) (initial_conditions.py:81-87)."""
    freqs = [2 * torch.pi * torch.fft.fftfreq(size, step) for size, step in zip(grid.shape, grid.step)]
    kx, ky = torch.meshgrid(*freqs, indexing="ij")
    k = torch.linalg.norm(torch.stack([kx, ky]), dim=0)
    if half:
        k = k[:, : grid.shape[1] // 2 + 1]
    return k.to(device) if device is not None else k


def _half_weights(n: int, device) -> torch.Tensor:
    """Multiplicity of every rfft2 column in a sum over the full spectrum (1 for ky = 0 and n/2, else 2)."""
    w = torch.full((n // 2 + 1,), 2.0, device=device)
    w[0] = 1.0
    if n % 2 == 0:
        w[-1] = 1.0
    return w


def spectral_filter(spectral_density: Callable, v: torch.Tensor, grid: Grid) -> torch.Tensor:
    """White noise -> field with the prescribed spectral density (initial_conditions.py:90-101), CUDA tensors."""
    k = _angular_frequency_magnitude(grid, half=True, device=v.device)
    filters = torch.where(k > 0, spectral_density(k), torch.zeros_like(k)).to(v.dtype)
    return _fft.irfft2(_fft.rfft2(v) * filters)


def streamfunc_normalize(k_half: torch.Tensor, psi_hat: torch.Tensor, n: int):
    """psi / sqrt(kinetic energy), kinetic energy = sum over the FULL spectrum of 2 |k psi^|^2 / n^4
    (initial_conditions.py:104-109), evaluated on the half spectrum."""
    e = (2 * (k_half * psi_hat).abs() ** 2 / (n * n) ** 2 * _half_weights(n, psi_hat.device)).sum(dim=(-2, -1), keepdim=True)
    return psi_hat / e.sqrt()


def vorticity_field(grid: Grid, peak_wavenumber: float = 3, random_state: int = 0, device=None, spectrum: bool = False):
    """McWilliams vorticity field (torch_cfd/initial_conditions.py:170-199).  ``device``: CUDA device of the result
    (default: the current one); ``spectrum=True`` returns ``rfft2`` of the field instead -- what every spectral
    caller computes next (fno/data_gen/data_gen_McWilliams2d.py:126) -- and saves the round trip."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    n = grid.shape[0]
    rng = torch.Generator()
    rng.manual_seed(random_state)
    noise = torch.randn(grid.shape, generator=rng).to(device)
    k = _angular_frequency_magnitude(grid, half=True, device=device).to(noise.dtype)
    filters = torch.where(k > 0, McWilliams_density(k, peak_wavenumber), torch.zeros_like(k))
    psi_hat = streamfunc_normalize(k, _fft.rfft2(noise) * filters, n)
    w_hat = psi_hat * k ** 2
    if spectrum:
        return w_hat
    return Field(_fft.irfft2(w_hat), grid.cell_faces, grid)


class GRF2d:
    """Gaussian random field on [0, 1]^2 with covariance (-Delta + tau^2)^-alpha (fno/data_gen/grf.py:13-115): same
    constructor arguments and ``sample(bsz, n, random_state)``; the random coefficients come from torch's generator
    for ``device`` exactly as upstream, the inverse transform ``ifftn(coeff).real`` runs as ``irfft2`` of the
    Hermitian part of ``coeff``."""

    def __init__(self, *, dim=2, n=128, alpha=2, tau=3, device="cuda", dtype=torch.float, normalize=False,
                 smoothing=False, **kwargs):
        assert dim == 2
        self.dim, self.n, self.device, self.dtype = dim, n, device, dtype
        self.normalize, self.alpha, self.tau, self.smoothing = normalize, alpha, tau, smoothing
        self.max_mesh_size = 2048
        self._initialize()

    def _initialize(self, n=None, device=None, alpha=None, tau=None, sigma=None):
        n = self.n if n is None else n
        device = self.device if device is None else device
        alpha = self.alpha if alpha is None else alpha
        tau = self.tau if tau is None else tau
        sigma = tau ** (0.5 * (2 * alpha - self.dim)) if sigma is None else sigma
        k = torch.fft.fftfreq(n, d=1 / n, device=device)
        kx, ky = torch.meshgrid(k, k, indexing="ij")
        sqrt_eig = (n ** self.dim) * math.sqrt(2.0) * sigma * ((4 * (math.pi ** 2) * (kx ** 2 + ky ** 2) + tau ** 2) ** (-alpha / 2.0))
        sqrt_eig[0, 0] = 0.0
        self.sqrt_eig = sqrt_eig
        self.n = n

    def sample(self, bsz, n=None, random_state=0, **kwargs):
        import torch.nn.functional as F
        if n is not None and n != self.n:
            self._initialize(n=n, **kwargs)
        n = self.n
        torch.cuda.manual_seed(random_state)
        torch.random.manual_seed(random_state)
        if self.smoothing:
            coeff = torch.randn(bsz, 2, self.max_mesh_size, self.max_mesh_size, dtype=self.dtype, device=self.device)
            coeff = F.interpolate(coeff, size=[n, n], mode="bilinear")
        else:
            coeff = torch.randn(bsz, 2, n, n, dtype=self.dtype, device=self.device)
        coeff = self.sqrt_eig * (coeff[:, 0] + 1j * coeff[:, 1])
        # Re(ifft2(c)) = irfft2 of the Hermitian part h(k) = (c(k) + conj(c(-k))) / 2 on the half spectrum
        neg = torch.roll(torch.flip(coeff, dims=(-2, -1)), shifts=(1, 1), dims=(-2, -1))
        h = (0.5 * (coeff + neg.conj()))[..., : n // 2 + 1].contiguous()
        s = _fft.irfft2(h)
        if self.normalize:
            s = s / torch.linalg.norm(s / n, dim=(-1, -2), keepdim=True)
        return s


def spectral_poisson_apply(value: torch.Tensor, multiplier: torch.Tensor) -> torch.Tensor:
    """``ifft(multiplier * fft(value), s=grid.shape).real`` of the fast-diagonalisation pressure solve on a periodic
    grid (torch_cfd/pressure.py:357-364 with ``fft = rfftn``): ``multiplier`` on the rfft2 half mesh."""
    return _fft.irfft2(_fft.rfft2(value) * multiplier.to(value.device))
