"""CPU oracle for hot path A: the pseudo-spectral 2-D vorticity Navier-Stokes RK4+CN step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product package
(``torch-cfd_b200/``); only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may use it, and only as the checker / the CPU arm.

What it is: a functional restatement, on CPU tensors, of the algorithm in the reference
(scaomath/torch-cfd @ 475c738).  Each function cites the reference lines it follows.

Where the arithmetic really lives: the 2-D real FFTs are third-party -- ``torch.fft.rfft2`` /
``torch.fft.irfft2`` of the pinned dependency ``torch>=2.5`` (requirements.txt:2; installed 2.11.0,
CPU backend MKL/pocketfft).  The oracle calls that same library (it is what the reference runs on
CPU), and ``oracle/dft_naive.py`` restates the published DFT definition it implements as O(n^2)
numpy matrices so the FFT convention itself (sign, normalisation, half-spectrum, C2R treatment of
the self-conjugate columns) is pinned independently at small n (tests/test_oracle_cpu.py).

Parity pin: the reference ships no golden vectors for this path, so the
oracle is pinned against outputs of the reference itself, generated in the build container by
``tests/golden/make_golden.py`` (imports /root/reference) and committed as ``tests/golden/*.npz``.
``tests/test_oracle_cpu.py`` requires bit-for-bit agreement with those outputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Optional, Sequence, Tuple

import torch

# Carpenter-Kennedy 2N-storage coefficients: torch_cfd/equations.py:294-317
CK_ALPHAS = [0, 0.1496590219993, 0.3704009573644, 0.6222557631345, 0.9582821306748, 1]
CK_BETAS = [0, -0.4178904745, -1.192151694643, -1.697784692471, -1.514183444257]
CK_GAMMAS = [0.1496590219993, 0.3792103129999, 0.8229550293869, 0.6994504559488, 0.1530572479681]
# classic RK4 written in the same (alpha, beta, gamma) form: torch_cfd/equations.py:320-324
RK4_ALPHAS = [0, 0.5, 0.5, 1.0, 1.0]
RK4_BETAS = [0, 0, 0, 0]
RK4_GAMMAS = [1 / 6, 1 / 3, 1 / 3, 1 / 6]


def rfft_mesh(n: int, diam: float, dtype: torch.dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    """Ordinal-frequency half mesh (kx, ky), each (n, n//2+1).

    torch_cfd/grids.py:159-170 (fft_axes: fftfreq(n, d=step)), :191-201 (meshgrid 'ij', keep the
    first n//2+1 columns).  step = diam / n as in grids.py:103-106.
    """
    step = (diam - 0.0) / n
    ax = torch.fft.fftfreq(n, d=step, dtype=dtype)
    kx, ky = torch.meshgrid(ax, ax, indexing="ij")
    nh = math.floor(n / 2.0) + 1
    return kx[..., :nh], ky[..., :nh]


def laplacian(kx: torch.Tensor, ky: torch.Tensor) -> torch.Tensor:
    """-4 pi^2 (|kx|^2 + |ky|^2): torch_cfd/equations.py:398."""
    return -4 * (torch.pi) ** 2 * (abs(kx) ** 2 + abs(ky) ** 2)


def laplacian_patched(kx: torch.Tensor, ky: torch.Tensor) -> torch.Tensor:
    """Same with the [0,0] entry set to 1 (used only for the stream function):
    torch_cfd/spectral.py:41-46."""
    lap = -4 * (torch.pi**2) * (abs(kx) ** 2 + abs(ky) ** 2)
    lap[..., 0, 0] = 1
    return lap


def brick_wall_bounds(n: int) -> Tuple[int, int, int]:
    """Integer bounds of the 2/3-rule mask: rows [0, r_lo) and the last r_hi rows, columns [0, c).

    torch_cfd/spectral.py:78-84.  Note the reference writes ``-int(2/3*n) // 2``: unary minus binds
    before ``//`` so the upper block has ceil(int(2n/3)/2) rows.
    """
    r_lo = int(2 / 3 * n) // 2
    r_hi = -(-int(2 / 3 * n) // 2)
    c = int(2 / 3 * (n // 2 + 1))
    return r_lo, r_hi, c


def brick_wall_mask(n: int, dtype: torch.dtype) -> torch.Tensor:
    """0/1 mask (n, n//2+1): torch_cfd/spectral.py:78-84."""
    r_lo, r_hi, c = brick_wall_bounds(n)
    m = torch.zeros((n, n // 2 + 1), dtype=dtype)
    m[:r_lo, :c] = 1
    m[n - r_hi :, :c] = 1
    return m


def cell_center_mesh(n: int, diam: float, dtype: torch.dtype, offset=(0.5, 0.5)):
    """Physical mesh x, y (n, n) at lower + (i + offset) * step: torch_cfd/grids.py:137-157,172-189."""
    step = (diam - 0.0) / n
    axes = [0.0 + (torch.arange(n).to(dtype) + o) * step for o in offset]
    x, y = torch.meshgrid(*axes, indexing="ij")
    return x, y


def kolmogorov_forcing_vorticity(n, diam, dtype, scale=1.0, wave_number=1, forcing_diam=2 * torch.pi,
                                 offset=(0, 0)):
    """Vorticity-form Kolmogorov forcing field f(x, y) = -a k c cos(k c y), c = 2 pi / forcing_diam.

    torch_cfd/forcings.py:181-210 (swap_xy=False branch), offsets default ((0,0),(0,0)) from
    forcings.py:139-153.  Returns the real (n, n) field; the solver adds rfft2 of it.
    """
    _, y = cell_center_mesh(n, diam, dtype, offset)
    c = 2 * torch.pi / forcing_diam
    return -scale * wave_number * c * torch.cos(wave_number * c * y)


def kolmogorov_forcing_velocity(n, diam, dtype, scale=1.0, wave_number=1, forcing_diam=2 * torch.pi,
                                offset=(0, 0)):
    """Velocity-form Kolmogorov forcing (fx, fy) = (a sin(k c y), 0): torch_cfd/forcings.py:158-179."""
    _, y = cell_center_mesh(n, diam, dtype, offset)
    c = 2 * torch.pi / forcing_diam
    fx = scale * torch.sin(wave_number * c * y)
    return fx, torch.zeros_like(fx)


@dataclass
class NS2DTables:
    """Everything NavierStokes2DSpectral._initialize registers (torch_cfd/equations.py:394-403)."""

    n: int
    kx: torch.Tensor
    ky: torch.Tensor
    laplace: torch.Tensor
    linear_term: torch.Tensor
    filter: torch.Tensor
    smooth: bool = True
    # time-independent forcing, one of: None, ("vorticity", f (n,n) real),
    # ("velocity", (fx, fy) real fields)
    forcing: Optional[tuple] = None


def make_tables(n: int, diam: float, viscosity: float, drag: float = 0.0, smooth: bool = True,
                forcing: Optional[tuple] = None, dtype: torch.dtype = torch.float32) -> NS2DTables:
    """torch_cfd/equations.py:394-403."""
    kx, ky = rfft_mesh(n, diam, dtype)
    lap = laplacian(kx, ky)
    linear_term = viscosity * lap - drag
    filt = brick_wall_mask(n, dtype)
    return NS2DTables(n, kx, ky, lap, linear_term, filt, smooth, forcing)


def vorticity_to_velocity(tb: NS2DTables, w_hat: torch.Tensor):
    """psi_hat = -w_hat / lap', (u_hat, v_hat) = (2 pi i ky psi_hat, -2 pi i kx psi_hat).

    torch_cfd/spectral.py:87-115, :68-75.
    """
    lap = laplacian_patched(tb.kx, tb.ky)
    psi_hat = -1 / lap * w_hat
    gx = 2j * torch.pi * tb.kx * psi_hat
    gy = 2j * torch.pi * tb.ky * psi_hat
    return (gy, -gx), psi_hat


def forcing_hat(tb: NS2DTables) -> Optional[torch.Tensor]:
    """Spectrum of the forcing as the solver adds it: torch_cfd/equations.py:429-437,
    curl form torch_cfd/spectral.py:49-56."""
    if tb.forcing is None:
        return None
    kind, f = tb.forcing
    if kind == "vorticity":
        return torch.fft.rfft2(f)
    fx_hat, fy_hat = torch.fft.rfft2(f[0]), torch.fft.rfft2(f[1])
    return 2j * torch.pi * (fy_hat * tb.kx - fx_hat * tb.ky)


def explicit_terms(tb: NS2DTables, w_hat: torch.Tensor) -> torch.Tensor:
    """F(w_hat) = filter * rfft2(-(dx w * u + dy w * v)) + f_hat: torch_cfd/equations.py:413-438."""
    (u_hat, v_hat), _ = vorticity_to_velocity(tb, w_hat)
    vx, vy = torch.fft.irfft2(u_hat), torch.fft.irfft2(v_hat)
    gx_hat = 2j * torch.pi * tb.kx * w_hat
    gy_hat = 2j * torch.pi * tb.ky * w_hat
    gx, gy = torch.fft.irfft2(gx_hat), torch.fft.irfft2(gy_hat)
    adv = -(gx * vx + gy * vy)
    adv_hat = torch.fft.rfft2(adv)
    if tb.smooth:
        adv_hat *= tb.filter
    terms = adv_hat
    fh = forcing_hat(tb)
    if fh is not None:
        terms += fh
    return terms


def implicit_terms(tb: NS2DTables, w_hat):
    """L w_hat: torch_cfd/equations.py:443-444."""
    return tb.linear_term * w_hat


def implicit_solve(tb: NS2DTables, w_hat, mu):
    """(1 / (1 - mu L)) w_hat: torch_cfd/equations.py:446-447."""
    return 1 / (1 - mu * tb.linear_term) * w_hat


def rk_coefficients(low_storage: bool = True, dtype: torch.dtype = torch.float32):
    """Coefficient tensors as the stepper stores them (default dtype at construction time):
    torch_cfd/equations.py:325-326."""
    if low_storage:
        a, b, g = CK_ALPHAS, CK_BETAS, CK_GAMMAS
    else:
        a, b, g = RK4_ALPHAS, RK4_BETAS, RK4_GAMMAS
    return (torch.tensor(a, dtype=dtype), torch.tensor(b, dtype=dtype), torch.tensor(g, dtype=dtype))


def rk4cn_step(tb: NS2DTables, u: torch.Tensor, dt: float, coeffs) -> torch.Tensor:
    """One low-storage RK (explicit) + Crank-Nicolson (implicit) step:
    torch_cfd/equations.py:328-358."""
    alphas, betas, gammas = coeffs
    h = 0
    for k in range(len(betas)):
        h = explicit_terms(tb, u) + betas[k] * h
        mu = 0.5 * dt * (alphas[k + 1] - alphas[k])
        u = implicit_solve(tb, u + gammas[k] * dt * h + mu * implicit_terms(tb, u), mu)
    return u


def forward(tb: NS2DTables, w_hat: torch.Tensor, dt: float, steps: int = 1, coeffs=None):
    """NavierStokes2DSpectral.forward: torch_cfd/equations.py:452-463."""
    if coeffs is None:
        coeffs = rk_coefficients(True, tb.kx.dtype)
    w_old = w_hat
    for _ in range(steps):
        w_hat = rk4cn_step(tb, w_hat, dt, coeffs)
    dwdt = 1 / (steps * dt) * (w_hat - w_old)
    return w_hat, dwdt


def imex_step(tb: NS2DTables, u: torch.Tensor, dt: float, order: float = 2, alpha=0.5, beta=0.5) -> torch.Tensor:
    """One IMEXStepper step (reference: torch_cfd/equations.py:176-229): order 1 / 1.5 -> ``_imex``
    (:176-193), order 2 -> ``_rk2_crank_nicolson`` (:195-229).  alpha, beta are 0-d tensors of the default
    dtype upstream (``params``), so they are taken as such here to round identically."""
    alpha = torch.as_tensor(alpha, dtype=tb.linear_term.dtype)
    beta = torch.as_tensor(beta, dtype=tb.linear_term.dtype)
    F = lambda v: explicit_terms(tb, v)
    G = lambda v: implicit_terms(tb, v)
    if order in (1, 1.5):
        g = u + dt * F(u) + (1 - alpha) * dt * G(u)
        return implicit_solve(tb, g, alpha * dt)
    g = u + beta * dt * G(u)
    h = F(u)
    u = implicit_solve(tb, g + dt * h, beta * dt)
    h = alpha * F(u) + (1 - alpha) * h
    return implicit_solve(tb, g + dt * h, beta * dt)


def imex_forward(tb: NS2DTables, w_hat: torch.Tensor, dt: float, steps: int = 1, order: float = 2, alpha=0.5,
                 beta=0.5):
    """NavierStokes2DSpectral.forward with an IMEXStepper (equations.py:449-463)."""
    w = w_hat
    for _ in range(steps):
        w = imex_step(tb, w, dt, order, alpha, beta)
    return w, 1 / (steps * dt) * (w - w_hat)


def residual(tb: NS2DTables, w_hat, wt_hat):
    """w_t - F(w) - L w: torch_cfd/equations.py:405-411."""
    return wt_hat - explicit_terms(tb, w_hat) - implicit_terms(tb, w_hat)


def trajectory(tb: NS2DTables, w0: torch.Tensor, dt: float, num_steps: int = 1,
               record_every_steps: int = 1, dtype: torch.dtype = torch.complex64, coeffs=None):
    """get_trajectory_imex without the progress bar: fno/data_gen/solvers.py:191-265.

    Records (w, psi, dw/dt, residual) after every step t with t % record_every_steps == 0,
    cast to ``dtype``, stacked on dim -3.
    """
    w = w0
    rec = {"vorticity": [], "stream": [], "vort_t": [], "residual": []}
    for t in range(num_steps):
        w, dwdt = forward(tb, w, dt, 1, coeffs)
        if t % record_every_steps == 0:
            _, psi = vorticity_to_velocity(tb, w)
            res = residual(tb, w, dwdt)
            for key, val in zip(["vorticity", "stream", "vort_t", "residual"], [w, psi, dwdt, res]):
                rec[key].append(val.detach().to(dtype).clone())
    return {k: torch.stack(v, dim=-3) for k, v in rec.items()}


def synthetic_vorticity_hat(n: int, batch: int, seed: int, dtype: torch.dtype,
                            peak_wavenumber: float = 4.0, amplitude: float = 4.0,
                            first_index: int = 0) -> torch.Tensor:
    """Seeded synthetic initial condition used by the GPU-box tests and the bench (the reference's
    McWilliams generator, torch_cfd/initial_conditions.py:170-199, does not travel to the GPU box).

    Sample i is white noise from a CPU generator seeded ``seed + first_index + i`` (so the field of
    a sample does not depend on how the batch is sharded), shaped in spectral space by a smooth
    envelope k / (1 + (k/kp)^4) peaked near ``peak_wavenumber`` and scaled so that max |w| is
    ``amplitude`` -- a decaying-turbulence-like field with the same dynamic range as config C2.
    Not claimed to equal the reference generator; parity runs feed identical tensors to both sides.
    """
    ax = torch.fft.fftfreq(n, d=1.0 / n, dtype=torch.float64)
    kx, ky = torch.meshgrid(ax, ax[: n // 2 + 1], indexing="ij")
    k = torch.sqrt(kx**2 + ky**2)
    env = k / (1 + (k / peak_wavenumber) ** 4)
    out = []
    for i in range(batch):
        g = torch.Generator()
        g.manual_seed(seed + first_index + i)
        noise = torch.randn((n, n), generator=g, dtype=torch.float64)
        wh = torch.fft.rfft2(noise) * env
        w = torch.fft.irfft2(wh, s=(n, n))
        w = w * (amplitude / w.abs().max())
        out.append(torch.fft.rfft2(w))
    cdtype = torch.complex64 if dtype == torch.float32 else torch.complex128
    return torch.stack(out).to(cdtype)
