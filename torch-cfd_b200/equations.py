"""Drop-in ``NavierStokes2DSpectral`` / ``RK4CrankNicolsonStepper`` backed by the fused sm_100a
kernels of libtcfd (reference: torch_cfd/equations.py:35-107, :249-463).

Same constructor arguments, buffer names (``kx, ky, laplace, linear_term, filter``), methods
(``forward/step, explicit_terms, implicit_terms, implicit_solve, residual``) and return values as
the reference.  Differences, all deliberate:

* the equation runs on CUDA tensors only -- there is no CPU fallback; a CPU tensor raises;
* ``forward`` with an ``RK4CrankNicolsonStepper`` is ONE call into the C ABI for all ``steps``
  (2 kernel launches per RK substage, no torch.fft, no eager elementwise kernels);
* a state-independent forcing (both shipped spectral forcings are) is evaluated once and added inside the
  kernel instead of 5x per step; a user forcing that DOES depend on the state takes
  the host-driven sub-stage loop: CUDA ``explicit_terms`` (advection) plus the forcing evaluated per
  sub-stage with the libtcfd transforms (``torch-cfd_b200/fft.py``), as upstream does (equations.py:429-437);
* the solver is inference-only (``torch.no_grad`` semantics), like every shipped caller.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from .grids import Grid
from .spectral import brick_wall_filter_2d, spectral_laplacian_2d, vorticity_to_velocity

Params = Dict[str, torch.Tensor]


def stable_time_step(dx: float = None, dt: float = None, max_velocity: float = 1.0,
                     max_courant_number: float = 0.5, viscosity: float = 1e-3,
                     implicit_diffusion: bool = True, ndim: int = 2) -> float:
    """Host scalar helper with the reference's signature and defaults (torch_cfd/equations.py:35-64):
    min(diffusive bound, advective CFL bound, dt); the diffusive bound is ``dx`` when diffusion is
    implicit and ``dx^2 / (viscosity 2^ndim)`` when it is explicit."""
    dt_diffusion = dx
    if not implicit_diffusion:
        dt_diffusion = dx**2 / (viscosity * 2 ** (ndim))
    dt_advection = max_courant_number * dx / max_velocity
    dt = dt_advection if dt is None else dt
    return min(dt_diffusion, dt_advection, dt)


class ImplicitExplicitODE(nn.Module):
    """du/dt = F(u) + G(u) with F explicit and G implicit (reference: equations.py:67-107)."""

    def explicit_terms(self, u):
        raise NotImplementedError

    def implicit_terms(self, u):
        raise NotImplementedError

    def implicit_solve(self, u, step_size):
        raise NotImplementedError

    def residual(self, u, u_t):
        raise NotImplementedError


_CK = {
    "alphas": [0, 0.1496590219993, 0.3704009573644, 0.6222557631345, 0.9582821306748, 1],
    "betas": [0, -0.4178904745, -1.192151694643, -1.697784692471, -1.514183444257],
    "gammas": [0.1496590219993, 0.3792103129999, 0.8229550293869, 0.6994504559488, 0.1530572479681],
}
_RK4 = {
    "alphas": [0.0, 0.5, 0.5, 1.0, 1.0],
    # upstream writes integer zeros here, which makes its constructor raise (nn.Parameter of an
    # int64 tensor, torch_cfd/equations.py:322,167); floats make the documented option usable
    "betas": [0.0, 0.0, 0.0, 0.0],
    "gammas": [1 / 6, 1 / 3, 1 / 3, 1 / 6],
}


class IMEXStepper(nn.Module):
    """Implicit-explicit steppers of configurable order (reference: torch_cfd/equations.py:110-246):
    order 1 / 1.5 -> ``g = u + dt F(u) + (1 - alpha) dt G(u); u = G_inv(g, alpha dt)`` (:176-193),
    order 2 -> RK2 + Crank-Nicolson (:195-229).  ``params`` holds ``alpha`` and ``beta`` like upstream.

    With ``alpha = 0.5`` the order-1 / 1.5 step is ONE sub-stage of the fused libtcfd step
    (beta_k = 0, gamma_k dt = dt, mu = dt / 2): it runs as a single launch per call.  Every other
    setting takes the generic path: the equation's CUDA ``explicit_terms`` plus torch elementwise ops
    (the equation module must then live on the state's device, as upstream)."""

    def __init__(self, order: float = 2, alpha: float = 0.5, beta: Optional[float] = 0.5,
                 requires_grad: bool = False, *args, **kwargs):
        super().__init__()
        if order not in (1, 1.5, 2):
            raise ValueError("IMEXStepper: order must be 1, 1.5 or 2")
        self.order = order
        params = {"alpha": torch.tensor(alpha), "beta": torch.tensor(beta)}
        self.params = nn.ParameterDict({k: nn.Parameter(v, requires_grad=requires_grad) for k, v in params.items()})
        self.requires_grad = requires_grad

    def fusable(self, dt: float, params: Optional[Params] = None) -> bool:
        """True when the step is one sub-stage of the fused kernel: numerator and denominator of the
        implicit part use the same factor, (1 - alpha) dt == alpha dt."""
        params = self.params if params is None else params
        a = params["alpha"]
        return self.order in (1, 1.5) and float((1 - a) * dt) == float(a * dt)

    def substage_scalars(self, dt: float, params: Optional[Params] = None):
        params = self.params if params is None else params
        return [0.0], [float(dt)], [float(params["alpha"] * dt)]

    def forward(self, u: torch.Tensor, dt: float, equation: ImplicitExplicitODE,
                params: Optional[Params] = None) -> torch.Tensor:
        if isinstance(equation, NavierStokes2DSpectral) and self.fusable(dt, params) and not equation.state_dependent_forcing:
            return equation._fused_steps(u, dt, 1, self, params, want_dudt=False)[0]
        params = self.params if params is None else params
        alpha, beta = params["alpha"], params["beta"]
        F, G, G_inv = equation.explicit_terms, equation.implicit_terms, equation.implicit_solve
        if self.order in (1, 1.5):
            g = u + dt * F(u) + (1 - alpha) * dt * G(u)
            return G_inv(g, alpha * dt)
        g = u + beta * dt * G(u)
        h = F(u)
        u = G_inv(g + dt * h, beta * dt)
        h = alpha * F(u) + (1 - alpha) * h
        return G_inv(g + dt * h, beta * dt)


class RK4CrankNicolsonStepper(IMEXStepper):
    """Low-storage Runge-Kutta (Carpenter-Kennedy 2N) for the explicit terms with Crank-Nicolson
    for the implicit ones (reference: torch_cfd/equations.py:249-358).

    ``params`` holds ``alphas`` (len s+1), ``betas`` and ``gammas`` (len s) exactly like upstream, and the
    class derives from ``IMEXStepper`` as upstream does (``isinstance(solver, IMEXStepper)`` keeps its meaning).
    For a libtcfd-backed equation the whole step is fused on the GPU; any other
    ``ImplicitExplicitODE`` takes the generic sub-stage loop.
    """

    def __init__(self, order: float = 4, requires_grad: bool = False, weights: Optional[Params] = None,
                 low_storage: bool = True, *args, **kwargs):
        nn.Module.__init__(self)  # upstream calls IMEXStepper.__init__(order) and then replaces params
        self.order = order
        table = _CK if low_storage else _RK4
        params = {k: torch.tensor(v) for k, v in table.items()}
        self.params = nn.ParameterDict({k: nn.Parameter(v, requires_grad=requires_grad) for k, v in params.items()})
        self.requires_grad = requires_grad

    def fusable(self, dt: float, params: Optional[Params] = None) -> bool:
        return True

    def substage_scalars(self, dt: float, params: Optional[Params] = None):
        """(beta_k, gamma_k*dt, mu_k) for every substage, evaluated with the same tensor
        arithmetic (and therefore the same rounding) as upstream equations.py:355-357."""
        params = self.params if params is None else params
        alphas, betas, gammas = params["alphas"], params["betas"], params["gammas"]
        beta, gdt, mu = [], [], []
        for k in range(len(betas)):
            beta.append(float(betas[k]))
            gdt.append(float(gammas[k] * dt))
            mu.append(float(0.5 * dt * (alphas[k + 1] - alphas[k])))
        return beta, gdt, mu

    def forward(self, u: torch.Tensor, dt: float, equation: ImplicitExplicitODE,
                params: Optional[Params] = None) -> torch.Tensor:
        if isinstance(equation, NavierStokes2DSpectral) and not equation.state_dependent_forcing:
            return equation._fused_steps(u, dt, 1, self, params, want_dudt=False)[0]
        params = self.params if params is None else params
        alphas, betas, gammas = params["alphas"], params["betas"], params["gammas"]
        F, G, G_inv = equation.explicit_terms, equation.implicit_terms, equation.implicit_solve
        h = 0
        for k in range(len(betas)):
            h = F(u) + betas[k] * h
            mu = 0.5 * dt * (alphas[k + 1] - alphas[k])
            u = G_inv(u + gammas[k] * dt * h + mu * G(u), mu)
        return u


def _probe_state_independent(forcing_fn, grid: Grid, vorticity: bool) -> bool:
    n = grid.shape[0]
    g = torch.Generator().manual_seed(0)

    def state():
        if vorticity:
            return torch.randn(n, n // 2 + 1, generator=g, dtype=torch.get_default_dtype()).to(torch.complex64)
        return (torch.randn(n, n, generator=g), torch.randn(n, n, generator=g))

    def flat(out):
        outs = out if isinstance(out, (tuple, list)) else (out,)
        return [getattr(o, "data", o) for o in outs]

    a, b = flat(forcing_fn(grid, state())), flat(forcing_fn(grid, state()))
    return all(torch.equal(x, y) for x, y in zip(a, b))


class NavierStokes2DSpectral(ImplicitExplicitODE):
    """2-D vorticity Navier-Stokes, pseudo-spectral, explicit advection + implicit diffusion
    (reference: torch_cfd/equations.py:361-463).

    Args:
      viscosity, grid, drag, smooth (2/3 rule), forcing_fn, solver: as upstream.
    """

    def __init__(self, viscosity: float, grid: Grid, drag: float = 0.0, smooth: bool = True,
                 forcing_fn: Optional[Callable] = None, solver: Optional[nn.Module] = None, **kwargs):
        super().__init__()
        self.viscosity = viscosity
        self.grid = grid
        self.drag = drag
        self.smooth = smooth
        self.forcing_fn = forcing_fn
        self.solver = solver
        self._plans = _lib.HandleStore()
        self._initialize()

    def _initialize(self):
        kx, ky = self.grid.rfft_mesh()
        self.register_buffer("kx", kx)
        self.register_buffer("ky", ky)
        laplace = -4 * (torch.pi) ** 2 * (abs(self.kx) ** 2 + abs(self.ky) ** 2)
        self.register_buffer("laplace", laplace)
        filter_ = brick_wall_filter_2d(self.grid)
        linear_term = self.viscosity * self.laplace - self.drag
        self.register_buffer("linear_term", linear_term)
        self.register_buffer("filter", filter_)

    # ------------------------------------------------------------------ plan management
    def forcing_hat(self) -> Optional[torch.Tensor]:
        """Spectrum of the (state-independent) forcing exactly as upstream adds it in every
        substage (equations.py:429-437), evaluated once on the host."""
        if self.forcing_fn is None:
            return None
        fn = self.forcing_fn
        vorticity = bool(getattr(fn, "vorticity", False))
        if self.state_dependent_forcing:
            return None  # added per sub-stage by explicit_terms (host-driven loop)
        kx, ky = self.kx.cpu(), self.ky.cpu()
        if vorticity:
            f = fn(self.grid, None)
            return torch.fft.rfft2(getattr(f, "data", f).cpu())
        fx, fy = fn(self.grid, None)
        fx_hat = torch.fft.rfft2(getattr(fx, "data", fx).cpu())
        fy_hat = torch.fft.rfft2(getattr(fy, "data", fy).cpu())
        return 2j * torch.pi * (fy_hat * kx - fx_hat * ky)

    @property
    def state_dependent_forcing(self) -> bool:
        """True when ``forcing_fn`` reads its state argument (probed once with two random states; the shipped
        ``KolmogorovForcing`` and anything that declares ``state_independent = True`` are known not to)."""
        cached = self.__dict__.get("_sdf")
        if cached is None:
            fn = self.forcing_fn
            if fn is None:
                cached = False
            else:
                known = getattr(fn, "state_independent", False) or type(fn).__name__ == "KolmogorovForcing"
                cached = not known and not _probe_state_independent(fn, self.grid, bool(getattr(fn, "vorticity", False)))
            self.__dict__["_sdf"] = cached
        return cached

    def _dynamic_forcing_hat(self, vort_hat: torch.Tensor) -> torch.Tensor:
        """Spectrum of a state-dependent forcing for the state ``vort_hat`` (B, n, nh), evaluated like upstream
        (equations.py:429-437) with the libtcfd transforms: the velocity form receives the physical velocity."""
        from . import fft as _fft
        from .spectral import spectral_curl_2d
        fn = self.forcing_fn
        kx, ky = self.kx.to(vort_hat.device), self.ky.to(vort_hat.device)
        data = lambda f: getattr(f, "data", f).to(vort_hat.device)
        if getattr(fn, "vorticity", False):
            return _fft.rfft2(data(fn(self.grid, vort_hat)).expand(vort_hat.shape[0], -1, -1).contiguous())
        (uhat, vhat), _ = vorticity_to_velocity(self.grid, vort_hat, (kx, ky))
        fx, fy = fn(self.grid, (_fft.irfft2(uhat), _fft.irfft2(vhat)))
        shape = (vort_hat.shape[0],) + tuple(self.kx.shape[:1]) * 2
        fx_hat = _fft.rfft2(data(fx).expand(shape).contiguous())
        fy_hat = _fft.rfft2(data(fy).expand(shape).contiguous())
        return spectral_curl_2d((fx_hat, fy_hat), (kx, ky))

    def invalidate_plan(self):
        """Drop cached device plans (call after editing a buffer or the forcing)."""
        for p in self._plans.values():
            p.close()
        self._plans = _lib.HandleStore()

    def _plan(self, device: torch.device, batch: int) -> "_lib.NS2DPlan":
        if device.type != "cuda":
            raise RuntimeError(
                "torch-cfd_b200 runs the spectral step on CUDA devices only (no CPU fallback); "
                f"got a tensor on {device}. Move the state to a B200 first.")
        key = (device.index if device.index is not None else torch.cuda.current_device())
        plan = self._plans.get(key)
        if plan is not None and plan.max_batch >= batch:
            return plan
        lib = _lib.load_library()  # (a superseded smaller plan is released when its last reference goes)
        kx, ky = self.kx.detach().cpu(), self.ky.detach().cpu()
        dtype = kx.dtype
        n = kx.shape[0]
        kappa_x = (2j * torch.pi * kx).imag[:, 0]
        kappa_y = (2j * torch.pi * ky).imag[0, :]
        neg_inv_lap = -1 / spectral_laplacian_2d((kx, ky))
        filt = self.filter.detach().cpu() if self.smooth else None
        with torch.cuda.device(key):
            plan = _lib.NS2DPlan(lib, n, dtype, batch, kappa_x, kappa_y, neg_inv_lap,
                                 self.linear_term.detach().cpu(), filt, self.forcing_hat())
        self._plans[key] = plan
        return plan

    def _as_batch(self, t: torch.Tensor, host: bool = False) -> Tuple[torch.Tensor, torch.Size]:
        n, nh = self.kx.shape
        if t.device.type != "cuda" and not host:
            raise RuntimeError(
                "torch-cfd_b200 runs the spectral step on CUDA devices only (no CPU fallback); "
                f"got a tensor on {t.device}. Move the state to a B200 first.")
        if t.dim() < 2 or tuple(t.shape[-2:]) != (n, nh):
            raise ValueError(f"expected a spectrum of shape (*, {n}, {nh}), got {tuple(t.shape)}")
        want = torch.complex64 if self.kx.dtype == torch.float32 else torch.complex128
        if t.dtype != want:
            raise TypeError(f"expected {want} (equation buffers are {self.kx.dtype}), got {t.dtype}")
        return t.detach().reshape(-1, n, nh).contiguous(), t.shape

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def explicit_terms(self, vort_hat):
        w, shape = self._as_batch(vort_hat)
        out = torch.empty_like(w)
        with torch.cuda.device(w.device):
            self._plan(w.device, w.shape[0]).explicit_terms(w, out)
            if self.state_dependent_forcing:
                out += self._dynamic_forcing_hat(w)
        return out.reshape(shape)

    def implicit_terms(self, vort_hat):
        return self.linear_term * vort_hat

    def implicit_solve(self, vort_hat, dt):
        return 1 / (1 - dt * self.linear_term) * vort_hat

    @torch.no_grad()
    def residual(self, vhat: torch.Tensor, vt_hat: torch.Tensor):
        w, shape = self._as_batch(vhat)
        wt, _ = self._as_batch(vt_hat)
        out = torch.empty_like(w)
        with torch.cuda.device(w.device):
            self._plan(w.device, w.shape[0]).residual(w, wt, out)
            if self.state_dependent_forcing:
                out -= self._dynamic_forcing_hat(w)
        return out.reshape(shape)

    def step(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    @torch.no_grad()
    def _fused_steps(self, vort_hat, dt, steps, solver, params=None, want_dudt=True):
        w, shape = self._as_batch(vort_hat)
        beta, gdt, mu = solver.substage_scalars(dt, params)
        out = torch.empty_like(w)
        dwdt = torch.empty_like(w) if want_dudt else None
        with torch.cuda.device(w.device):
            self._plan(w.device, w.shape[0]).step(w, out, dwdt, steps, beta, gdt, mu, 1 / (steps * dt))
        return out.reshape(shape), (dwdt.reshape(shape) if want_dudt else None)

    def forward(self, vort_hat, dt, steps=1) -> Tuple[torch.Tensor, torch.Tensor]:
        """vort_hat: (B, n, n//2+1), (n_t, n, n//2+1) or (n, n//2+1) complex spectrum.
        Returns (vort_hat after ``steps`` steps, (new - old) / (steps * dt))."""
        if torch.is_grad_enabled() and (vort_hat.requires_grad or (
                self.solver is not None and any(p.requires_grad for p in self.solver.parameters()))):
            raise NotImplementedError(
                "torch-cfd_b200: the CUDA step is inference-only (no autograd through the solver, SURVEY 8b); "
                "run it under torch.no_grad() or detach the state / freeze the stepper parameters")
        if isinstance(self.solver, IMEXStepper) and self.solver.fusable(dt) and not self.state_dependent_forcing:
            return self._fused_steps(vort_hat, dt, steps, self.solver)  # includes RK4CrankNicolsonStepper
        if self.solver is None:
            raise TypeError("NavierStokes2DSpectral.solver is None: pass solver=RK4CrankNicolsonStepper()")
        vort_old = vort_hat
        for _ in range(steps):
            vort_hat = self.solver(vort_hat, dt, self)
        return vort_hat, 1 / (steps * dt) * (vort_hat - vort_old)

    def check_kernels(self, device=None):
        """Raise if a kernel of an earlier (now synchronised) call on ``device`` reported a failure, e.g. the bounded
        dependency wait of the dataflow launch (tcfd_ns2d_check; bound: TCFD_FLOW_TIMEOUT_S seconds, 0 disables it)."""
        want = None
        if device is not None:
            d = torch.device(device)
            want = d.index if d.index is not None else torch.cuda.current_device()
        for key, plan in list(self._plans.items()):
            if want is None or key == want:
                plan.check()

    @torch.no_grad()
    def forward_host(self, vort_hat_host: torch.Tensor, dt, steps=1, device=None, out=None, dvdt_out=None):
        """End-to-end variant for HOST-resident states: uploads the (preferably pinned) CPU
        tensor, steps on the GPU and downloads both results (tcfd_ns2d_step_host).  Returns pinned
        CPU tensors (``out`` / ``dvdt_out`` when given -- pass pre-allocated pinned buffers to
        avoid a cudaHostAlloc per call); synchronises the stream before returning."""
        if vort_hat_host.is_cuda:
            raise ValueError("forward_host expects a CPU tensor")
        if not isinstance(self.solver, RK4CrankNicolsonStepper):
            raise TypeError("forward_host needs an RK4CrankNicolsonStepper")
        if not torch.cuda.is_available():
            raise RuntimeError("torch-cfd_b200: no CUDA device (there is no CPU fallback)")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        w, shape = self._as_batch(vort_hat_host, host=True)
        beta, gdt, mu = self.solver.substage_scalars(dt)
        if out is None:
            out = torch.empty(w.shape, dtype=w.dtype, pin_memory=True)
        if dvdt_out is None:
            dvdt_out = torch.empty(w.shape, dtype=w.dtype, pin_memory=True)
        o, d = out.view(w.shape), dvdt_out.view(w.shape)
        if o.data_ptr() == w.data_ptr():
            raise ValueError("out must not alias the input")
        with torch.cuda.device(dev):
            plan = self._plan(dev, w.shape[0])
            plan.step(w, o, d, steps, beta, gdt, mu, 1 / (steps * dt), host=True)
            torch.cuda.current_stream().synchronize()
            plan.check()  # a kernel-side failure (dependency time-out of the dataflow launch) surfaces here, not later
        return o.view(shape), d.view(shape)
