"""``get_trajectory_imex`` (reference: fno/data_gen/solvers.py:191-265) on the fused CUDA step, plus the
batch-sharded multi-GPU form with the path's single collective (an all-gather of the recorded fields).

Reference semantics kept: the state is stepped ``num_steps`` times; after every step ``t`` with
``t % record_every_steps == 0`` the tuple (w, psi, dw/dt of that step, residual(w, dw/dt)) is recorded,
cast to ``dtype`` (complex64 by default even for fp64 runs) and stacked on dim -3; the result is a
dict ``vorticity, stream, vort_t, residual`` of CPU tensors.  Differences: steps after the last
recorded one are not executed (their result is never returned upstream either), the snapshots are
written by one kernel into device buffers and leave the GPU in ONE copy at the end instead of one
synchronising ``.cpu()`` per field per snapshot, and the upstream NameError (``tqdm`` is never
imported there) is not reproduced.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import fft as _fft
from .equations import IMEXStepper, NavierStokes2DSpectral
from .spectral import vorticity_to_velocity

FIELDS = ("vorticity", "stream", "vort_t", "residual")


_RING = {}  # device index -> (two pinned byte buffers, two events)
_RING_BYTES = 256 << 20


def _to_host(v: torch.Tensor) -> torch.Tensor:
    """Device -> host copy of a (large) result into ordinary pageable memory, like ``v.cpu()``, but staged
    through two pinned 256 MB buffers: the DMA of chunk i+1 runs while the CPU moves chunk i into place.
    Measured on the B200 box for 3.4 GB of snapshots: 0.20 s against 1.57 s for ``.cpu()`` (and no
    result-sized pinned allocation, which alone costs 1.7 s)."""
    if not v.is_cuda or v.numel() * v.element_size() < (8 << 20):
        return v.cpu()
    src = v.contiguous()
    out = torch.empty(src.shape, dtype=src.dtype)
    real_s = torch.view_as_real(src) if src.is_complex() else src
    real_o = torch.view_as_real(out) if out.is_complex() else out
    bs, bo = real_s.reshape(-1).view(torch.uint8), real_o.reshape(-1).view(torch.uint8)
    key = src.device.index
    if key not in _RING:
        _RING[key] = ([torch.empty(_RING_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(2)],
                      [torch.cuda.Event() for _ in range(2)])
    ring, ev = _RING[key]
    n = bs.numel()
    nchunks = (n + _RING_BYTES - 1) // _RING_BYTES
    with torch.cuda.device(src.device):
        for i in range(nchunks + 1):
            if i < nchunks:
                a, b = i * _RING_BYTES, min(n, (i + 1) * _RING_BYTES)
                ring[i % 2][: b - a].copy_(bs[a:b], non_blocking=True)
                ev[i % 2].record()
            if i > 0:
                j = i - 1
                a, b = j * _RING_BYTES, min(n, (j + 1) * _RING_BYTES)
                ev[j % 2].synchronize()
                bo[a:b].copy_(ring[j % 2][: b - a])
    return out


def _record_steps(num_steps: int, record_every_steps: int):
    return list(range(0, num_steps, record_every_steps))


def _trajectory_device(equation, w0, dt, num_steps, record_every_steps, dtype, fields, pbar, pbar_desc):
    """Snapshots on the device of ``w0``: dict field -> (B, n_t, n, nh) tensor of ``dtype``."""
    rec = _record_steps(num_steps, record_every_steps)
    n_t = len(rec)
    # every stepper the fused launch serves: RK4CrankNicolsonStepper (an IMEXStepper, as upstream) and the
    # order-1 / 1.5 IMEXStepper with alpha = 0.5; other IMEX settings take the generic loop below
    fused = isinstance(equation, NavierStokes2DSpectral) and isinstance(equation.solver, IMEXStepper) \
        and equation.solver.fusable(dt) and w0.is_cuda and not equation.state_dependent_forcing
    n, nh = w0.shape[-2:]
    w = w0.detach().reshape(-1, n, nh).contiguous()
    B = w.shape[0]
    snaps = {k: (torch.empty(B, n_t, n, nh, dtype=dtype, device=w.device) if k in fields else None) for k in FIELDS}
    bar = None
    if pbar:
        try:
            from tqdm import tqdm
            bar = tqdm(total=num_steps, desc=pbar_desc)
        except ImportError:
            bar = None
    done = -1
    for it, t_step in enumerate(rec):
        gap = t_step - done - 1          # un-recorded steps before this one
        if fused:
            if gap > 0:
                w, _ = equation._fused_steps(w, dt, gap, equation.solver, want_dudt=False)
            w, dwdt = equation._fused_steps(w, dt, 1, equation.solver)
            res = equation.residual(w, dwdt) if snaps["residual"] is not None else None
            plan = equation._plan(w.device, B)
            with torch.cuda.device(w.device):
                plan.record(w, dwdt, res, snaps, it)
        else:  # any ImplicitExplicitODE-like object (used by the CPU tests with an oracle-backed equation)
            for _ in range(gap):
                w, _ = equation.forward(w, dt=dt)
            w, dwdt = equation.forward(w, dt=dt)
            vals = {"vorticity": w, "vort_t": dwdt}
            if snaps["stream"] is not None:
                if hasattr(equation, "stream_function"):
                    vals["stream"] = equation.stream_function(w)
                else:  # what upstream does for any equation (fno/data_gen/solvers.py:230-231)
                    vals["stream"] = vorticity_to_velocity(equation.grid, w, (equation.kx, equation.ky))[1]
            if snaps["residual"] is not None:
                vals["residual"] = equation.residual(w, dwdt)
            for k, v in vals.items():
                if snaps[k] is not None:
                    snaps[k][:, it] = v.detach().to(dtype)
        done = t_step
        if bar is not None:
            bar.update(gap + 1)
    if bar is not None:
        bar.close()
    return {k: v for k, v in snaps.items() if v is not None}


def postprocess_trajectory(result: Dict[str, torch.Tensor], subsample: int = 1, dtype: torch.dtype = torch.float32,
                           device_result: bool = False) -> Dict[str, torch.Tensor]:
    """The post-processing loop of the reference's data-generation scripts
    (fno/data_gen/data_gen_Kolmogorov2d.py:178-188, data_gen_McWilliams2d.py:154-165) ON THE DEVICE:

        value = fft.irfft2(value).real.cpu().to(dtype)
        if subsample > 1: value = F.interpolate(value, size=(n // subsample,) * 2, mode="bilinear")

    ``result``: dict of (*, n_t, n, n//2+1) complex CUDA tensors (``get_trajectory_imex(..., device_result=True)``).
    Returns the physical-space fields (*, n_t, ns, ns) of ``dtype`` -- on the host like upstream (the copy moves
    the subsampled tensors: subsample^2 fewer bytes than the spectra), or on the device with ``device_result``."""
    out = {}
    for k, v in result.items():
        if not torch.is_tensor(v) or not v.is_complex():
            out[k] = v
            continue
        n = v.shape[-2]
        f = _fft.irfft2(v)
        if subsample > 1:
            f = _fft.interpolate_bilinear(f, n // subsample, dtype)
        elif f.dtype != dtype:
            f = f.to(dtype)
        out[k] = f if device_result else _to_host(f)
    return out


def get_trajectory_imex(equation, w0: torch.Tensor, dt: float, num_steps: int = 1, record_every_steps: int = 1,
                        pbar: bool = False, pbar_desc: str = "generating trajectories using RK4",
                        require_grad: bool = False, dtype: torch.dtype = torch.complex64,
                        fields: Sequence[str] = FIELDS, device_result: bool = False) -> Dict[str, torch.Tensor]:
    """Same signature and return value as the reference (plus ``fields`` / ``device_result``)."""
    if require_grad:
        raise NotImplementedError("torch-cfd_b200: the fused step is inference-only (SURVEY 8b)")
    lead = w0.shape[:-2]
    n, nh = w0.shape[-2:]
    with torch.no_grad():
        snaps = _trajectory_device(equation, w0, dt, num_steps, record_every_steps, dtype, tuple(fields), pbar, pbar_desc)
    out = {}
    for k, v in snaps.items():
        v = v.reshape(*lead, v.shape[1], n, nh)
        out[k] = v if device_result else _to_host(v)
    if not device_result and w0.is_cuda and hasattr(equation, "check_kernels"):
        equation.check_kernels(w0.device)  # the copies above synchronised: surface a kernel-side failure here
    return out


def get_trajectory_imex_sharded(equation, w0_local: torch.Tensor, dt: float, num_steps: int = 1,
                                record_every_steps: int = 1, dtype: torch.dtype = torch.complex64,
                                fields: Sequence[str] = ("vorticity",), group=None,
                                device_result: bool = False, gather_to="all", physical: bool = False,
                                subsample: int = 1, out_dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    """Multi-GPU trajectory: one process per GPU, each stepping ITS contiguous slice ``w0_local`` of
    the global batch (samples never interact: no halo, no data-path collective), followed by the
    path's only collective on the recorded fields (NCCL over NVLink on GPUs, gloo in the CPU tests).

    ``gather_to``: ``"all"`` (default) -- ``all_gather_into_tensor``: every rank returns the global
    (B_total, n_t, n, nh) tensors, ranks ordered along the batch axis; an ``int`` r -- the fields are gathered on
    rank r only (the other ranks return an empty dict): one copy of the result crosses NVLink and one rank
    copies it to its host; ``None`` -- no collective, every rank returns its own shard (what a data-generation
    job that writes one file per rank wants).  All ranks must hold the same local batch size.

    ``physical=True`` applies ``postprocess_trajectory`` (irfft2 + bilinear subsample + cast, the reference
    scripts' post-processing loop) on every rank's own shard BEFORE the collective and the device->host copy,
    which then move ``subsample^2`` fewer bytes."""
    import torch.distributed as dist
    with torch.no_grad():
        snaps = _trajectory_device(equation, w0_local, dt, num_steps, record_every_steps, dtype, tuple(fields), False, "")
        if physical:
            snaps = postprocess_trajectory(snaps, subsample, out_dtype, device_result=True)
    single = not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1
    if single or gather_to is None:
        return {k: (v if device_result else _to_host(v)) for k, v in snaps.items()}
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    out = {}
    for k, v in snaps.items():
        real = torch.view_as_real(v.contiguous()) if v.is_complex() else v.contiguous()
        if gather_to == "all":
            gathered = torch.empty((world * real.shape[0],) + tuple(real.shape[1:]), dtype=real.dtype, device=real.device)
            dist.all_gather_into_tensor(gathered, real, group=group)
        else:
            dst = int(gather_to)
            gathered = None
            if rank == dst:
                gathered = torch.empty((world * real.shape[0],) + tuple(real.shape[1:]), dtype=real.dtype, device=real.device)
            dist.gather(real, list(gathered.chunk(world)) if rank == dst else None, dst=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
            if rank != dst:
                continue
        g = torch.view_as_complex(gathered) if v.is_complex() else gathered
        out[k] = g if device_result else _to_host(g)
    return out


# ------------------------------------------------------------------------------------------------
# Legacy first-order IMEX / Crank-Nicolson path (reference: fno/data_gen/solvers.py:49-188, :268-448), SURVEY 8f
# rank 3.  One step is ONE sub-stage of the fused kernel (beta = 0, gamma dt = dt, mu = dt / 2) run on a plan whose
# tables are the legacy ones: linear term nu * lap' (with the patched lap'[0, 0] = 1 the reference also uses
# there), the |kx|, |ky| <= 2/3 k_max mask, the caller's forcing spectrum.
_LEGACY_PLANS = {}


def _legacy_tables(n, diam, dtype, device, scale_kmax=True):
    """(kx, ky, lap', mask) as fno/data_gen/solvers.py:316-345 builds them (scale_kmax: the trajectory driver
    compares |k| with 2/3 k_max / diam, the stand-alone step -- :150-154 -- with 2/3 k_max)."""
    import math
    k_max = math.floor(n / 2.0)
    k = torch.fft.fftfreq(n, d=diam / n, dtype=dtype, device=device)
    kx, ky = torch.meshgrid([k, k], indexing="ij")
    kx, ky = kx[..., : k_max + 1], ky[..., : k_max + 1]
    lap = -4 * (math.pi ** 2) * (kx ** 2 + ky ** 2)
    lap[0, 0] = 1.0
    km = (1 / diam) * k_max if scale_kmax else k_max
    filt = torch.logical_and(torch.abs(kx) <= (2.0 / 3.0) * km, torch.abs(ky) <= (2.0 / 3.0) * km).to(dtype)
    return kx, ky, lap, filt


def _legacy_plan(w, f_hat, visc, rfftmesh, laplacian, dealias_filter, dealias):
    """NS2DPlan configured with the legacy tables; cached per (device, n, dtype, visc, identity of the tables).
    The forcing is batch-shared inside the kernel: a plan serves ONE forcing spectrum at a time (re-uploaded when
    the caller passes another tensor)."""
    from . import _lib
    import math
    n = w.shape[-2]
    real = torch.float32 if w.dtype == torch.complex64 else torch.float64
    kx, ky = rfftmesh
    kx2, ky2 = kx.reshape(-1, *kx.shape[-2:])[0], ky.reshape(-1, *ky.shape[-2:])[0]
    lap2 = laplacian.reshape(-1, *laplacian.shape[-2:])[0]
    filt2 = None
    if dealias and dealias_filter is not None and torch.is_tensor(dealias_filter):
        filt2 = dealias_filter.reshape(-1, *dealias_filter.shape[-2:])[0]
    tag = lambda t: None if t is None else (t.data_ptr(), t._version, tuple(t.shape))
    dev = w.device.index if w.device.index is not None else torch.cuda.current_device()
    key = (dev, n, real, float(visc), tag(kx2), tag(ky2), tag(lap2), tag(filt2))
    hit = _LEGACY_PLANS.get(key)
    if hit is None or hit[0].max_batch < w.shape[0]:
        kappa_x = (2 * math.pi * kx2.to(real)).cpu()[:, 0]
        kappa_y = (2 * math.pi * ky2.to(real)).cpu()[0, :]
        lap_c = lap2.to(real).cpu()
        with torch.cuda.device(dev):
            plan = _lib.NS2DPlan(_lib.load_library(), n, real, w.shape[0], kappa_x, kappa_y, -1 / lap_c, visc * lap_c,
                                 None if filt2 is None else filt2.to(real).cpu(), None)
        hit = [plan, None]
        _LEGACY_PLANS[key] = hit
    ftag = tag(f_hat)
    if hit[1] != ftag:
        hit[0].set_forcing(f_hat.reshape(n, n // 2 + 1).to(w.dtype))
        hit[1] = ftag
    return hit[0]


def _legacy_fusable(w, f, tensors):
    if not w.is_cuda or w.dim() < 3:
        return False
    if torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in [w, f] + list(tensors)):
        raise NotImplementedError(
            "torch-cfd_b200: the CUDA step is inference-only (no autograd through the solver); call it under "
            "torch.no_grad() / with detached tensors (fine-tuning through the step needs the reference's torch path)")
    return True


def imex_crank_nicolson_step(w, f, visc, delta_t, diam: float = 1, rfftmesh=None, laplacian=None, dealias_filter=None,
                             dealias: bool = False, output_rfft: bool = False, debug=False, **kwargs):
    """Same signature and return values as the reference (fno/data_gen/solvers.py:93-188):
    ``(w_next, dwdt, w, psi_h, res_h[, (kx, ky), laplacian, dealias_filter])`` for spectra ``w``, ``f`` of shape
    (B, *, n, n//2+1).  CUDA tensors only; a forcing shared by the batch runs as one fused launch, a per-sample
    forcing takes the CUDA explicit terms plus the reference's update formula as elementwise torch ops."""
    import math
    bsz, *size = w.shape
    assert (size[-1] - 1) * 2 == size[-2]
    n = size[-2]
    real = torch.float32 if w.dtype == torch.complex64 else torch.float64
    if rfftmesh is None or laplacian is None or (dealias and dealias_filter is None):
        kx0, ky0, lap0, filt0 = _legacy_tables(n, diam, real, w.device, scale_kmax=False)
        rfftmesh = (kx0, ky0) if rfftmesh is None else rfftmesh
        laplacian = lap0 if laplacian is None else laplacian
        dealias_filter = filt0 if dealias_filter is None else dealias_filter
    kx, ky = rfftmesh
    if f.ndim < w.ndim:
        f = f.unsqueeze(0)
    if not _legacy_fusable(w, f, [laplacian]):
        raise RuntimeError("torch-cfd_b200 runs the spectral step on CUDA devices only (no CPU fallback)")
    lead = w.shape[:-2]
    wb = w.reshape(-1, n, n // 2 + 1).contiguous()
    shared_f = f.numel() == n * (n // 2 + 1)
    lap_b = laplacian.to(w.device)
    with torch.no_grad(), torch.cuda.device(w.device):
        if shared_f:
            plan = _legacy_plan(wb, f, visc, (kx, ky), laplacian, dealias_filter, dealias)
            w_next, dwdt = torch.empty_like(wb), torch.empty_like(wb)
            plan.step(wb, w_next, dwdt, 1, [0.0], [float(delta_t)], [float(0.5 * delta_t)], 1 / delta_t)
            w_next, dwdt = w_next.reshape(w.shape), dwdt.reshape(w.shape)
            # res = dwdt + conv - nu lap w - f with (w_next - w)/dt = -conv + f + nu lap (w_next + w)/2:
            res_h = 0.5 * delta_t * visc * lap_b * dwdt
        else:
            plan = _legacy_plan(wb, torch.zeros(n, n // 2 + 1, dtype=w.dtype, device=w.device), visc, (kx, ky), laplacian,
                                dealias_filter, dealias)
            conv = torch.empty_like(wb)
            plan.explicit_terms(wb, conv)            # = -mask * rfft2(u w_x + v w_y)
            convection_h = -conv.reshape(w.shape)
            w_next = (-delta_t * convection_h + delta_t * f + (1.0 + 0.5 * delta_t * visc * lap_b) * w) / (
                1.0 - 0.5 * delta_t * visc * lap_b)
            dwdt = (w_next - w) / delta_t
            res_h = dwdt + convection_h - visc * lap_b * w - f
        psi_h = -w / lap_b
    if output_rfft:
        return w_next, dwdt, w, psi_h, res_h, (kx, ky), laplacian, dealias_filter
    return w_next, dwdt, w, psi_h, res_h


def update_residual(w_h, w_h_t, f_h, visc, rfftmesh, laplacian, dealias_filter=None, dealias=True, **kwargs):
    """``w_t + (u . grad) w - nu lap w - f`` in spectral space (fno/data_gen/solvers.py:49-90) on the CUDA kernels."""
    n = w_h.shape[-2]
    if not _legacy_fusable(w_h, f_h, [w_h_t, laplacian]):
        raise RuntimeError("torch-cfd_b200 runs the spectral residual on CUDA devices only (no CPU fallback)")
    wb = w_h.reshape(-1, n, n // 2 + 1).contiguous()
    wt = w_h_t.reshape(-1, n, n // 2 + 1).contiguous()
    with torch.no_grad(), torch.cuda.device(w_h.device):
        if f_h.numel() == n * (n // 2 + 1):
            plan = _legacy_plan(wb, f_h, visc, rfftmesh, laplacian, dealias_filter, dealias)
            out = torch.empty_like(wb)
            plan.residual(wb, wt, out)
            return out.reshape(w_h.shape)
        plan = _legacy_plan(wb, torch.zeros(n, n // 2 + 1, dtype=w_h.dtype, device=w_h.device), visc, rfftmesh, laplacian,
                            dealias_filter, dealias)
        out = torch.empty_like(wb)
        plan.residual(wb, wt, out)
        return out.reshape(w_h.shape) - f_h


def get_trajectory_imex_crank_nicolson(w0, f, visc: float = 1e-3, T: float = 1, delta_t: float = 1e-3,
                                       record_steps: int = 1, diam: float = 1, dealias: bool = True, subsample: int = 1,
                                       dtype: torch.dtype = None, pbar: bool = True, **kwargs):
    """Same signature and result as the reference (fno/data_gen/solvers.py:268-448): physical-space ``w0`` (B, n, n)
    and forcing ``f`` ((n, n) or (B, n, n)) -> dict of CPU tensors ``vorticity, vorticity_t, stream, residual``
    (B, record_steps, n // subsample, n // subsample) and ``t_steps``.  The un-recorded steps between two records
    run as ONE fused launch; transforms, residual and the bilinear subsampling run on the device."""
    import math
    if not w0.is_cuda:
        raise RuntimeError("torch-cfd_b200 runs the solver on CUDA devices only (no CPU fallback)")
    dtype = w0.dtype if dtype is None else dtype
    bsz, n = w0.size(0), w0.size(-1)
    ns = n // subsample
    total_steps = math.ceil(T / delta_t)
    record_every = math.floor(total_steps / record_steps)
    cdtype = torch.complex64 if dtype == torch.float32 else torch.complex128
    kx, ky, lap, filt = _legacy_tables(n, diam, dtype, w0.device)
    out_dtype = torch.get_default_dtype()
    vort, vort_t, stream, residual = [torch.empty(bsz, record_steps, ns, ns, dtype=out_dtype) for _ in range(4)]
    t_steps = torch.empty(record_steps)
    bar = None
    if pbar:
        try:
            from tqdm import tqdm
            bar = tqdm(total=total_steps)
        except ImportError:
            bar = None
    with torch.no_grad(), torch.cuda.device(w0.device):
        w_h = _fft.rfft2(w0.to(dtype).contiguous())
        f_h = _fft.rfft2(f.to(dtype).to(w0.device).contiguous())
        if f_h.ndim < w_h.ndim:
            f_h = f_h.unsqueeze(0)
        shared_f = f_h.numel() == n * (n // 2 + 1)
        kws = dict(diam=diam, rfftmesh=(kx, ky), laplacian=lap, dealias_filter=filt, dealias=dealias)
        c, j, t = 0, 0, 0.0
        while j < total_steps and c < record_steps:
            gap = record_every - 1 if j + record_every <= total_steps else 0
            if gap > 0 and shared_f:
                plan = _legacy_plan(w_h, f_h, visc, (kx, ky), lap, filt, dealias)
                nxt = torch.empty_like(w_h)
                plan.step(w_h, nxt, None, gap, [0.0], [float(delta_t)], [float(0.5 * delta_t)], 1 / (gap * delta_t))
                w_h = nxt
            else:
                for _ in range(gap):
                    w_h = imex_crank_nicolson_step(w_h, f_h, visc, delta_t, **kws)[0]
            j += gap
            w_h, w_h_t, _, psi_h, _ = imex_crank_nicolson_step(w_h, f_h, visc, delta_t, **kws)
            j += 1
            if not torch.isfinite(torch.view_as_real(w_h)).all():
                raise ValueError("Solution diverged")
            res_h = update_residual(w_h, w_h_t, f_h, visc, (kx, ky), lap, dealias_filter=filt, dealias=dealias)
            fields = [_fft.irfft2(z.to(cdtype).contiguous()) for z in (w_h, w_h_t, psi_h, res_h)]
            if subsample > 1:
                fields = [_fft.interpolate_bilinear(z, ns) for z in fields]
            for dst, z in zip((vort, vort_t, stream, residual), fields):
                dst[:, c] = z.to(out_dtype).cpu()
            for _ in range(gap + 1):
                t += delta_t
            t_steps[c] = t
            c += 1
            if bar is not None:
                bar.update(gap + 1)
        if bar is not None:
            bar.close()
    return dict(vorticity=vort, vorticity_t=vort_t, stream=stream, residual=residual, t_steps=t_steps)
