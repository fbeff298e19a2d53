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
