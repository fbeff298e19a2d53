"""Minimal periodic ``Grid`` for the spectral hot path.

Only what hot path A touches is provided (reference: torch_cfd/grids.py:36-218 -- constructor,
``axes``/``mesh`` and the ``fft_axes``/``fft_mesh``/``rfft_mesh`` family, all host-side and run
once per equation).  The staggered-grid machinery (GridArray / GridVariable / boundary conditions)
belongs to the finite-volume solver and is out of scope.
"""
from __future__ import annotations

import math
import numbers
import operator
from typing import Optional, Sequence, Tuple, Union

import torch


class Grid:
    """Uniform n-d box.  ``Grid(shape, step=...)`` or ``Grid(shape, domain=...)``; ``domain`` is a
    number (upper bound of every axis) or one ``(lower, upper)`` pair per axis."""

    def __init__(self, shape: Sequence[int], step: Optional[Union[float, Sequence[float]]] = None,
                 domain: Optional[Union[float, Sequence[Tuple[float, float]]]] = None,
                 device: Optional[torch.device] = "cpu"):
        self.shape = tuple(operator.index(s) for s in shape)
        nd = len(self.shape)
        if step is not None and domain is not None:
            raise TypeError("cannot provide both step and domain")
        if domain is not None:
            if isinstance(domain, (int, float)):
                domain = ((0, domain),) * nd
            else:
                if len(domain) != nd:
                    raise ValueError(f"length of domain does not match ndim: {len(domain)} != {nd}")
                for b in domain:
                    if len(b) != 2:
                        raise ValueError(f"domain is not sequence of pairs of numbers: {domain}")
            domain = tuple((float(lo), float(hi)) for lo, hi in domain)
        else:
            if step is None:
                step = 1
            if isinstance(step, numbers.Number):
                step = (step,) * nd
            elif len(step) != nd:
                raise ValueError(f"length of step does not match ndim: {len(step)} != {nd}")
            domain = tuple((0.0, float(s * n)) for s, n in zip(step, self.shape))
        self.domain = domain
        self.step = tuple((hi - lo) / n for (lo, hi), n in zip(domain, self.shape))
        self.device = device

    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def cell_center(self) -> Tuple[float, ...]:
        return self.ndim * (0.5,)

    @property
    def cell_faces(self) -> Tuple[Tuple[float, ...], ...]:
        d = self.ndim
        return tuple(tuple(1.0 if i == j else 0.5 for j in range(d)) for i in range(d))

    def axes(self, offset: Optional[Sequence[float]] = None) -> Tuple[torch.Tensor, ...]:
        """Grid-point coordinates per axis: lower + (i + offset) * step."""
        if offset is None:
            offset = self.cell_center
        if len(offset) != self.ndim:
            raise ValueError(f"unexpected offset length: {len(offset)} vs {self.ndim}")
        return tuple(lo + (torch.arange(n) + o) * h
                     for (lo, _), o, n, h in zip(self.domain, offset, self.shape, self.step))

    def fft_axes(self) -> Tuple[torch.Tensor, ...]:
        """Ordinal FFT frequencies per axis (multiply by 2 pi for angular ones)."""
        return tuple(torch.fft.fftfreq(n, d=h) for n, h in zip(self.shape, self.step))

    def mesh(self, offset: Optional[Sequence[float]] = None) -> Tuple[torch.Tensor, ...]:
        x, y = torch.meshgrid(*self.axes(offset), indexing="ij")
        return x.to(self.device), y.to(self.device)

    def fft_mesh(self) -> Tuple[torch.Tensor, ...]:
        kx, ky = torch.meshgrid(*self.fft_axes(), indexing="ij")
        return kx.to(self.device), ky.to(self.device)

    def rfft_mesh(self) -> Tuple[torch.Tensor, ...]:
        k_max = math.floor(self.shape[-1] / 2.0)
        return tuple(m[..., : k_max + 1] for m in self.fft_mesh())

    def __repr__(self):
        return f"Grid(shape={self.shape}, domain={self.domain})"
