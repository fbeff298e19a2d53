"""Time-independent forcings for the spectral solver (reference: torch_cfd/forcings.py:60-210).

The reference re-evaluates the forcing field (mesh + cos + rfft2) in every RK substage although
both shipped spectral forcings ignore their state argument; here the field is evaluated once
per equation on the host and its spectrum is added inside the fused kernel.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from .grids import Grid


class Field:
    """Bare stand-in for the reference's GridArray: the solver only reads ``.data``."""

    def __init__(self, data: torch.Tensor, offset=None, grid: Optional[Grid] = None):
        self.data, self.offset, self.grid = data, offset, grid


class ForcingFn(nn.Module):
    """Base class: ``forcing(grid, state)`` returns a vorticity field (``vorticity=True``) or a
    pair of velocity-component fields."""

    state_independent = False

    def __init__(self, grid: Grid, scale: float = 1, wave_number: int = 1, diam: float = 1.0,
                 swap_xy: bool = False, vorticity: bool = False, offsets=None, device=None, **kwargs):
        super().__init__()
        self.grid = grid
        self.scale = scale
        self.wave_number = wave_number
        self.diam = diam
        self.swap_xy = swap_xy
        self.vorticity = vorticity
        self.offsets = grid.cell_faces if offsets is None else offsets
        self.device = grid.device if device is None else device

    def velocity_eval(self, grid, velocity):
        raise NotImplementedError

    def vorticity_eval(self, grid, vorticity):
        raise NotImplementedError

    def forward(self, grid=None, velocity=None, vorticity=None):
        if not self.vorticity:
            return self.velocity_eval(grid, velocity)
        return self.vorticity_eval(grid, vorticity)


class KolmogorovForcing(ForcingFn):
    """f = (a sin(k c y), 0) in velocity form, or its curl -a k c cos(k c y) in vorticity form,
    c = 2 pi / diam (x and y exchanged when ``swap_xy``)."""

    state_independent = True

    def __init__(self, diam=2 * torch.pi, offsets=((0, 0), (0, 0)), vorticity=False, *args, **kwargs):
        super().__init__(*args, diam=diam, offsets=offsets, vorticity=vorticity, **kwargs)

    def _coordinate(self, grid):
        grid = self.grid if grid is None else grid
        if self.swap_xy:
            return grid, grid.mesh(self.offsets[1])[0], self.offsets[1]
        return grid, grid.mesh(self.offsets[0])[1], self.offsets[0]

    def velocity_eval(self, grid, velocity=None) -> Tuple[Field, Field]:
        grid, s, off = self._coordinate(grid)
        c = 2 * torch.pi / self.diam
        wave = Field(self.scale * torch.sin(self.wave_number * c * s), off, grid)
        zero = Field(torch.zeros_like(wave.data), off, grid)
        return (zero, wave) if self.swap_xy else (wave, zero)

    def vorticity_eval(self, grid, vorticity=None) -> Field:
        grid, s, off = self._coordinate(grid)
        c = 2 * torch.pi / self.diam
        return Field(-self.scale * self.wave_number * c * torch.cos(self.wave_number * c * s), off, grid)
