"""Drop-in ``SpectralConv3d`` (fno/fno3d.py:19-116), ``SpectralConvS`` (fno/sfno.py:332-394 on
fno/base.py:114-237) and ``SpectralConvT`` (fno/sfno.py:398-457) with the reference's constructor
arguments, parameter names/shapes/initialisation and ``forward`` signatures.

``forward`` is ONE autograd node: rfftn -> corner-block complex channel mix (+ bias) -> irfftn run as
five hand-written kernels that never materialise the full spectrum (csrc/sconv_kernels.cuh), and the
backward pass is the same kernels with conjugate-transposed tables.  CUDA fp32 tensors only; there
is no torch.fft / einsum on this path and no CPU fallback.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib


class _SConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, plan, delta, nbias, *params):
        w = [p.detach() for p in params[:4]]
        bias = [p.detach() for p in params[4:4 + nbias]] if nbias else None
        X, Y, T_in, t_pad, T_out, Ci, Co, mx, my, mt, norm = plan.geom
        xc = x.detach().contiguous()
        y = torch.empty(xc.shape[0], Co, X, Y, T_out, dtype=torch.float32, device=xc.device)
        need_grad = any(ctx.needs_input_grad)
        xhat = torch.empty(plan.xhat_elems(xc.shape[0]), dtype=torch.complex64, device=xc.device) if need_grad else None
        with torch.cuda.device(xc.device):
            plan.forward(xc, w, bias, delta, y, xhat)
        ctx.plan, ctx.delta, ctx.nbias = plan, delta, nbias
        ctx.save_for_backward(xhat, *params) if need_grad else None
        ctx.x_shape = xc.shape
        return y

    @staticmethod
    def backward(ctx, gy):
        xhat, *params = ctx.saved_tensors
        plan, delta, nbias = ctx.plan, ctx.delta, ctx.nbias
        w = [p.detach() for p in params[:4]]
        gy = gy.contiguous()
        need_x = ctx.needs_input_grad[0]
        need_w = any(ctx.needs_input_grad[4:8])
        need_b = nbias and any(ctx.needs_input_grad[8:8 + nbias])
        gx = torch.empty(ctx.x_shape, dtype=torch.float32, device=gy.device) if need_x else None
        # gradients come back in the parameters' own layout (complex, or real (..., 2))
        gw = [torch.empty_like(p) for p in params[:4]] if (need_w or need_b) else None
        gb = [torch.empty_like(p) for p in params[4:4 + nbias]] if need_b else None
        with torch.cuda.device(gy.device):
            plan.backward(gy, xhat, w, gx, gw, gb, delta)
        grads = [gx, None, None, None]
        grads += list(gw) if gw is not None else [None] * 4
        grads += list(gb) if gb is not None else [None] * nbias
        return tuple(grads)


def _check_input(x: torch.Tensor, Ci: int):
    if x.device.type != "cuda":
        raise RuntimeError("torch-cfd_b200 runs the spectral convolution on CUDA devices only (no CPU fallback); "
                           f"got a tensor on {x.device}")
    if x.dtype != torch.float32:
        raise TypeError(f"expected float32 input, got {x.dtype}")
    if x.dim() != 5 or x.shape[1] != Ci:
        raise ValueError(f"expected input of shape (b, {Ci}, X, Y, T), got {tuple(x.shape)}")


class _PlanCache:
    """Per-module cache of libtcfd handles keyed by (device, geometry)."""

    def __init__(self):
        self._plans = {}

    def get(self, device, batch, geom):
        key = (device.index if device.index is not None else torch.cuda.current_device(),) + tuple(geom)
        plan = self._plans.get(key)
        if plan is not None and plan.max_batch >= batch:
            return plan
        if plan is not None:
            plan.close()
        lib = _lib.load_library()
        with torch.cuda.device(key[0]):
            plan = _lib.SConv3dPlan(lib, *geom, max_batch=batch)
        self._plans[key] = plan
        return plan


def spectral_conv3d(x, weights: Sequence[torch.Tensor], bias: Optional[Sequence[torch.Tensor]], cache: _PlanCache,
                    modes, T_out=None, t_pad=0, delta=1.0, norm="backward"):
    """y = irfftn(W (.) rfftn(pad_t(x)), s=(X, Y, T_out + t_pad))[..., -T_out:] on the four corner blocks."""
    Ci, Co = weights[0].shape[0], weights[0].shape[1]
    _check_input(x, Ci)
    b, _, X, Y, T = x.shape
    mx, my, mt = modes
    T_out = T if T_out is None else int(T_out)
    geom = (X, Y, T, int(t_pad), T_out, Ci, Co, mx, my, mt, norm)
    plan = cache.get(x.device, b, geom)
    params = list(weights) + (list(bias) if bias is not None else [])
    return _SConvFn.apply(x, plan, float(delta), 0 if bias is None else len(bias), *params)


class SpectralConv3d(nn.Module):
    """3-D Fourier layer (reference: fno/fno3d.py:19-116): complex parameters ``weights1..4`` of shape
    (in, out, modes1, modes2, modes3), initialised ``rand / (in * out)``."""

    def __init__(self, in_channels, out_channels, modes1, modes2, modes3):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.modes1, self.modes2, self.modes3 = modes1, modes2, modes3
        self.scale = 1 / (in_channels * out_channels)
        for i in range(1, 5):
            setattr(self, f"weights{i}", nn.Parameter(
                self.scale * torch.rand(in_channels, out_channels, modes1, modes2, modes3, dtype=torch.cfloat)))
        self._cache = _PlanCache()

    def forward(self, x):
        w = [self.weights1, self.weights2, self.weights3, self.weights4]
        return spectral_conv3d(x, w, None, self._cache, (self.modes1, self.modes2, self.modes3))


class SpectralConvS(nn.Module):
    """Space-time Fourier layer of SFNO (reference: fno/sfno.py:332-394 over fno/base.py:114-237):
    ``weight`` = ParameterList of four real tensors (in, out, mx, my, mt, 2) initialised
    ``0.5 / (in * out) * rand``; optional ``bias`` = four real zero tensors (mx, my, mt, 2)."""

    def __init__(self, in_channels: int, out_channels: int, modes_x: int, modes_y: int, modes_t: int, dim: int = 3,
                 bias: bool = False, delta: float = 1, norm="backward") -> None:
        super().__init__()
        assert dim == 3, "only the (2+1)-D layer is implemented"
        self.in_channels, self.out_channels, self.dim = in_channels, out_channels, dim
        self.modes_x, self.modes_y, self.modes_t = modes_x, modes_y, modes_t
        self.delta, self.norm = delta, norm
        size = [in_channels, out_channels, modes_x, modes_y, modes_t, 2]
        gain = 0.5 / (in_channels * out_channels)
        self.weight = nn.ParameterList([nn.Parameter(gain * torch.rand(*size)) for _ in range(4)])
        self.bias = nn.ParameterList([nn.Parameter(gain * torch.zeros(*size[2:])) for _ in range(4)]) if bias else bias
        self._cache = _PlanCache()

    def _apply_conv(self, v, T_out=None, t_pad=0):
        bias = list(self.bias) if self.bias else None
        return spectral_conv3d(v, list(self.weight), bias, self._cache, (self.modes_x, self.modes_y, self.modes_t),
                               T_out=T_out, t_pad=t_pad, delta=self.delta, norm=self.norm)

    def forward(self, v, out_mesh_size=None, **kwargs):
        T_out = None
        if out_mesh_size is not None:
            if list(out_mesh_size[:2]) != list(v.shape[2:4]):
                raise NotImplementedError("resampling in x / y (out_mesh_size != input mesh) is not implemented")
            T_out = out_mesh_size[2]
        return self._apply_conv(v, T_out=T_out)


class SpectralConvT(SpectralConvS):
    """Temporal-resampling Fourier layer (reference: fno/sfno.py:398-457): optional front zero padding
    T -> 2T, output of ``out_steps`` time samples."""

    def __init__(self, in_channels: int, out_channels: int, modes_x: int, modes_y: int, modes_t: int,
                 delta: float = 1e-1, out_steps: int = None, norm: str = "backward", bias: bool = True,
                 temporal_padding: bool = False, postprocess: nn.Module = None, **kwargs) -> None:
        super().__init__(in_channels, out_channels, modes_x, modes_y, modes_t, norm=norm, delta=delta, bias=bias)
        self.out_steps = out_steps
        self.temporal_padding = temporal_padding
        if postprocess is not None and not isinstance(postprocess, nn.Identity):
            raise NotImplementedError("a spectral postprocess (e.g. HelmholtzProjection) is outside the fused path")
        self.postprocess = nn.Identity()

    def forward(self, v, out_steps: int = None):
        if out_steps is None and self.out_steps is not None:
            out_steps = self.out_steps
        t_pad = v.size(-1) if self.temporal_padding else 0
        return self._apply_conv(v, T_out=out_steps, t_pad=t_pad)
