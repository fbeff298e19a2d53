"""Drop-in ``SpectralConv3d`` (fno/fno3d.py:19-116), ``SpectralConvS`` (fno/sfno.py:332-394 on
fno/base.py:114-237) and ``SpectralConvT`` (fno/sfno.py:398-457) with the reference's constructor
arguments, parameter names/shapes/initialisation and ``forward`` signatures.

``forward`` is ONE autograd node: rfftn -> corner-block complex channel mix (+ bias) -> irfftn run as
five hand-written kernels that never materialise the full spectrum (csrc/sconv_kernels.cuh), and the
backward pass is the same kernels with conjugate-transposed tables.  A change of the x / y mesh
(``out_mesh_size``) or a spectral ``postprocess`` (HelmholtzProjection) splits the layer into its two halves
with the small spectral tensor handed to torch in between (see ``spectral_conv3d``).  CUDA fp32 tensors only;
there is no torch.fft / einsum on this path and no CPU fallback.
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


class _SConvAnalysisFn(torch.autograd.Function):
    """x -> truncated output spectrum yhat (b, Co, 2mx, 2my, mt) complex64: pruned rfftn + corner-block channel mix
    (+ bias).  First half of the layer (tcfd_sconv3d_analysis); used when something sits between the halves."""

    @staticmethod
    def forward(ctx, x, plan, delta, nbias, *params):
        w = [p.detach() for p in params[:4]]
        bias = [p.detach() for p in params[4:4 + nbias]] if nbias else None
        xc = x.detach().contiguous()
        yhat = torch.empty(plan.yhat_shape(xc.shape[0]), dtype=torch.complex64, device=xc.device)
        need_grad = any(ctx.needs_input_grad)
        xhat = torch.empty(plan.xhat_elems(xc.shape[0]), dtype=torch.complex64, device=xc.device) if need_grad else None
        with torch.cuda.device(xc.device):
            plan.analysis(xc, w, bias, delta, yhat, xhat)
        ctx.plan, ctx.delta, ctx.nbias, ctx.x_shape = plan, delta, nbias, xc.shape
        ctx.save_for_backward(xhat, *params) if need_grad else None
        return yhat

    @staticmethod
    def backward(ctx, gyhat):
        xhat, *params = ctx.saved_tensors
        plan, delta, nbias = ctx.plan, ctx.delta, ctx.nbias
        w = [p.detach() for p in params[:4]]
        need_x = ctx.needs_input_grad[0]
        need_w = any(ctx.needs_input_grad[4:8])
        need_b = nbias and any(ctx.needs_input_grad[8:8 + nbias])
        gyhat = gyhat.contiguous()
        gx = torch.empty(ctx.x_shape, dtype=torch.float32, device=gyhat.device) if need_x else None
        gw = [torch.empty_like(p) for p in params[:4]] if (need_w or need_b) else None
        gb = [torch.empty_like(p) for p in params[4:4 + nbias]] if need_b else None
        with torch.cuda.device(gyhat.device):
            plan.analysis_backward(gyhat, xhat, w, gx, gw, gb, delta)
        grads = [gx, None, None, None]
        grads += list(gw) if gw is not None else [None] * 4
        grads += list(gb) if gb is not None else [None] * nbias
        return tuple(grads)


class _SConvSynthesisFn(torch.autograd.Function):
    """truncated spectrum yhat (b, Co, 2mx, 2my, mt) -> y (b, Co, X, Y, T_out): zero-padded inverse transforms of the
    plan's geometry (tcfd_sconv3d_synthesis); the adjoint is the matching pruned analysis of the cotangent."""

    @staticmethod
    def forward(ctx, yhat, plan):
        X, Y, T_in, t_pad, T_out, Ci, Co, mx, my, mt, norm = plan.geom
        yh = yhat.detach().contiguous()
        y = torch.empty(yh.shape[0], Co, X, Y, T_out, dtype=torch.float32, device=yh.device)
        with torch.cuda.device(yh.device):
            plan.synthesis(yh, y)
        ctx.plan, ctx.yh_shape = plan, yh.shape
        return y

    @staticmethod
    def backward(ctx, gy):
        gy = gy.contiguous()
        gyhat = torch.empty(ctx.yh_shape, dtype=torch.complex64, device=gy.device)
        with torch.cuda.device(gy.device):
            ctx.plan.synthesis_backward(gy, gyhat)
        return gyhat, None


def _kept_positions(n_in: int, m: int, n_out: int, device):
    """Index (in a length-n_out frequency axis) of every kept row of a truncated spectrum whose 2m rows are the m
    lowest and the m highest indices of a length-n_in axis -- irfftn(s=...) pads / trims at the END of an axis, so
    an entry keeps its INDEX, not its frequency (fno/base.py:236) -- and the mask of rows that survive the trim."""
    pos = torch.cat([torch.arange(m), torch.arange(n_in - m, n_in)]).to(device)
    return pos, pos < n_out


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
        self._plans = _lib.HandleStore()

    def get(self, device, batch, geom):
        key = (device.index if device.index is not None else torch.cuda.current_device(),) + tuple(geom)
        plan = self._plans.get(key)
        if plan is not None and plan.max_batch >= batch:
            return plan
        # a superseded (smaller) plan is NOT closed here: an autograd graph built with it may still run its
        # backward; the handle is released when the last reference goes (SConv3dPlan.__del__)
        lib = _lib.load_library()
        with torch.cuda.device(key[0]):
            plan = _lib.SConv3dPlan(lib, *geom, max_batch=batch)
        self._plans[key] = plan
        return plan


def spectral_conv3d(x, weights: Sequence[torch.Tensor], bias: Optional[Sequence[torch.Tensor]], cache: _PlanCache,
                    modes, T_out=None, t_pad=0, delta=1.0, norm="backward", out_xy=None, postprocess=None):
    """y = irfftn(post(W (.) rfftn(pad_t(x))), s=(X_out, Y_out, T_out + t_pad))[..., -T_out:] on the four corner blocks.

    Default (same mesh, no spectral post-process): ONE autograd node, five launches.  With ``out_xy`` != (X, Y)
    (fno/base.py:229-237) or a ``postprocess`` module acting on the spectrum (fno/sfno.py:452) the layer runs as its
    two halves: analysis on the input geometry -> the truncated spectrum is placed into the FULL spectrum of the
    output geometry (entries keep their index, as irfftn(s=...) does) -> postprocess -> synthesis on a handle of the
    output geometry.  Only the small spectral tensor is touched by torch ops in between; both transforms stay in
    the CUDA kernels, and autograd sees the two halves as nodes."""
    Ci, Co = weights[0].shape[0], weights[0].shape[1]
    _check_input(x, Ci)
    b, _, X, Y, T = x.shape
    mx, my, mt = modes
    T_out = T if T_out is None else int(T_out)
    geom = (X, Y, T, int(t_pad), T_out, Ci, Co, mx, my, mt, norm)
    plan = cache.get(x.device, b, geom)
    params = list(weights) + (list(bias) if bias is not None else [])
    nb = 0 if bias is None else len(bias)
    Xo, Yo = (X, Y) if out_xy is None else (int(out_xy[0]), int(out_xy[1]))
    if (Xo, Yo) == (X, Y) and postprocess is None:
        return _SConvFn.apply(x, plan, float(delta), nb, *params)
    yhat = _SConvAnalysisFn.apply(x, plan, float(delta), nb, *params)
    # full spectrum of the output geometry: (b, Co, Xo, Yo, mt_full) in plain FFT index order
    Tn_in = T + int(t_pad)
    mt_full = Tn_in // 2 + 1 if postprocess is not None else mt
    px, vx = _kept_positions(X, mx, Xo, x.device)
    py, vy = _kept_positions(Y, my, Yo, x.device)
    full = torch.zeros(b, Co, Xo, Yo, mt_full, dtype=yhat.dtype, device=x.device)
    full[:, :, px[vx][:, None], py[vy][None, :], :mt] = yhat[:, :, vx.nonzero()[:, 0][:, None], vy.nonzero()[:, 0][None, :], :]
    if postprocess is not None:
        full = postprocess(full)
    geom_out = (Xo, Yo, T, int(t_pad), T_out, Co, Co, Xo // 2, Yo // 2, mt_full, norm)
    plan_out = cache.get(x.device, b, geom_out)
    return _SConvSynthesisFn.apply(full, plan_out)


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

    def _apply_conv(self, v, T_out=None, t_pad=0, out_xy=None, postprocess=None):
        bias = list(self.bias) if self.bias else None
        return spectral_conv3d(v, list(self.weight), bias, self._cache, (self.modes_x, self.modes_y, self.modes_t),
                               T_out=T_out, t_pad=t_pad, delta=self.delta, norm=self.norm, out_xy=out_xy,
                               postprocess=postprocess)

    def forward(self, v, out_mesh_size=None, **kwargs):
        """``out_mesh_size`` = (X_out, Y_out, T_out): output mesh of ``irfftn(s=out_mesh_size)`` (fno/base.py:229-237)."""
        T_out, out_xy = None, None
        if out_mesh_size is not None:
            out_xy, T_out = tuple(out_mesh_size[:2]), out_mesh_size[2]
        return self._apply_conv(v, T_out=T_out, out_xy=out_xy)


class SpectralConvT(SpectralConvS):
    """Temporal-resampling Fourier layer (reference: fno/sfno.py:398-457): optional front zero padding
    T -> 2T, output of ``out_steps`` time samples."""

    def __init__(self, in_channels: int, out_channels: int, modes_x: int, modes_y: int, modes_t: int,
                 delta: float = 1e-1, out_steps: int = None, norm: str = "backward", bias: bool = True,
                 temporal_padding: bool = False, postprocess: nn.Module = None, **kwargs) -> None:
        super().__init__(in_channels, out_channels, modes_x, modes_y, modes_t, norm=norm, delta=delta, bias=bias)
        self.out_steps = out_steps
        self.temporal_padding = temporal_padding
        self.postprocess = nn.Identity() if postprocess is None else postprocess

    def forward(self, v, out_steps: int = None):
        if out_steps is None and self.out_steps is not None:
            out_steps = self.out_steps
        t_pad = v.size(-1) if self.temporal_padding else 0
        post = None if isinstance(self.postprocess, nn.Identity) else self.postprocess
        return self._apply_conv(v, T_out=out_steps, t_pad=t_pad, postprocess=post)
