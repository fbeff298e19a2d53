"""``FNO3d`` (reference: fno/fno3d.py:119-236): lift -> n x [spectral conv + pointwise MLP + pointwise
skip, GELU] -> project.  Same constructor arguments, sub-module names (``p, spectral_conv, mlp, w,
activation, q``) and therefore ``state_dict`` keys as upstream; the spectral convolutions are the
fused libtcfd layers.  Inference (no grad, CUDA, fp32, width <= 32) also fuses the pointwise glue --
lifting ``p``, per layer ``nonlinear(mlp(conv(x)) + w(x))``, projection ``q`` -- into one libtcfd
launch each (csrc/fno_glue.cu); with gradients enabled the 1x1x1 channel mixes and
GELUs are the reference's torch ops so that autograd sees them."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib
from .spectral_conv import SpectralConv3d


class MLP(nn.Module):
    """Pointwise two-layer channel MLP (reference: fno/fno3d.py:119-130)."""

    def __init__(self, in_channels, out_channels, mid_channels, activation=True):
        super().__init__()
        self.mlp1 = nn.Conv3d(in_channels, mid_channels, 1)
        self.mlp2 = nn.Conv3d(mid_channels, out_channels, 1)
        self.activation = nn.GELU() if activation else nn.Identity()

    def forward(self, x):
        return self.mlp2(self.activation(self.mlp1(x)))


class FNO3d(nn.Module):
    def __init__(self, modes1, modes2, modes3, width, dim=3, input_channel=10, num_spectral_layers=4,
                 last_activation=False, padding=0, extra_mlp=True, channel_expansion=128, debug=False):
        super().__init__()
        self.modes1, self.modes2, self.modes3, self.width = modes1, modes2, modes3, width
        self.input_channel, self.padding = input_channel, padding
        self.extra_mlp, self.channel_expansion, self.debug = extra_mlp, channel_expansion, debug
        self.p = nn.Conv3d(input_channel + dim, width, 1)
        self.spectral_conv = nn.ModuleList(
            [SpectralConv3d(width, width, modes1, modes2, modes3) for _ in range(num_spectral_layers)])
        self.mlp = nn.ModuleList([MLP(width, width, width) for _ in range(num_spectral_layers)])
        self.w = nn.ModuleList([nn.Conv3d(width, width, 1) for _ in range(num_spectral_layers)])
        self.activation = nn.ModuleList([nn.GELU() for _ in range(num_spectral_layers - 1)])
        self.activation.append(nn.GELU() if last_activation else nn.Identity())
        self.q = MLP(width, 1, channel_expansion, activation=last_activation)

    def _fused_ok(self, x):
        gelu_ok = all(isinstance(a, (nn.GELU, nn.Identity)) and getattr(a, "approximate", "none") == "none"
                      for a in list(self.activation) + [m.activation for m in self.mlp] + [self.q.activation])
        return (not torch.is_grad_enabled() and x.is_cuda and x.dtype == torch.float32 and self.width <= 32
                and self.q.mlp2.out_channels == 1 and gelu_ok)

    def _glue_weights(self, k, mlp, w):
        """Host copies of layer k's pointwise weights (they travel as a kernel parameter), cached until a
        parameter is modified in place (``_version``) or replaced."""
        ps = (mlp.mlp1.weight, mlp.mlp1.bias, mlp.mlp2.weight, mlp.mlp2.bias, w.weight, w.bias)
        tag = tuple((id(t), t._version, str(t.device)) if t is not None else None for t in ps)
        cache = self.__dict__.setdefault("_glue_cache", {})
        hit = cache.get(k)
        if hit is None or hit[0] != tag:
            hit = (tag, _lib.fno_glue_host_weights(*ps))
            cache[k] = hit
        return hit[1]

    def _collapsed_q(self):
        """W = W2 W1 (1 x width), b = W2 b1 + b2 of the activation-free projection MLP, accumulated in
        float64 and cached until a parameter changes."""
        ps = (self.q.mlp1.weight, self.q.mlp1.bias, self.q.mlp2.weight, self.q.mlp2.bias)
        tag = tuple((id(t), t._version, str(t.device)) if t is not None else None for t in ps)
        hit = self.__dict__.get("_q_cache")
        if hit is None or hit[0] != tag:
            w1 = ps[0].detach().double().reshape(ps[0].shape[0], -1)
            w2 = ps[2].detach().double().reshape(1, -1)
            b1 = ps[1].detach().double() if ps[1] is not None else torch.zeros(w1.shape[0], dtype=torch.float64, device=w1.device)
            b2 = ps[3].detach().double() if ps[3] is not None else torch.zeros(1, dtype=torch.float64, device=w1.device)
            hit = (tag, ((w2 @ w1).float().contiguous(), (w2 @ b1 + b2).float().contiguous()))
            self.__dict__["_q_cache"] = hit
        return hit[1]

    def _forward_fused(self, x):
        """Inference path: every pointwise stage is one libtcfd launch (csrc/fno_glue.cu)."""
        lib = _lib.load_library()
        with torch.cuda.device(x.device):
            x = _lib.fno_pointwise_linear(lib, x.contiguous(), self.p.weight, self.p.bias)
            if self.padding != 0:
                x = F.pad(x, [0, 0, self.padding, self.padding, self.padding, self.padding], mode="circular").contiguous()
            for k, (conv, mlp, w, nonlinear) in enumerate(zip(self.spectral_conv, self.mlp, self.w, self.activation)):
                x = _lib.fno_layer_glue(lib, conv(x), x, self._glue_weights(k, mlp, w), isinstance(nonlinear, nn.GELU))
            if self.padding != 0:
                x = x[..., self.padding:-self.padding, self.padding:-self.padding, :].contiguous()
            if isinstance(self.q.activation, nn.GELU):
                x = _lib.fno_project(lib, x, self.q.mlp1.weight, self.q.mlp1.bias, self.q.mlp2.weight,
                                     self.q.mlp2.bias, True)
            else:
                # no activation between the two 1x1x1 convolutions of q (last_activation=False, the
                # default): the projection is ONE linear map width -> 1, evaluated as such
                wq, bq = self._collapsed_q()
                x = _lib.fno_pointwise_linear(lib, x, wq, bq)
        return x.squeeze(1), None

    def forward(self, x):
        """x: (b, input_channel + 3, X, Y, T) -> ((b, X, Y, T), None) like upstream (fno/fno3d.py:205-236)."""
        if self._fused_ok(x):
            return self._forward_fused(x)
        x = self.p(x)
        x = F.pad(x, [0, 0, self.padding, self.padding, self.padding, self.padding], mode="circular")
        for conv, mlp, w, nonlinear in zip(self.spectral_conv, self.mlp, self.w, self.activation):
            x = nonlinear(mlp(conv(x)) + w(x))
        if self.padding != 0:
            x = x[..., self.padding:-self.padding, self.padding:-self.padding, :]
        x = self.q(x)
        return x.squeeze(1), None
