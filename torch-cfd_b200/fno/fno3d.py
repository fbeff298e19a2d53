"""``FNO3d`` (reference: fno/fno3d.py:119-236): lift -> n x [spectral conv + pointwise MLP + pointwise
skip, GELU] -> project.  Same constructor arguments, sub-module names (``p, spectral_conv, mlp, w,
activation, q``) and therefore ``state_dict`` keys as upstream; the spectral convolutions are the
fused libtcfd layers, the 1x1x1 channel mixes and GELU stay torch (cuDNN/cuBLAS) ops -- SURVEY 8a
row B5 / 8f rank 2: their fusion into the inverse-FFT epilogue is the next step on this path."""
import torch
import torch.nn as nn
import torch.nn.functional as F

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

    def forward(self, x):
        """x: (b, input_channel + 3, X, Y, T) -> ((b, X, Y, T), None) like upstream (fno/fno3d.py:205-236)."""
        x = self.p(x)
        x = F.pad(x, [0, 0, self.padding, self.padding, self.padding, self.padding], mode="circular")
        for conv, mlp, w, nonlinear in zip(self.spectral_conv, self.mlp, self.w, self.activation):
            x = nonlinear(mlp(conv(x)) + w(x))
        if self.padding != 0:
            x = x[..., self.padding:-self.padding, self.padding:-self.padding, :]
        x = self.q(x)
        return x.squeeze(1), None
