"""``SFNO`` and its building blocks (reference: fno/sfno.py:26-620 on fno/base.py:60-400) around the libtcfd
spectral convolutions: same constructor arguments, sub-module / parameter names (``state_dict`` keys of an
upstream checkpoint load unchanged) and ``forward`` signatures.

What runs where: every Fourier layer (``SpectralConvS`` in the latent blocks, ``SpectralConvT`` in the lifting and
output operators) is the fused CUDA layer of ``spectral_conv.py``; the positional encoding, the group norm and the
pointwise 1x1x1 channel maps are the reference's own torch modules (they are cuDNN / elementwise work, not part
of the transform path); ``HelmholtzProjection`` is a per-mode 2x2 projection applied to the spectrum between the two
halves of ``SpectralConvT`` (see ``spectral_conv.spectral_conv3d``)."""
from __future__ import annotations

from copy import deepcopy
from typing import List, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..spectral import fft_expand_dims, fft_mesh_2d, spectral_div_2d, spectral_grad_2d, spectral_laplacian_2d
from .spectral_conv import SpectralConvS, SpectralConvT

ActivationType = Union[str]


class LayerNormnd(nn.GroupNorm):
    """GroupNorm with one group = layer norm over (C, *) of a (b, C, *) tensor (fno/base.py:60-83)."""

    def __init__(self, num_channels, eps=1e-07, elementwise_affine=True, device=None, dtype=None):
        super().__init__(num_groups=1, num_channels=num_channels, eps=eps, affine=elementwise_affine,
                         device=device, dtype=dtype)


class PointwiseFFN(nn.Module):
    """Two pointwise (kernel size 1) convolutions with an activation in between (fno/base.py:86-111)."""

    def __init__(self, in_channels: int, out_channels: int, mid_channels: int, activation: ActivationType = "ReLU",
                 dim: int = 3):
        super().__init__()
        convs = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}
        if dim not in convs:
            raise ValueError(f"Unsupported dimension: {dim}, expected 1, 2, or 3")
        self.linear1 = convs[dim](in_channels, mid_channels, 1)
        self.linear2 = convs[dim](mid_channels, out_channels, 1)
        self.activation = getattr(nn, activation)()

    def forward(self, v: torch.Tensor):
        return self.linear2(self.activation(self.linear1(v)))


class SpaceTimePositionalEncoding(nn.Module):
    """Sinusoidal space-time encoding added to the (b, 1, x, y, t) input (fno/sfno.py:26-113): channels
    0..2 are the coordinates (x, y in [0, 1], t on a ``max_time_steps`` ruler), the others
    ``exp(beta t) sin|cos(pi (k+1) t)``; with ``spatial_random_feats`` a product basis in (x, y, t) projected to
    ``num_channels`` by a pointwise convolution.  The table ``pe`` is a plain attribute rebuilt when the mesh changes."""

    def __init__(self, modes_x: int = 16, modes_y: int = 16, modes_t: int = 5, num_channels: int = 20,
                 input_shape: Union[List, Tuple] = (64, 64, 10), spatial_random_feats: bool = False,
                 max_time_steps: int = 100, time_exponential_scale: float = 1e-2, **kwargs):
        super().__init__()
        assert num_channels % 2 == 0 and num_channels > 3
        self.num_channels = num_channels
        self.max_time_steps = max_time_steps
        self.time_exponential_scale = time_exponential_scale
        self.modes_x, self.modes_y, self.modes_t = modes_x, modes_y, modes_t
        self._pe = self._pe_expanded if spatial_random_feats else self._pe
        self._pe(*input_shape)
        if spatial_random_feats:
            self.proj = nn.Conv3d(modes_x * modes_y * modes_t + 3, num_channels, kernel_size=1)
        else:
            self.proj = nn.Identity()

    def _coords(self, nx, ny, nt):
        gx, gy = torch.linspace(0, 1, nx), torch.linspace(0, 1, ny)
        gt = torch.linspace(0, 1, self.max_time_steps + 1)[1: nt + 1]
        return gt, torch.meshgrid(gx, gy, gt, indexing="ij")

    def _pe_expanded(self, *shape):
        nx, ny, nt = shape
        _, (gx, gy, gt) = self._coords(nx, ny, nt)
        feats = [gx, gy, gt]
        trig = lambda idx: torch.sin if idx % 2 == 0 else torch.cos
        for i in range(1, self.modes_x + 1):
            for j in range(1, self.modes_y + 1):
                for k in range(1, self.modes_t + 1):
                    feats.append(1 / (i * j * k) * torch.exp(self.time_exponential_scale * gt)
                                 * trig(i)(torch.pi * i * gx) * trig(j)(torch.pi * j * gy) * trig(k)(torch.pi * k * gt))
        self.pe = torch.stack(feats).unsqueeze(0)

    def _pe(self, *shape):
        nx, ny, nt = shape
        t1d, (gx, gy, gt) = self._coords(nx, ny, nt)
        feats = [gx, gy, gt]
        for k in range(self.num_channels - 3):
            basis = torch.sin if k % 2 == 0 else torch.cos
            col = torch.exp(self.time_exponential_scale * t1d) * basis(torch.pi * (k + 1) * t1d)
            feats.append(col.reshape(1, 1, nt).repeat(nx, ny, 1))
        self.pe = torch.stack(feats).unsqueeze(0)  # (1, num_channels, nx, ny, nt)

    def forward(self, v: torch.Tensor):
        if self.pe is None or self.pe.shape[-3:] != v.shape[-3:]:
            *_, nx, ny, nt = v.size()
            self._pe(nx, ny, nt)
        return v + self.proj(self.pe.to(v.dtype).to(v.device))


class HelmholtzProjection(nn.Module):
    """Projection of a 2-component spectrum onto divergence-free fields, ``w^ = u^ - grad(div u^) / lap``
    (fno/sfno.py:116-193); buffers ``lap, kx, ky`` as upstream.  Input / output (b, 2, nx, ny, nt//2+1) complex."""

    def __init__(self, n_grid: int = 64, diam: float = 2 * torch.pi, dtype: torch.dtype = torch.float32):
        super().__init__()
        self.n_grid = n_grid
        self.diam = diam
        self._update_fft_mesh(n_grid, diam, dtype)

    def _update_fft_mesh(self, n, diam=None, dtype=torch.float32):
        diam = diam if diam is not None else self.diam
        kx, ky = fft_mesh_2d(n, diam)
        lap = spectral_laplacian_2d(fft_mesh=(kx, ky))
        dev = self.lap.device if hasattr(self, "lap") else None
        self.register_buffer("lap", lap.to(dtype).to(dev))
        self.register_buffer("kx", kx.to(dtype).to(dev))
        self.register_buffer("ky", ky.to(dtype).to(dev))

    @staticmethod
    def div(uhat, fft_mesh):
        kx, ky = fft_expand_dims(fft_mesh, uhat.size(0))
        return spectral_div_2d([uhat[:, 0], uhat[:, 1]], (kx, ky))

    @staticmethod
    def grad(uhat, fft_mesh):
        kx, ky = fft_expand_dims(fft_mesh, uhat.size(0))
        return torch.stack(spectral_grad_2d(uhat, (kx, ky)), dim=1)

    def forward(self, uhat):
        bsz, _, nx, ny, nt = uhat.shape
        if nx != self.kx.shape[0]:
            # evaluation on another mesh: rebuild the tables first (upstream rebuilds them AFTER reading the old
            # ones, which fails on the shape mismatch -- SURVEY appendix B style defect, not reproduced)
            self._update_fft_mesh(nx, dtype=self.kx.dtype)
        fft_mesh = (self.kx, self.ky)
        grad_div_u = self.grad(self.div(uhat, fft_mesh), fft_mesh)
        lap = self.lap[None, None, :, :, None].expand(bsz, 2, nx, ny, nt)
        return uhat - grad_div_u / lap


class LiftingOperator(nn.Module):
    """(b, 1, x, y, t_in) -> (b, width, x, y, latent_steps) (fno/sfno.py:196-260): positional encoding, layer norm,
    pointwise lift, then ``act(v[..., -1:] + mlp(sconv(v)))`` with a ``SpectralConvT`` that resamples time."""

    def __init__(self, width: int, modes_x: int, modes_y: int, modes_t: int, latent_steps: int = 10,
                 norm: str = "backward", activation: ActivationType = "GELU", beta: float = 0.1,
                 spatial_random_feats: bool = False, channel_expansion: int = 4, nonlinear: bool = True, **kwargs) -> None:
        super().__init__()
        pe_modes_t = modes_t - 1 if modes_t % 2 != 0 else modes_t
        self.pe = SpaceTimePositionalEncoding(modes_x // 2, modes_y // 2, pe_modes_t // 2, num_channels=width,
                                              time_exponential_scale=beta, spatial_random_feats=spatial_random_feats)
        in_channels = self.pe.num_channels
        self.norm = LayerNormnd(in_channels)
        self.proj = nn.Conv3d(in_channels, width, kernel_size=1)
        self.sconv = SpectralConvT(width, width, modes_x, modes_y, modes_t, out_steps=latent_steps, norm=norm, bias=False)
        self.latent_steps = latent_steps
        if nonlinear:
            self.activation = getattr(nn, activation)()
            self.mlp = PointwiseFFN(width, width, channel_expansion * width, activation)
        else:
            self.activation = nn.Identity()
            self.mlp = nn.Conv3d(width, width, kernel_size=1)

    def forward(self, v):
        assert self.latent_steps <= v.size(-1)
        v = self.proj(self.norm(self.pe(v)))
        w = self.mlp(self.sconv(v.contiguous()))
        return self.activation(v[..., -1:] + w)


class OutConv(nn.Module):
    """Latent steps -> output steps (fno/sfno.py:263-328): the last input frame is prepended to the latent series,
    one ``SpectralConvT`` (zero-padded in time, spectral bias, Helmholtz projection when ``out_dim == 2``) maps it
    to ``out_steps + 1`` frames, and the result is added to the last input frame."""

    def __init__(self, modes_x: int, modes_y: int, modes_t: int, delta: float = 0.1, out_dim: int = 1, diam: float = 1,
                 n_grid: int = 64, out_steps: int = None, spatial_padding: int = 0, temporal_padding: bool = True,
                 norm: str = "backward", **kwargs) -> None:
        super().__init__()
        self.size = [out_dim, out_dim, modes_x, modes_y, modes_t]
        if out_dim == 2:
            postprocess = HelmholtzProjection(n_grid=n_grid, diam=diam)
        elif out_dim == 1:
            postprocess = nn.Identity()
        self.conv = SpectralConvT(*self.size, norm=norm, delta=delta, out_steps=out_steps, bias=True,
                                  temporal_padding=temporal_padding, postprocess=postprocess)
        self.n_grid, self.norm, self.delta = n_grid, norm, delta
        self.spatial_padding, self.temporal_padding = spatial_padding, temporal_padding

    def forward(self, v, v_res, out_steps: int, **kwargs):
        v_res = v_res.unsqueeze(1).expand(-1, v.size(1), -1, -1, -1)  # "b x y t -> b d x y t"
        v = torch.cat([v_res[..., -1:], v], dim=-1)
        sp = self.spatial_padding
        if sp > 0:
            v = F.pad(v, pad=(0, 0, sp, sp, sp, sp), mode="constant")
        v = self.conv(v.contiguous(), out_steps=out_steps + 1)
        if sp > 0:
            v = v[..., sp:-sp, sp:-sp, :]
        v = v_res[..., -1:] + v[..., -out_steps:]
        return v.squeeze(1)


class FNOBase(nn.Module):
    """Bookkeeping shared by the FNO variants (fno/base.py:240-400): hyper-parameters, the latent blocks
    ``spectral_conv / mlp / w / activations`` and the latent-tensor hooks."""

    latent_tensors = {}

    def __init__(self, *, num_spectral_layers: int = 4, fft_norm="backward", activation: ActivationType = "ReLU",
                 spatial_padding: int = 0, channel_expansion: int = 4, spatial_random_feats: bool = False,
                 lift_activation: bool = False, debug=False, **kwargs):
        super().__init__()
        self.spatial_padding, self.fft_norm, self.activation = spatial_padding, fft_norm, activation
        self.spatial_random_feats, self.lift_activation = spatial_random_feats, lift_activation
        self.channel_expansion, self.debug, self.num_spectral_layers = channel_expansion, debug, num_spectral_layers

    @staticmethod
    def _set_modulelist(module, num_layers, *args):
        return nn.ModuleList([deepcopy(module(*args)) for _ in range(num_layers)])

    def _set_spectral_layers(self, num_layers: int, modes: List[int], width: int, activation: ActivationType,
                             spectral_conv, mlp, linear, channel_expansion: int = 4) -> None:
        spec = {"spectral_conv": (spectral_conv, (width, width, *modes)),
                "mlp": (mlp, (width, width, channel_expansion * width, activation)),
                "w": (linear, (width, width, 1)),
                "activations": (getattr(nn, activation), ())}
        for attr, (module, args) in spec.items():
            setattr(self, attr, self._set_modulelist(module, num_layers, *args))

    def add_latent_hook(self, layer_name: str):
        def make(name):
            def hook(model, input, output):
                self.latent_tensors[name] = output.detach()
            return hook
        module = getattr(self, layer_name)
        if hasattr(module, "__iter__"):
            for k, b in enumerate(module):
                b.register_forward_hook(make(f"{layer_name}_{k}"))
        else:
            module.register_forward_hook(make(layer_name))

    def forward(self, *args, **kwargs):
        raise NotImplementedError("Subclasses of FNO must implement the forward method")


class SFNO(FNOBase):
    """Spectral-refiner FNO for (2+1)-D fields (fno/sfno.py:460-620): (b, x, y, t_in) -> (b, x, y, out_steps).
    ``lifting_operator`` -> ``num_spectral_layers - 1`` blocks ``act(mlp(conv(v)) + w(v))`` -> ``reduction`` ->
    ``output_operator``; arbitrary input / output steps, ``latent_steps`` frames inside."""

    def __init__(self, modes_x: int, modes_y: int, modes_t: int, width: int, out_dim: int = 1, beta: float = -1e-2,
                 delta: float = 1e-1, num_spectral_layers: int = 4, fft_norm: str = "backward",
                 activation: ActivationType = "ReLU", spatial_padding: int = 0, temporal_padding: bool = True,
                 channel_expansion: int = 4, spatial_random_feats: bool = False, lift_activation: bool = True,
                 latent_steps: int = 10, output_steps: int = None, debug=False, **kwargs):
        super().__init__(num_spectral_layers=num_spectral_layers, fft_norm=fft_norm, activation=activation,
                         spatial_padding=spatial_padding, channel_expansion=channel_expansion,
                         spatial_random_feats=spatial_random_feats, lift_activation=lift_activation, debug=debug, **kwargs)
        self.modes_x, self.modes_y, self.modes_t, self.width = modes_x, modes_y, modes_t, width
        assert num_spectral_layers > 1
        self._set_spectral_layers(num_spectral_layers - 1, [modes_x, modes_y, modes_t], width, spectral_conv=SpectralConvS,
                                  mlp=PointwiseFFN, linear=nn.Conv3d, activation=activation,
                                  channel_expansion=channel_expansion)
        self.lifting_operator = LiftingOperator(width, modes_x, modes_y, modes_t, latent_steps=latent_steps, norm=fft_norm,
                                                beta=beta, activation=activation, spatial_random_feats=spatial_random_feats,
                                                channel_expansion=channel_expansion, nonlinear=lift_activation)
        self.output_operator = OutConv(modes_x, modes_y, modes_t, out_dim=out_dim, delta=delta, out_steps=output_steps,
                                       spatial_padding=spatial_padding, temporal_padding=temporal_padding, norm=fft_norm)
        self.reduction = nn.Conv3d(width, 1, kernel_size=1)
        self.out_steps = output_steps
        self.debug = debug

    @property
    def set_lifting_operator(self):
        return self.lifting_operator

    @property
    def set_output_operator(self):
        return self.output_operator

    def forward(self, v, out_steps=None):
        if out_steps is None:
            out_steps = self.out_steps if self.out_steps is not None else v.size(-1)
        v_res = v
        v = self.lifting_operator(v.unsqueeze(1))
        for conv, mlp, w, nonlinear in zip(self.spectral_conv, self.mlp, self.w, self.activations):
            v = nonlinear(mlp(conv(v.contiguous())) + w(v))
        v = self.reduction(v)
        return self.output_operator(v, v_res, out_steps=out_steps)
