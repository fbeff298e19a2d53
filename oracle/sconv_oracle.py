"""CPU oracle for hot path B: the FNO3d / SFNO spectral-convolution layer.

TEST INFRASTRUCTURE ONLY (same rules as oracle/ns2d_oracle.py).  Functional restatement on CPU
tensors of

* ``SpectralConv3d.forward``      fno/fno3d.py:86-116   (complex weights ``weights1..4``)
* ``SpectralConv.forward``        fno/base.py:229-237   (rfftn -> spectral_conv -> irfftn(s=out))
* ``SpectralConvS.spectral_conv`` fno/sfno.py:364-391   (real (...,2) weights, optional bias*delta)
* ``SpectralConvT.forward``       fno/sfno.py:433-457   (front zero-pad in t, out_steps resampling)

The FFTs are the pinned third-party ``torch.fft.rfftn / irfftn`` (torch 2.11.0) that the reference
calls; ``oracle/dft_naive.py`` pins their convention at small sizes.  Gradients of the oracle come
from torch autograd applied to this restatement (the reference relies on autograd as well).
Pinned bit-for-bit against reference outputs in tests/golden/sconv.npz (tests/test_oracle_cpu.py).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn.functional as F


def corner_slices(mx: int, my: int):
    """The four (x, y) corner blocks in the order the reference visits them.

    fno/fno3d.py:101-112 uses weights1..4 <-> (lo x, lo y), (hi x, lo y), (lo x, hi y), (hi, hi);
    fno/sfno.py:377-384 indexes ``weight[ix + 2*iy]`` -- the same order.
    """
    sx = [slice(0, mx), slice(-mx, None)]
    sy = [slice(0, my), slice(-my, None)]
    return [(sx[ix], sy[iy]) for iy in range(2) for ix in range(2)]


def spectral_mix(v_hat: torch.Tensor, weights: Sequence[torch.Tensor], co: int, mx: int, my: int,
                 mt: int, bias: Optional[Sequence[torch.Tensor]] = None, delta: float = 1.0):
    """out[b,o,corner,:mt] = sum_i v_hat[b,i,corner,:mt] * W[i,o,...] (+ delta*bias); zeros elsewhere.

    weights: four complex tensors (Ci, Co, mx, my, mt).  Later corners overwrite earlier ones where
    they overlap (2*mx > X), exactly as the reference's slice assignments do.
    fno/fno3d.py:92-112, fno/sfno.py:364-391.
    """
    b, ci, kx, ky, kt = v_hat.shape
    out = torch.zeros(b, co, kx, ky, kt, dtype=v_hat.dtype)
    st = slice(0, mt)
    for idx, (sx, sy) in enumerate(corner_slices(mx, my)):
        out[..., sx, sy, st] = torch.einsum("bixyz,ioxyz->boxyz", v_hat[..., sx, sy, st], weights[idx])
        if bias is not None:
            out[..., sx, sy, st] += delta * bias[idx][None, None, ...]
    return out


def spectral_conv3d(x: torch.Tensor, weights: Sequence[torch.Tensor], mx: int, my: int, mt: int):
    """fno/fno3d.py:86-116."""
    co = weights[0].shape[1]
    x_ft = torch.fft.rfftn(x, dim=[-3, -2, -1])
    out_ft = spectral_mix(x_ft, weights, co, mx, my, mt)
    return torch.fft.irfftn(out_ft, s=(x.size(-3), x.size(-2), x.size(-1)))


def spectral_conv_s(v: torch.Tensor, weight_real: Sequence[torch.Tensor], mx: int, my: int, mt: int,
                    bias_real: Optional[Sequence[torch.Tensor]] = None, delta: float = 1.0,
                    out_mesh_size: Optional[Sequence[int]] = None, norm: str = "backward"):
    """SpectralConvS.forward = SpectralConv.forward: fno/base.py:229-237 + fno/sfno.py:364-391.
    weight_real: four real tensors (Ci, Co, mx, my, mt, 2); bias_real: four (mx, my, mt, 2)."""
    mesh = list(v.shape[2:])
    out_mesh = mesh if out_mesh_size is None else list(out_mesh_size)
    w = [torch.view_as_complex(t) for t in weight_real]
    bs = None if bias_real is None else [torch.view_as_complex(t) for t in bias_real]
    v_hat = torch.fft.rfftn(v, dim=(-3, -2, -1), norm=norm)
    out_hat = spectral_mix(v_hat, w, w[0].shape[1], mx, my, mt, bs, delta)
    return torch.fft.irfftn(out_hat, s=out_mesh, dim=(-3, -2, -1), norm=norm)


def spectral_conv_t(v: torch.Tensor, weight_real, mx: int, my: int, mt: int, out_steps: int,
                    bias_real=None, delta: float = 0.1, temporal_padding: bool = False,
                    norm: str = "backward"):
    """SpectralConvT.forward with postprocess=Identity: fno/sfno.py:433-457."""
    if temporal_padding:
        t_pad = v.size(-1)
        v = F.pad(v, (t_pad, 0))
    else:
        t_pad = 0
    nx, ny, ntp = v.shape[-3:]
    w = [torch.view_as_complex(t) for t in weight_real]
    bs = None if bias_real is None else [torch.view_as_complex(t) for t in bias_real]
    v_hat = torch.fft.rfftn(v, dim=(-3, -2, -1), norm=norm)
    out_hat = spectral_mix(v_hat, w, w[0].shape[1], mx, my, mt, bs, delta)
    out = torch.fft.irfftn(out_hat, s=(nx, ny, out_steps + t_pad), dim=(-3, -2, -1), norm=norm)
    if temporal_padding:
        out = out[..., -out_steps:]
    return out
