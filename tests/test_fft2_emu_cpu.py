"""CPU tier: the stand-alone 2-D transforms and the bilinear resampling (csrc/fft2d.cu) compiled for the
host (tests/emu/libtcfd_emu.so) against torch.fft / F.interpolate -- the operations the reference's
data-generation scripts apply to a recorded trajectory (fno/data_gen/data_gen_Kolmogorov2d.py:178-188)."""
import pytest
import torch
import torch.nn.functional as F

from _common import ensure_emu_lib, rel_l2


def _lib():
    from torch_cfd_b200 import _lib
    return _lib, _lib.TcfdLibrary(ensure_emu_lib())


@pytest.mark.parametrize("n,dtype,tol", [(32, torch.float32, 2e-6), (64, torch.float64, 1e-14), (128, torch.float32, 2e-6),
                                         (256, torch.float64, 1e-14)])
def test_emu_rfft2_irfft2_vs_torch(n, dtype, tol):
    L, lib = _lib()
    plan = L.FFT2Plan(lib, n, dtype)
    g = torch.Generator().manual_seed(n)
    x = torch.randn(3, n, n, generator=g, dtype=dtype)
    xh = plan.rfft2(x)
    ref = torch.fft.rfft2(x)
    assert xh.shape == ref.shape and xh.dtype == ref.dtype
    assert rel_l2(xh, ref) < tol
    # a NON-Hermitian spectrum: C2R semantics (imaginary parts of the ky = 0, n/2 bins dropped after the kx pass)
    cd = ref.dtype
    yh = torch.randn(2, 2, n, n // 2 + 1, generator=g, dtype=dtype).to(cd) + 1j * torch.randn(2, 2, n, n // 2 + 1, generator=g, dtype=dtype).to(cd)
    y = plan.irfft2(yh)
    yr = torch.fft.irfft2(yh)
    assert y.shape == yr.shape and y.dtype == yr.dtype
    assert rel_l2(y, yr) < tol
    assert rel_l2(plan.irfft2(xh), x) < 2 * tol


@pytest.mark.parametrize("n_in,n_out,din,dout", [(64, 32, torch.float32, torch.float32), (64, 16, torch.float64, torch.float32),
                                                 (128, 48, torch.float32, torch.float32), (32, 64, torch.float64, torch.float64)])
def test_emu_bilinear_vs_torch(n_in, n_out, din, dout):
    L, lib = _lib()
    x = torch.randn(2, 3, n_in, n_in, generator=torch.Generator().manual_seed(1), dtype=din)
    y = L.resample_bilinear(lib, x, n_out, dout)
    ref = F.interpolate(x.to(dout), size=(n_out, n_out), mode="bilinear")
    assert y.shape == ref.shape and y.dtype == ref.dtype
    assert (y - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()
