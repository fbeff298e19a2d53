"""CPU tier for the FNO3d layer glue (csrc/fno_glue.cu compiled for the host, test infrastructure only):
each fused pointwise stage against the reference's own torch ops (nn.Conv3d kernel 1 + nn.GELU,
fno/fno3d.py:119-130, :214-235) on the same weights."""
import pytest
import torch
import torch.nn as nn

from _common import ensure_emu_lib


def _lib():
    from torch_cfd_b200 import _lib
    return _lib, _lib.TcfdLibrary(ensure_emu_lib())


def _rel(a, b):
    return (torch.linalg.norm(a - b) / torch.linalg.norm(b)).item()


@pytest.mark.parametrize("Ci,Co,shape", [(13, 20, (2, 6, 5, 4)), (13, 10, (1, 3, 3, 3)), (5, 32, (1, 4, 4, 2))])
def test_pointwise_linear_matches_conv3d(Ci, Co, shape):
    L, lib = _lib()
    torch.manual_seed(0)
    conv = nn.Conv3d(Ci, Co, 1)
    x = torch.randn(shape[0], Ci, *shape[1:])
    with torch.no_grad():
        ref = conv(x)
    y = L.fno_pointwise_linear(lib, x, conv.weight.detach(), conv.bias.detach())
    assert y.shape == ref.shape and _rel(y, ref) < 2e-6


@pytest.mark.parametrize("C,shape,act", [(20, (2, 4, 6, 10), True), (10, (1, 3, 5, 7), False), (32, (1, 2, 4, 4), True),
                                         (7, (1, 3, 3, 3), True)])
def test_layer_glue_matches_reference_ops(C, shape, act):
    L, lib = _lib()
    torch.manual_seed(1)
    mlp1, mlp2, w = nn.Conv3d(C, C, 1), nn.Conv3d(C, C, 1), nn.Conv3d(C, C, 1)
    gelu = nn.GELU()
    c, x = torch.randn(shape[0], C, *shape[1:]), torch.randn(shape[0], C, *shape[1:])
    with torch.no_grad():
        ref = mlp2(gelu(mlp1(c))) + w(x)
        if act:
            ref = gelu(ref)
    hw = L.fno_glue_host_weights(mlp1.weight, mlp1.bias, mlp2.weight, mlp2.bias, w.weight, w.bias)
    y = L.fno_layer_glue(lib, c, x, hw, act)
    assert _rel(y, ref) < 2e-6


@pytest.mark.parametrize("C,M,shape,act", [(20, 128, (2, 4, 4, 10), False), (10, 32, (1, 3, 3, 5), True)])
def test_project_matches_reference_ops(C, M, shape, act):
    L, lib = _lib()
    torch.manual_seed(2)
    mlp1, mlp2 = nn.Conv3d(C, M, 1), nn.Conv3d(M, 1, 1)
    x = torch.randn(shape[0], C, *shape[1:])
    with torch.no_grad():
        h = mlp1(x)
        ref = mlp2(nn.GELU()(h) if act else h)
    y = L.fno_project(lib, x, mlp1.weight.detach(), mlp1.bias.detach(), mlp2.weight.detach(), mlp2.bias.detach(), act)
    assert y.shape == ref.shape and _rel(y, ref) < 2e-6


def test_argument_errors():
    L, lib = _lib()
    x = torch.randn(1, 40, 2, 2, 2)
    with pytest.raises(RuntimeError, match="at most 32"):
        L.fno_pointwise_linear(lib, x, torch.randn(40, 40), None)
    with pytest.raises(ValueError):
        L.fno_layer_glue(lib, torch.randn(1, 4, 2, 2, 2), torch.randn(1, 4, 2, 2, 3),
                         (torch.randn(4, 4), None, torch.randn(4, 4), None, torch.randn(4, 4), None), True)
