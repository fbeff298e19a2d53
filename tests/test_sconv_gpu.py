"""GPU parity tests of hot path B (pytest -m gpu): module shim -> autograd.Function -> ctypes ->
C ABI -> sm_100a kernels, against the reference-generated fixtures, the oracle (values and torch
autograd gradients) and, at BASELINE's C4/C5 sizes, adjointness / linearity properties."""
import pytest
import torch

from _common import load_golden, rel_l2
from _sconv_common import GOLDEN32, golden_params
from oracle import sconv_oracle as SO

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _module(kind, cfg, wr, br):
    from torch_cfd_b200.fno import SpectralConv3d, SpectralConvS, SpectralConvT
    mx, my, mt = cfg["modes"]
    if kind == "c3d":
        m = SpectralConv3d(cfg["Ci"], cfg["Co"], mx, my, mt)
        with torch.no_grad():
            for i, w in enumerate(wr, 1):
                getattr(m, f"weights{i}").copy_(torch.view_as_complex(w))
        return m.to(DEV)
    if kind == "cs":
        m = SpectralConvS(cfg["Ci"], cfg["Co"], mx, my, mt, bias=bool(cfg.get("bias")), delta=cfg.get("delta", 1),
                          norm=cfg.get("norm", "backward"))
    else:
        m = SpectralConvT(cfg["Ci"], cfg["Co"], mx, my, mt, delta=cfg.get("delta", 0.1), out_steps=cfg.get("out_steps"),
                          bias=bool(cfg.get("bias")), temporal_padding=bool(cfg.get("temporal_padding")),
                          norm=cfg.get("norm", "backward"))
    with torch.no_grad():
        for p, w in zip(m.weight, wr):
            p.copy_(w)
        if br is not None:
            for p, b in zip(m.bias, br):
                p.copy_(b)
    return m.to(DEV)


@pytest.mark.parametrize("tag,kind,cfg", GOLDEN32)
def test_modules_vs_reference_golden(tag, kind, cfg):
    g = load_golden("sconv32")
    x = torch.from_numpy(g[f"{tag}_x"]).to(DEV).requires_grad_(True)
    yref, cot = torch.from_numpy(g[f"{tag}_y"]), torch.from_numpy(g[f"{tag}_cot"])
    wr, br, gwr, gbr = golden_params(g, tag, kind, cfg)
    m = _module(kind, cfg, wr, br)
    y = m(x, out_mesh_size=cfg["out_mesh"]) if "out_mesh" in cfg else m(x)
    assert y.shape == yref.shape and y.is_cuda and y.dtype == torch.float32
    (y * cot.to(DEV)).sum().backward()
    tol = 2e-6
    assert rel_l2(y, yref) < tol
    assert rel_l2(x.grad, torch.from_numpy(g[f"{tag}_gx"])) < tol
    params = [getattr(m, f"weights{i}") for i in range(1, 5)] if kind == "c3d" else list(m.weight)
    for p, ref in zip(params, gwr):
        got = torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad
        assert rel_l2(got, ref) < tol
    if gbr is not None:
        for p, ref in zip(m.bias, gbr):
            assert rel_l2(p.grad, ref) < tol


@pytest.mark.parametrize("X,Y,T,modes", [(64, 128, 10, (12, 20, 6)), (256, 64, 16, (20, 8, 8)), (128, 128, 7, (8, 8, 4)),
                                            (64, 256, 10, (6, 28, 6)), (32, 64, 6, (4, 12, 4)), (32, 512, 4, (4, 20, 3))])
def test_modules_vs_oracle_sizes(X, Y, T, modes):
    _modules_vs_oracle(X, Y, T, modes)


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"TCFD_SCONV_PLANES": "2"}, {"TCFD_SCONV_PLANES": "1"}, {"TCFD_SCONV_XAXIS": "1"},
                                 {"TCFD_SCONV_MIX": "1"}, {"TCFD_SCONV_MIX": "2"}])
def test_kernel_generations_vs_oracle(env, monkeypatch):
    """The older kernel generations kept as fall-backs (odd sizes, unaligned views) and A/B switches: same parity bar."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    _modules_vs_oracle(64, 128, 10, (12, 20, 6))


def _modules_vs_oracle(X, Y, T, modes):
    from torch_cfd_b200.fno import SpectralConvS
    torch.manual_seed(5)
    mx, my, mt = modes
    m = SpectralConvS(3, 2, mx, my, mt, bias=True, delta=0.7)
    with torch.no_grad():
        for b in m.bias:
            b.copy_(0.2 * torch.randn_like(b))
    x = torch.randn(2, 3, X, Y, T)
    xr = x.clone().requires_grad_(True)
    wr = [w.detach().clone().requires_grad_() for w in m.weight]
    br = [b.detach().clone().requires_grad_() for b in m.bias]
    yr = SO.spectral_conv_s(xr, wr, mx, my, mt, br, 0.7)
    cot = torch.randn_like(yr)
    yr.backward(cot)
    m = m.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    y = m(xg)
    y.backward(cot.to(DEV))
    assert rel_l2(y, yr) < 1e-5          # forward 1e-5 relative
    assert rel_l2(xg.grad, xr.grad) < 1e-4  # gradients 1e-4 relative
    for p, w in zip(m.weight, wr):
        assert rel_l2(p.grad, w.grad) < 1e-4
    for p, b in zip(m.bias, br):
        assert rel_l2(p.grad, b.grad) < 1e-4


def test_c4_size_properties():
    """C4's geometry (b reduced to 4): SpectralConvT with temporal padding, 256^2, T=10, modes (20,20,8),
    width 20.  Adjointness <y, A x'> == <A^T y, x'> of forward/backward, linearity in x, and an oracle
    check of one sample."""
    from torch_cfd_b200.fno import SpectralConvT
    torch.manual_seed(7)
    m = SpectralConvT(20, 20, 20, 20, 8, out_steps=10, temporal_padding=True, bias=False).to(DEV)
    x1 = torch.randn(4, 20, 256, 256, 10, device=DEV, requires_grad=True)
    x2 = torch.randn(4, 20, 256, 256, 10, device=DEV)
    y1 = m(x1)
    cot = torch.randn_like(y1)
    (gx,) = torch.autograd.grad(y1, x1, cot)
    with torch.no_grad():
        y2 = m(x2)
        lhs = (cot * y2).double().sum().item()
        rhs = (gx * x2).double().sum().item()
        assert abs(lhs - rhs) < 1e-4 * max(abs(lhs), 1.0)
        y12 = m(x1.detach() + 2 * x2)
        assert rel_l2(y12, y1.detach() + 2 * y2) < 1e-5
        yr = SO.spectral_conv_t(x2[:1].cpu(), [w.detach().cpu() for w in m.weight], 20, 20, 8, 10, None, 0.1, True)
        assert rel_l2(y2[:1], yr) < 1e-5


def test_errors_on_gpu():
    from torch_cfd_b200.fno import SpectralConv3d
    m = SpectralConv3d(2, 2, 4, 4, 8).to(DEV)
    with pytest.raises(ValueError, match="modes_t"):
        m(torch.zeros(1, 2, 32, 32, 10, device=DEV))  # T/2+1 = 6 < 8: the reference raises as well
    with pytest.raises(TypeError):
        m(torch.zeros(1, 2, 32, 32, 16, device=DEV, dtype=torch.float64))
    with pytest.raises(ValueError):
        m(torch.zeros(1, 3, 32, 32, 16, device=DEV))


def test_fno3d_forward_matches_reference_structure():
    """C5's model on a reduced batch: FNO3d(8, 8, 5, width 20) on (2, 13, 128, 128, 10); every
    spectral layer replaced by the oracle must give the same output (1e-5), state_dict keys as upstream."""
    from torch_cfd_b200.fno import FNO3d
    torch.manual_seed(11)
    m = FNO3d(8, 8, 5, 20, input_channel=10).to(DEV).eval()
    keys = set(m.state_dict().keys())
    assert {"p.weight", "spectral_conv.0.weights1", "spectral_conv.3.weights4", "mlp.0.mlp1.weight", "w.3.bias",
            "q.mlp2.weight"} <= keys
    x = torch.randn(2, 13, 128, 128, 10, device=DEV)
    with torch.no_grad():
        y, aux = m(x)
        assert aux is None and y.shape == (2, 128, 128, 10)
        # same network on the CPU (the reference's own fp32 path: torch's GPU convolutions default to
        # TF32 and are NOT the yardstick) with the spectral layers evaluated by the oracle
        mc = FNO3d(8, 8, 5, 20, input_channel=10).eval()
        mc.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()})
        h = mc.p(x.cpu())
        for conv, mlp, w, act in zip(mc.spectral_conv, mc.mlp, mc.w, mc.activation):
            ws = [getattr(conv, f"weights{i}").detach() for i in range(1, 5)]
            h = act(mlp(SO.spectral_conv3d(h, ws, 8, 8, 5)) + w(h))
        yr = mc.q(h).squeeze(1)
        assert rel_l2(y, yr) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("width,last_act,padding", [(20, False, 0), (10, True, 0), (12, False, 2)])
def test_fno3d_fused_inference_matches_torch_op_path(width, last_act, padding):
    """FNO3d forward (fno/fno3d.py:205-236): the inference path with the fused pointwise glue
    (tcfd_fno_pointwise_linear / layer_glue / project) against the same module evaluated with the
    reference's torch ops (TF32 off) around the same spectral convolutions."""
    from torch_cfd_b200.fno import FNO3d
    torch.manual_seed(0)
    m = FNO3d(4, 4, 3, width, input_channel=10, last_activation=last_act, padding=padding).to("cuda:0").eval()
    n = 32 - 2 * padding  # the padded grid must be a power of two for the spectral convolution
    x = torch.randn(3, 13, n, n, 10, device="cuda:0")
    with torch.no_grad():
        y_fused, _ = m(x)
    assert y_fused.shape == (3, n, n, 10)
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False  # exact fp32 torch ops
    try:
        with torch.enable_grad():  # the torch-op path (what autograd uses)
            y_ops, _ = m(x)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    y_ops = y_ops.detach()
    err = (torch.linalg.norm(y_fused - y_ops) / torch.linalg.norm(y_ops)).item()
    assert err < 1e-5, err


@pytest.mark.parametrize("C,shape,act", [(20, (2, 16, 16, 10), True), (10, (3, 7, 9, 11), False), (32, (2, 16, 16, 10), True),
                                         (2, (1, 4, 8, 8), True), (14, (2, 8, 8, 9), False), (24, (1, 8, 16, 3), True),
                                         (7, (1, 5, 5, 5), True)])
def test_layer_glue_tensor_core_path_vs_reference_ops(C, shape, act, monkeypatch):
    """tcfd_fno_layer_glue (fno/fno3d.py:223-230): the tcgen05 path (3xTF32 products, accumulators in TMEM; even C)
    and the CUDA-core path against the reference's torch ops in exact fp32; full and ragged 128-point tiles."""
    import ctypes
    import torch.nn as nn
    from torch_cfd_b200 import _lib
    lib = _lib.load_library()
    torch.manual_seed(1)
    mlp1, mlp2, w = nn.Conv3d(C, C, 1), nn.Conv3d(C, C, 1), nn.Conv3d(C, C, 1)
    c, x = torch.randn(shape[0], C, *shape[1:]), torch.randn(shape[0], C, *shape[1:])
    with torch.no_grad():
        ref = mlp2(nn.functional.gelu(mlp1(c))) + w(x)
        ref = nn.functional.gelu(ref) if act else ref
    hw = _lib.fno_glue_host_weights(mlp1.weight, mlp1.bias, mlp2.weight, mlp2.bias, w.weight, w.bias)
    lib.c.tcfd_fno_layer_glue_path.restype = ctypes.c_int
    y = _lib.fno_layer_glue(lib, c.to(DEV), x.to(DEV), hw, act)
    assert lib.c.tcfd_fno_layer_glue_path() == (1 if C % 2 == 0 else 0)
    assert rel_l2(y, ref) < 3e-6           # (1e-5 is the bar of SURVEY 8d)
    monkeypatch.setenv("TCFD_GLUE_TC", "0")
    y0 = _lib.fno_layer_glue(lib, c.to(DEV), x.to(DEV), hw, act)
    assert lib.c.tcfd_fno_layer_glue_path() == 0
    assert rel_l2(y0, ref) < 1e-6
