"""GPU parity tests of the SFNO side of hot path B (pytest -m gpu): change of mesh in SpectralConv.forward,
SpectralConvT with the Helmholtz post-process, and the SFNO model shim loaded with the REFERENCE's state_dict --
all against fixtures generated from the unmodified reference (tests/golden/make_golden.py sfno -> sfno.npz) --
plus the shape cases of the reference's own test file (fno/sfno_pytest.py:134-296)."""
import pytest
import torch

from _common import load_golden, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _exact_fp32():
    """The pointwise 1x1x1 convolutions are torch modules: keep them in true fp32 (cuDNN / cuBLAS default to TF32)."""
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def _check(tag, m, g, **fw):
    with torch.no_grad():
        for name, p in m.named_parameters():
            p.copy_(torch.from_numpy(g[f"{tag}_p_{name.replace('.', '_')}"]))
    m = m.to(DEV)
    x = torch.from_numpy(g[f"{tag}_x"]).to(DEV).requires_grad_(True)
    y = m(x, **fw)
    yref = torch.from_numpy(g[f"{tag}_y"])
    assert y.shape == yref.shape and y.is_cuda
    (y * torch.from_numpy(g[f"{tag}_cot"]).to(DEV)).sum().backward()
    assert rel_l2(y, yref) < 5e-6, tag
    assert rel_l2(x.grad, torch.from_numpy(g[f"{tag}_gx"])) < 1e-5, tag
    for name, p in m.named_parameters():
        ref = torch.from_numpy(g[f"{tag}_g_{name.replace('.', '_')}"])
        scale = max(1e-30, float(torch.linalg.norm(ref)))
        assert float(torch.linalg.norm(p.grad.cpu() - ref)) / scale < 1e-4, (tag, name)


def test_out_mesh_xy_vs_reference_golden():
    from torch_cfd_b200.fno import SpectralConvS
    g = load_golden("sfno")
    _check("cs_up", SpectralConvS(2, 2, 4, 4, 3), g, out_mesh_size=[64, 64, 12])
    _check("cs_down", SpectralConvS(2, 3, 6, 5, 3, norm="ortho"), g, out_mesh_size=[32, 32, 6])


def test_helmholtz_postprocess_vs_reference_golden():
    from torch_cfd_b200.fno import HelmholtzProjection, SpectralConvT
    g = load_golden("sfno")
    m = SpectralConvT(2, 2, 4, 4, 3, out_steps=6, temporal_padding=True, bias=True,
                      postprocess=HelmholtzProjection(n_grid=32, diam=1))
    _check("ct_helm", m, g)
    # the projected field is divergence free: spectral divergence of the output ~ 0
    x = torch.from_numpy(g["ct_helm_x"]).to(DEV)
    with torch.no_grad():
        y = m.to(DEV)(x)
    yh = torch.fft.fft2(y.cpu().double(), dim=(2, 3))
    k = torch.fft.fftfreq(32, d=1 / 32, dtype=torch.float64)
    div = yh[:, 0] * k[None, :, None, None] + yh[:, 1] * k[None, None, :, None]
    assert float(torch.linalg.norm(div)) < 1e-5 * float(torch.linalg.norm(yh))


def test_sfno_with_reference_state_dict():
    from torch_cfd_b200.fno import SFNO
    g = load_golden("sfno")
    model = SFNO(4, 4, 3, 8, num_spectral_layers=3, latent_steps=5)
    sd = {k[len("sfno_sd_"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sfno_sd_")}
    assert set(sd) == set(model.state_dict()), set(sd) ^ set(model.state_dict())
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    x = torch.from_numpy(g["sfno_x"]).to(DEV)
    with torch.no_grad():
        y, y9 = model(x), model(x, out_steps=9)
    assert rel_l2(y, torch.from_numpy(g["sfno_y"])) < 2e-5
    assert rel_l2(y9, torch.from_numpy(g["sfno_y9"])) < 2e-5
    # training step through the shim: gradients reach every parameter
    model.train()
    out = model(x)
    out.square().mean().backward()
    missing = [n for n, p in model.named_parameters() if p.grad is None or not torch.isfinite(p.grad).all()]
    assert not missing, missing


@pytest.mark.parametrize("mesh_size", [(64, 64, 10), (128, 128, 20)])
def test_sfno_shapes_like_reference_tests(mesh_size):
    """fno/sfno_pytest.py:251-296."""
    from torch_cfd_b200.fno import SFNO, LiftingOperator, OutConv, SpectralConvT
    model = SFNO(8, 8, 4, 16).to(DEV)
    x = torch.randn(2, *mesh_size, device=DEV)
    with torch.no_grad():
        assert model(x).shape == (2, *mesh_size)
        assert model(x, out_steps=40).shape == (2, *mesh_size[:2], 40)
        nx, ny, nt = mesh_size
        lift = LiftingOperator(width=16, modes_x=8, modes_y=8, modes_t=3, latent_steps=5).to(DEV)
        assert lift(torch.randn(2, 1, nx, ny, nt, device=DEV)).shape == (2, 16, nx, ny, 5)
        oc = OutConv(modes_x=8, modes_y=8, modes_t=3, out_dim=1).to(DEV)
        assert oc(torch.randn(2, 1, nx, ny, 10, device=DEV), torch.randn(2, nx, ny, nt, device=DEV), out_steps=40).shape == (2, nx, ny, 40)
        conv = SpectralConvT(16, 16, 8, 8, 4).to(DEV)
        for steps in (10, 20, 40):
            assert conv(torch.randn(2, 16, 64, 64, 10, device=DEV), out_steps=steps).shape == (2, 16, 64, 64, steps)
