"""Shared helpers of the parity tests (test infrastructure; may import oracle/)."""
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libtcfd_emu.so")
CSRC = os.path.join(ROOT, "torch-cfd_b200", "csrc")

from oracle import ns2d_oracle as O  # noqa: E402


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().cpu(), b.detach().cpu()
    return (torch.linalg.norm((a - b).reshape(-1)) / torch.linalg.norm(b.reshape(-1))).item()


def oracle_forcing(kind, n, diam, dtype):
    if kind in (None, "None"):
        return None
    if kind == "vorticity":
        return ("vorticity", O.kolmogorov_forcing_vorticity(n, diam, dtype))
    return ("velocity", O.kolmogorov_forcing_velocity(n, diam, dtype))


def oracle_tables(n, dtype, viscosity=1e-3, drag=0.0, forcing=None, smooth=True, diam=2 * torch.pi):
    return O.make_tables(n, diam, viscosity, drag, smooth, oracle_forcing(forcing, n, diam, dtype), dtype)


def substage_scalars(dtype, dt, low_storage=True):
    a, b, g = O.rk_coefficients(low_storage, dtype)
    ns = len(b)
    return ([float(b[k]) for k in range(ns)], [float(g[k] * dt) for k in range(ns)],
            [float(0.5 * dt * (a[k + 1] - a[k])) for k in range(ns)])


def build_module(n, dtype, viscosity=1e-3, drag=0.0, forcing=None, smooth=True, diam=2 * torch.pi):
    """The product nn.Module configured like the oracle tables (needs default dtype == dtype)."""
    import torch_cfd_b200 as T
    assert torch.get_default_dtype() == dtype
    grid = T.Grid(shape=(n, n), domain=((0, diam), (0, diam)))
    fn = None
    if forcing not in (None, "None"):
        fn = T.KolmogorovForcing(diam=diam, wave_number=1, grid=grid, scale=1, vorticity=(forcing == "vorticity"))
    return T.NavierStokes2DSpectral(viscosity=viscosity, grid=grid, drag=drag, smooth=smooth, forcing_fn=fn,
                                    solver=T.RK4CrankNicolsonStepper())


class default_dtype:
    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        self.prev = torch.get_default_dtype()
        torch.set_default_dtype(self.dtype)

    def __exit__(self, *a):
        torch.set_default_dtype(self.prev)


_EMU_CHECKED = False


def ensure_emu_lib():
    """Build tests/emu/libtcfd_emu.so (plain g++, the kernels compiled for host threads) if absent OR older than the
    kernel sources (one `make emu` per test process: a no-op when up to date, so a stale library is never tested)."""
    global _EMU_CHECKED
    if not _EMU_CHECKED or not os.path.exists(EMU_LIB):
        subprocess.run(["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1)), "emu"], check=True,
                       stdout=subprocess.DEVNULL)
        _EMU_CHECKED = True
    return EMU_LIB


def emu_plan(tb, batch, dtype):
    """An NS2DPlan on the host-emulation build, fed with the oracle's tables."""
    from torch_cfd_b200 import _lib
    lib = _lib.TcfdLibrary(ensure_emu_lib())
    kappa_x = (2j * torch.pi * tb.kx).imag[:, 0]
    kappa_y = (2j * torch.pi * tb.ky).imag[0, :]
    nil = -1 / O.laplacian_patched(tb.kx, tb.ky)
    return _lib.NS2DPlan(lib, tb.n, dtype, batch, kappa_x, kappa_y, nil, tb.linear_term,
                         tb.filter if tb.smooth else None, O.forcing_hat(tb))
