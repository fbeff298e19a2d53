#!/usr/bin/env python
"""One process driving two GPUs through the module API (per-device plans, per-device kernel attributes):
the same state stepped on cuda:0 and cuda:1 must agree bit for bit, for the NS step and the spectral conv."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch_cfd_b200 as T  # noqa: E402
from torch_cfd_b200.fno import FNO3d  # noqa: E402
from bench import make_state, VISC, DRAG, DT  # noqa: E402

assert torch.cuda.device_count() >= 2, "needs two visible GPUs"
n, B = 512, 16
diam = 2 * torch.pi
grid = T.Grid(shape=(n, n), domain=((0, diam), (0, diam)))
ns = T.NavierStokes2DSpectral(viscosity=VISC, grid=grid, drag=DRAG, smooth=True,
                              forcing_fn=T.KolmogorovForcing(diam=diam, wave_number=1, grid=grid, scale=1, vorticity=True),
                              solver=T.RK4CrankNicolsonStepper())
w0 = make_state(n, B, torch.float32, 0)
outs = []
for d in (0, 1, 0):
    w, dw = ns(w0.to(f"cuda:{d}"), DT, steps=3)
    f = ns.explicit_terms(w0.to(f"cuda:{d}"))
    outs.append((w.cpu(), dw.cpu(), f.cpu()))
assert all(torch.equal(a, b) for a, b in zip(outs[0], outs[1])) and all(torch.equal(a, b) for a, b in zip(outs[0], outs[2]))
torch.manual_seed(0)
m = FNO3d(4, 4, 3, 20, input_channel=10).eval()
x = torch.randn(2, 13, 64, 64, 10)
ys = []
with torch.no_grad():
    for d in (0, 1):
        ys.append(m.to(f"cuda:{d}")(x.to(f"cuda:{d}"))[0].cpu())
assert torch.equal(ys[0], ys[1])
print("ONE_PROCESS_TWO_GPUS OK")
