#!/usr/bin/env python
"""BASELINE configs[0] ("C1"): Kolmogorov-forced 2-D vorticity, 64 x 64, batch 1, fp64, 100 RK4+CN steps -- on the GPU
(one C-ABI call for all 100 steps), next to the oracle port of the reference on the host cores.  Prints one JSON
line: ms for the 100 steps, steps/s, launches, the 2 S-per-step and 18 S-per-step HBM figures."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch_cfd_b200 as T
from oracle import ns2d_oracle as O

torch.set_default_dtype(torch.float64)
n, steps, dev = 64, 100, torch.device("cuda", 0)
diam = 2 * torch.pi
grid = T.Grid(shape=(n, n), domain=((0, diam), (0, diam)))
forcing = T.KolmogorovForcing(diam=diam, wave_number=1, grid=grid, scale=1, vorticity=True)
ns = T.NavierStokes2DSpectral(viscosity=1e-3, grid=grid, drag=0.1, smooth=True, forcing_fn=forcing,
                              solver=T.RK4CrankNicolsonStepper())
w0 = O.synthetic_vorticity_hat(n, 1, 0, torch.float64)
w = w0.to(dev)
for _ in range(3):
    ns(w, 1e-3, steps=steps)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    out, _ = ns(w, 1e-3, steps=steps)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
launches = ns._plans[0].last_launch_count
tb = O.make_tables(n, diam, 1e-3, 0.1, True, ("vorticity", O.kolmogorov_forcing_vorticity(n, diam, torch.float64)), torch.float64)
torch.set_num_threads(os.cpu_count() or 1)
O.forward(tb, w0, 1e-3, 5)
t0 = time.perf_counter()
wr, _ = O.forward(tb, w0, 1e-3, steps)
cpu_ms = (time.perf_counter() - t0) * 1e3
err = (torch.linalg.norm(out.cpu() - wr) / torch.linalg.norm(wr)).item()
S = n * (n // 2 + 1) * 16
print(json.dumps({"config": "C1: 64x64, batch 1, fp64, forced, 100 steps", "ms_100_steps": ms, "steps_per_s": steps / ms * 1e3,
                  "launches_per_call": launches, "us_per_launch": ms * 1e3 / launches, "rel_l2_vs_oracle": err,
                  "GBps_18S": 18 * S * steps / ms / 1e6, "GBps_2S": 2 * S * steps / ms / 1e6,
                  "cpu_oracle_ms_100_steps": cpu_ms, "cpu_cores": os.cpu_count()}))
