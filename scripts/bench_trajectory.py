#!/usr/bin/env python
"""Secondary benchmark (BASELINE.json configs[2], SURVEY 8d "C3"): Kolmogorov-forced 512^2 trajectory of
2000 RK4+CN steps, 32 samples per GPU (batch 256 over 8 GPUs), fp32, recorded every 20 steps (100
snapshots), through get_trajectory_imex_sharded: steps are fused launches, the recorded fields stay on the
device and are all-gathered once at the end (NCCL) when there is more than one rank.  Prints one JSON line on
rank 0: steps/s of the whole trajectory (max over ranks), with and without the final device->host copy."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--per-gpu", type=int, default=32)
ap.add_argument("--n", type=int, default=512)
ap.add_argument("--steps", type=int, default=2000)
ap.add_argument("--every", type=int, default=20)
ap.add_argument("--fields", default="vorticity", help="comma list of vorticity,stream,vort_t,residual")
ap.add_argument("--gather-to", default="none", help="all | none | <rank>: collection of the recorded fields (get_trajectory_imex_sharded)")
ap.add_argument("--physical", type=int, default=1, help="1: irfft2 (+ bilinear subsample) on the device before the copy (data-gen post-processing)")
ap.add_argument("--subsample", type=int, default=1)
a = ap.parse_args()
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
import torch_cfd_b200 as T
from bench import make_state, VISC, DRAG, DT

n, B = a.n, a.per_gpu
diam = 2 * torch.pi
grid = T.Grid(shape=(n, n), domain=((0, diam), (0, diam)))
ns = T.NavierStokes2DSpectral(viscosity=VISC, grid=grid, drag=DRAG, smooth=True,
                              forcing_fn=T.KolmogorovForcing(diam=diam, wave_number=1, grid=grid, scale=1, vorticity=True),
                              solver=T.RK4CrankNicolsonStepper())
w0 = make_state(n, B, torch.float32, rank * B).to(dev)
fields = tuple(a.fields.split(","))
gather_to = None if a.gather_to == "none" else ("all" if a.gather_to == "all" else int(a.gather_to))
kw = dict(record_every_steps=a.every, fields=fields, gather_to=gather_to, physical=bool(a.physical), subsample=a.subsample)
T.get_trajectory_imex_sharded(ns, w0, DT, num_steps=2 * a.every, device_result=True, **kw)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
# (1) device part only: steps + recording (+ post-processing + collective), results left on the device
t0 = time.perf_counter()
out = T.get_trajectory_imex_sharded(ns, w0, DT, num_steps=a.steps, device_result=True, **kw)
torch.cuda.synchronize()
t_dev = time.perf_counter() - t0
shape = tuple(next(iter(out.values())).shape) if out else ()
finite = bool(torch.isfinite(next(iter(out.values()))[:, -1].float() if out and not next(iter(out.values())).is_complex()
                             else torch.view_as_real(next(iter(out.values()))[:, -1])).all().item()) if out else True
del out
if world > 1:
    dist.barrier()
# (2) end to end: the same call returning HOST tensors, as the data-generation scripts consume them
t0 = time.perf_counter()
host = T.get_trajectory_imex_sharded(ns, w0, DT, num_steps=a.steps, device_result=False, **kw)
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
host_bytes = sum(v.numel() * v.element_size() for v in host.values())
if world > 1:
    t = torch.tensor([t_dev, t_all], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_dev, t_all = t.tolist()
    dist.destroy_process_group()
if rank == 0:
    rec_steps = len(range(0, a.steps, a.every))
    print(json.dumps({"case": f"trajectory {a.steps} steps, {n}^2, {B} samples per GPU, every {a.every} steps, fields {fields}",
                      "n_gpus": world, "global_batch": B * world, "snapshots": rec_steps, "result_shape_rank0": shape,
                      "gather_to": a.gather_to, "physical": bool(a.physical), "subsample": a.subsample,
                      "seconds_device": t_dev, "steps_per_s_device": (a.steps - a.every + 1) / t_dev,
                      "seconds_end_to_end_host_result": t_all, "host_bytes_rank0": host_bytes, "finite": finite}))
