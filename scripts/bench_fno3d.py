#!/usr/bin/env python
"""Secondary benchmark (BASELINE.json configs[4], SURVEY 8d "C5"): FNO3d full forward (4 spectral
layers) on synthetic (b=128, 13, 128, 128, 10) fp32, eval + no_grad, batch-sharded over N GPUs
(one process per GPU, no collective).  FNO3d(8, 8, 5, width=20) -- modes3 clipped to 5 because T=10
allows at most 6.  Prints one JSON line on rank 0: samples/s over all GPUs (max over ranks)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128, help="GLOBAL batch")
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
from torch_cfd_b200.fno import FNO3d
torch.manual_seed(0)
m = FNO3d(8, 8, 5, 20, input_channel=10).to(dev).eval()
b = a.batch // world
x = torch.randn(b, 13, 128, 128, 10, device=dev)
with torch.no_grad():
    for _ in range(3):
        m(x)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        y, _ = m(x)
    e1.record()
    torch.cuda.synchronize()
    # the spectral layers alone
    h = m.p(x)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(a.iters):
        for conv in m.spectral_conv:
            conv(h)
    f1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
ms_sc = f0.elapsed_time(f1) / a.iters
if world > 1:
    t = torch.tensor([ms, ms_sc], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_sc = t.tolist()
    dist.destroy_process_group()
if rank == 0:
    print(json.dumps({"case": "FNO3d(8,8,5,width=20) forward, (128,13,128,128,10) fp32", "n_gpus": world,
                      "global_batch": a.batch, "ms_forward": ms, "samples_per_s": a.batch / ms * 1e3,
                      "ms_spectral_layers": ms_sc, "spectral_share": ms_sc / ms}))
