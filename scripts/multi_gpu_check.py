#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU): batch-sharded trajectory with the NCCL
all-gather against the CPU oracle, on rank 0."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
world, rank, local = (int(os.environ[k]) for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
import torch_cfd_b200 as T
from _common import O, build_module, oracle_tables, rel_l2
n, per = 128, 3
ns = build_module(n, torch.float32, 1e-3, 0.1, "vorticity")
w0 = O.synthetic_vorticity_hat(n, per * world, 4, torch.float32)
out = T.get_trajectory_imex_sharded(ns, w0[rank * per:(rank + 1) * per].to(dev), 1e-3, num_steps=5, record_every_steps=2,
                                    fields=("vorticity", "stream"))
ok = True
if rank == 0:
    ref = O.trajectory(oracle_tables(n, torch.float32, 1e-3, 0.1, "vorticity"), w0, 1e-3, 5, 2)
    for k in ("vorticity", "stream"):
        e = rel_l2(out[k], ref[k])
        print(f"world={world} {k}: shape {tuple(out[k].shape)} rel-L2 vs oracle {e:.2e}")
        ok = ok and out[k].shape == ref[k].shape and e < 2e-5
    print("MULTI_GPU_CHECK", "OK" if ok else "FAILED")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
