#!/usr/bin/env python
"""Bring-up / timing of the FNO3d layer glue kernels: tensor-core path (default), its descriptor-stride twin
(TCFD_GLUE_SWAP=1, bring-up only) and the CUDA-core path (TCFD_GLUE_TC=0) against the reference's torch ops."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn as nn
from torch_cfd_b200 import _lib

lib = _lib.load_library()
dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def run(C, shape, act, env):
    for k in ("TCFD_GLUE_TC", "TCFD_GLUE_SWAP"):
        os.environ.pop(k, None)
    os.environ.update(env)
    torch.manual_seed(1)
    mlp1, mlp2, w = nn.Conv3d(C, C, 1), nn.Conv3d(C, C, 1), nn.Conv3d(C, C, 1)
    c, x = torch.randn(shape[0], C, *shape[1:]), torch.randn(shape[0], C, *shape[1:])
    with torch.no_grad():
        ref = mlp2(nn.functional.gelu(mlp1(c))) + w(x)
        if act:
            ref = nn.functional.gelu(ref)
    hw = _lib.fno_glue_host_weights(mlp1.weight, mlp1.bias, mlp2.weight, mlp2.bias, w.weight, w.bias)
    y = _lib.fno_layer_glue(lib, c.to(dev), x.to(dev), hw, act)
    torch.cuda.synchronize()
    return (torch.linalg.norm(y.cpu() - ref) / torch.linalg.norm(ref)).item()


for env in ({"TCFD_GLUE_TC": "0"}, {}, {"TCFD_GLUE_SWAP": "1"}):
    for C, shape, act in ((20, (2, 8, 16, 10), True), (24, (1, 8, 16, 3), False)):
        try:
            print(json.dumps({"env": env, "C": C, "shape": shape, "rel_err": run(C, shape, act, env)}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"env": env, "C": C, "error": str(e)[:200]}), flush=True)
for C, shape, act in ((7, (1, 5, 5, 5), True), (10, (3, 7, 9, 11), False), (32, (2, 16, 16, 10), True), (16, (1, 4, 8, 8), True), (2, (1, 4, 8, 8), True), (14, (2, 8, 8, 9), False), (22, (1, 16, 16, 5), True), (30, (1, 16, 8, 4), True), (20, (4, 32, 32, 10), True)):
    print(json.dumps({"env": {}, "C": C, "shape": shape, "rel_err": run(C, shape, act, {})}), flush=True)
# timing at BASELINE config C5's layer size
for env in ({"TCFD_GLUE_TC": "0"}, {}):
    for k in ("TCFD_GLUE_TC", "TCFD_GLUE_SWAP"):
        os.environ.pop(k, None)
    os.environ.update(env)
    C = 20
    c = torch.randn(128, C, 128, 128, 10, device=dev)
    x = torch.randn(128, C, 128, 128, 10, device=dev)
    hw = tuple(torch.randn(C, C) * 0.2 if i % 2 == 0 else torch.randn(C) * 0.1 for i in range(6))
    for _ in range(3):
        _lib.fno_layer_glue(lib, c, x, hw, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _lib.fno_layer_glue(lib, c, x, hw, True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"env": env, "layer_glue_ms_C5": ms, "GBps": 3 * c.numel() * 4 / ms / 1e6}), flush=True)
