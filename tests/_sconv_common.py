"""Helpers shared by the CPU (emulation) and GPU tests of hot path B."""
import numpy as np
import torch

from _common import load_golden, rel_l2  # noqa: F401

# (tag, kind, ctor args) of the reference-generated fixtures in tests/golden/sconv32.npz
GOLDEN32 = [
    ("c3d_32", "c3d", dict(Ci=3, Co=2, modes=(6, 5, 4))),
    ("cs_32_bias", "cs", dict(Ci=2, Co=3, modes=(8, 16, 3), bias=True, delta=0.5)),
    ("cs_32_outT", "cs", dict(Ci=2, Co=2, modes=(4, 4, 4), norm="ortho", out_mesh=[32, 32, 11])),
    ("ct_32_pad", "ct", dict(Ci=2, Co=2, modes=(5, 4, 6), out_steps=7, temporal_padding=True, bias=True, delta=0.1)),
]


def golden_params(g, tag, kind, cfg):
    """(weights as real (...,2) tensors, biases or None) of a fixture."""
    if kind == "c3d":
        w = [torch.from_numpy(g[f"{tag}_p_weights{i}"]) for i in range(1, 5)]
        gw = [torch.from_numpy(g[f"{tag}_g_weights{i}"]) for i in range(1, 5)]
        return w, None, gw, None
    w = [torch.from_numpy(g[f"{tag}_p_weight_{i}"]) for i in range(4)]
    gw = [torch.from_numpy(g[f"{tag}_g_weight_{i}"]) for i in range(4)]
    if cfg.get("bias"):
        b = [torch.from_numpy(g[f"{tag}_p_bias_{i}"]) for i in range(4)]
        gb = [torch.from_numpy(g[f"{tag}_g_bias_{i}"]) for i in range(4)]
        return w, b, gw, gb
    return w, None, gw, None


def geometry(x_shape, y_shape, cfg):
    """(X, Y, T_in, t_pad, T_out, Ci, Co, mx, my, mt, norm) of a fixture."""
    _, Ci, X, Y, T = x_shape
    t_pad = T if cfg.get("temporal_padding") else 0
    return (X, Y, T, t_pad, y_shape[-1], Ci, y_shape[1], *cfg["modes"], cfg.get("norm", "backward"))
