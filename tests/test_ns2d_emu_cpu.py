"""CPU tier for hot path A: the product kernels compiled for host threads (tests/emu/libtcfd_emu.so,
-DTCFD_EMU, test infrastructure only) against the reference-generated goldens and the oracle;
the C-ABI surface of the product library; host-side logic of the nn.Module shims."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from _common import (O, ROOT, default_dtype, emu_plan, load_golden, oracle_tables, rel_l2,
                     substage_scalars)


def _golden_tables(g, dtype):
    return oracle_tables(int(g["n"]), dtype, float(g["viscosity"]), float(g["drag"]), str(g["forcing"]),
                         True, float(g["diam"]))


@pytest.mark.parametrize("name,dtype,tol", [
    ("ns2d_c1_fp64", torch.float64, 1e-12),
    ("ns2d_fp32_unforced", torch.float32, 2e-6),
    ("ns2d_fp64_velforce", torch.float64, 1e-12),
])
@pytest.mark.parametrize("small", ["1", "0"])
def test_emu_kernels_vs_reference_golden(name, dtype, tol, small, monkeypatch):
    """small = "1": grids up to 64^2 take the shared-memory-resident kernel (ONE launch per call, ns2d_small.cuh);
    "0": two launches per sub-stage everywhere below 256^2."""
    monkeypatch.setenv("TCFD_SMALL", small)
    g = load_golden(name)
    tb = _golden_tables(g, dtype)
    w0 = torch.from_numpy(g["w0_hat"]).reshape(-1, tb.n, tb.n // 2 + 1)
    plan = emu_plan(tb, w0.shape[0], dtype)
    dt = float(g["dt"])
    beta, gdt, mu = substage_scalars(dtype, dt)
    F = torch.empty_like(w0)
    plan.explicit_terms(w0, F)
    assert rel_l2(F, torch.from_numpy(g["F0"]).reshape(F.shape)) < tol
    for s in (1, 2):
        if f"w_{s}" not in g.files:
            continue
        out, dw = torch.empty_like(w0), torch.empty_like(w0)
        plan.step(w0, out, dw, s, beta, gdt, mu, 1 / (s * dt))
        assert plan.last_launch_count == (1 if (small == "1" and tb.n <= 64) else 1 + 2 * 5 * s)
        assert rel_l2(out, torch.from_numpy(g[f"w_{s}"]).reshape(out.shape)) < tol
        # dw/dt is a difference of nearly equal states: looser in relative terms
        assert rel_l2(dw, torch.from_numpy(g[f"dwdt_{s}"]).reshape(out.shape)) < tol * 1e4
    r = torch.empty_like(w0)
    w1 = torch.from_numpy(g["w_1"]).reshape(w0.shape)
    d1 = torch.from_numpy(g["dwdt_1"]).reshape(w0.shape)
    plan.residual(w1, d1, r)
    # the residual is a cancellation (w_t - F - L w ~ 0): measure its error on the scale of w_t
    rr = torch.from_numpy(g["res_1"]).reshape(w0.shape)
    assert (torch.linalg.norm(r - rr) / torch.linalg.norm(d1)).item() < tol


@pytest.mark.parametrize("n,batch,dtype,forcing,smooth", [
    (32, 3, torch.float32, "vorticity", True),
    (128, 2, torch.float64, None, True),
    (64, 5, torch.float32, "velocity", False),
    (256, 1, torch.float32, "vorticity", True),
])
def test_emu_kernels_vs_oracle(n, batch, dtype, forcing, smooth):
    tb = oracle_tables(n, dtype, 1e-3, 0.1, forcing, smooth)
    w0 = O.synthetic_vorticity_hat(n, batch, 7, dtype)
    plan = emu_plan(tb, batch, dtype)
    dt = 1e-3
    beta, gdt, mu = substage_scalars(dtype, dt)
    out = torch.empty_like(w0)
    plan.step(w0, out, None, 2, beta, gdt, mu, 1 / (2 * dt))
    ref, _ = O.forward(tb, w0, dt, 2)
    assert rel_l2(out, ref) < (1e-12 if dtype == torch.float64 else 2e-6)


def test_emu_flow_schedule_matches_two_launch_schedule(monkeypatch):
    """The persistent dataflow schedule (ns2d_flow.cuh: one launch per call, chunk-major tickets,
    W-slot workspaces re-used from chunk to chunk) against the two-launch schedule: bit-identical
    state and dw/dt, with the batch spanning several chunks (3 samples, W = 2 and W = 1)."""
    n, batch, dtype = 256, 3, torch.float32
    tb = oracle_tables(n, dtype, 1e-3, 0.1, "vorticity", True)
    w0 = O.synthetic_vorticity_hat(n, batch, 7, dtype)
    beta, gdt, mu = substage_scalars(dtype, 1e-3)
    res = {}
    for flow, W in [("0", "1"), ("1", "2"), ("1", "1")]:
        monkeypatch.setenv("TCFD_FLOW", flow)
        monkeypatch.setenv("TCFD_FLOW_W", W)
        plan = emu_plan(tb, batch, dtype)
        out, dw = torch.empty_like(w0), torch.empty_like(w0)
        plan.step(w0, out, dw, 1, beta, gdt, mu, 1e3)
        assert plan.last_launch_count == (1 if flow == "1" else 11)
        res[(flow, W)] = (out, dw)
        plan.close()
    for key in [("1", "2"), ("1", "1")]:
        assert torch.equal(res[key][0], res[("0", "1")][0]) and torch.equal(res[key][1], res[("0", "1")][1])
    ref, _ = O.forward(tb, w0, 1e-3, 1)
    assert rel_l2(res[("1", "2")][0], ref) < 2e-6


def test_emu_flow_grouped_items_512(monkeypatch):
    """The grouped-item variant of the dataflow kernel (3 double rows / 4 column quads per ticket, units
    pipelined inside the item, cross-item staging) is what 512^2 and larger grids run by default with a wide
    window; here it is forced (TCFD_FLOW_G) on a two-sample batch with a one-sample window, so that the
    slots are re-used, and checked against the oracle."""
    n, batch, dtype = 512, 2, torch.float32
    tb = oracle_tables(n, dtype, 1e-3, 0.1, "vorticity", True)
    w0 = O.synthetic_vorticity_hat(n, batch, 7, dtype)
    beta, gdt, mu = substage_scalars(dtype, 1e-3)
    monkeypatch.setenv("TCFD_FLOW", "1")
    monkeypatch.setenv("TCFD_FLOW_W", "1")
    monkeypatch.setenv("TCFD_FLOW_G", "3,4,4")
    plan = emu_plan(tb, batch, dtype)
    out, dw = torch.empty_like(w0), torch.empty_like(w0)
    plan.step(w0, out, dw, 1, beta, gdt, mu, 1e3)
    assert plan.last_launch_count == 1
    ref, dref = O.forward(tb, w0, 1e-3, 1)
    assert rel_l2(out, ref) < 2e-6
    assert (torch.linalg.norm(dw - dref) * 1e-3 / torch.linalg.norm(ref)).item() < 2e-6
    plan.close()


def test_emu_imex_crank_nicolson_is_one_fused_substage():
    """IMEXStepper(order=1.5) (torch_cfd/equations.py:176-193) == one sub-stage of the fused step
    (beta 0, gamma dt = dt, mu = dt/2): kernels against the reference-generated fixture; and the host
    logic that decides which IMEX settings may take the fused path."""
    import torch_cfd_b200 as T
    g = load_golden("ns2d_imex")
    dtype = torch.float64
    tb = _golden_tables(g, dtype)
    w0 = torch.from_numpy(g["w0_hat"])
    dt = float(g["dt"])
    with default_dtype(dtype):
        st = T.IMEXStepper(order=1.5)
        assert st.fusable(dt) and T.IMEXStepper(order=1).fusable(dt)
        assert not T.IMEXStepper(order=1, alpha=1.0).fusable(dt) and not T.IMEXStepper(order=2).fusable(dt)
        beta, gdt, mu = st.substage_scalars(dt)
    assert (beta, gdt, mu) == ([0.0], [dt], [0.5 * dt])
    plan = emu_plan(tb, w0.shape[0], dtype)
    for s in (1, 3):
        out, dw = torch.empty_like(w0), torch.empty_like(w0)
        plan.step(w0, out, dw, s, beta, gdt, mu, 1 / (s * dt))
        assert rel_l2(out, torch.from_numpy(g[f"o15_w_{s}"])) < 1e-12
        assert rel_l2(dw, torch.from_numpy(g[f"o15_dwdt_{s}"])) < 1e-8


def test_emu_batch_smaller_than_plan_and_errors():
    tb = oracle_tables(32, torch.float32)
    plan = emu_plan(tb, 4, torch.float32)
    w0 = O.synthetic_vorticity_hat(32, 4, 1, torch.float32)
    beta, gdt, mu = substage_scalars(torch.float32, 1e-3)
    full = torch.empty_like(w0)
    plan.step(w0, full, None, 1, beta, gdt, mu, 1e3)
    part = torch.empty_like(w0[:3])
    plan.step(w0[:3].contiguous(), part, None, 1, beta, gdt, mu, 1e3)
    assert torch.equal(part, full[:3])  # samples are independent: bit-identical
    with pytest.raises(RuntimeError, match="alias"):
        plan.step(w0, w0, None, 1, beta, gdt, mu, 1e3)
    with pytest.raises(ValueError):
        plan.step(w0[:, :16].contiguous(), full, None, 1, beta, gdt, mu, 1e3)
    big = torch.cat([w0, w0])
    with pytest.raises(RuntimeError, match="max_batch"):
        plan.step(big, torch.empty_like(big), None, 1, beta, gdt, mu, 1e3)


def test_unsupported_size_is_reported():
    from torch_cfd_b200 import _lib
    from _common import ensure_emu_lib
    lib = _lib.TcfdLibrary(ensure_emu_lib())
    n = 48
    z = torch.zeros(n, n // 2 + 1)
    with pytest.raises(RuntimeError, match="unsupported grid size"):
        _lib.NS2DPlan(lib, n, torch.float32, 1, torch.zeros(n), torch.zeros(n // 2 + 1), z, z, None, None)
    n = 2048  # fp64 stops at 1024: the column tile would not fit in shared memory
    z = torch.zeros(n, n // 2 + 1, dtype=torch.float64)
    with pytest.raises(RuntimeError, match="fp32 only"):
        _lib.NS2DPlan(lib, n, torch.float64, 1, torch.zeros(n, dtype=torch.float64),
                      torch.zeros(n // 2 + 1, dtype=torch.float64), z, z, None, None)


def test_product_library_exports_every_declared_symbol():
    """libtcfd.so (nvcc, sm_100a) loads without a GPU and exports all of include/tcfd.h."""
    hdr = open(os.path.join(ROOT, "include", "tcfd.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(tcfd_[a-z0-9_]+)\s*\(", hdr)))
    assert "tcfd_ns2d_step" in names and "tcfd_last_error" in names
    path = os.path.join(ROOT, "torch-cfd_b200", "libtcfd.so")
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(path)
    for nm in names:
        assert hasattr(lib, nm), nm
    lib.tcfd_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.tcfd_version()


def test_module_host_logic_matches_reference_tables():
    """Buffers, mask, forcing spectrum and substage scalars of the shim == reference goldens."""
    import torch_cfd_b200 as T
    from _common import build_module
    for name, dtype in [("ns2d_c1_fp64", torch.float64), ("ns2d_fp32_n128_nobatch", torch.float32),
                        ("ns2d_fp64_velforce", torch.float64)]:
        g = load_golden(name)
        with default_dtype(dtype):
            ns = build_module(int(g["n"]), dtype, float(g["viscosity"]), float(g["drag"]), str(g["forcing"]))
            for key in ["kx", "ky", "laplace", "linear_term", "filter"]:
                assert np.array_equal(getattr(ns, key).numpy(), g[key]), (name, key)
            tb = _golden_tables(g, dtype)
            assert torch.equal(ns.forcing_hat(), O.forcing_hat(tb))
            assert ns.solver.substage_scalars(float(g["dt"])) == substage_scalars(dtype, float(g["dt"]))
            w0 = torch.from_numpy(g["w0_hat"])
            assert np.array_equal(ns.implicit_terms(w0).numpy(), g["G0"])
            assert np.array_equal(ns.implicit_solve(w0, 0.5 * float(g["dt"])).numpy(), g["solve0"])
    gm = load_golden("mask_bounds")
    for n in [16, 32, 64, 128, 256, 512, 1024, 2048]:
        m = T.brick_wall_filter_2d(T.Grid((n, n), domain=((0, 1), (0, 1))))
        from torch_cfd_b200.spectral import brick_wall_bounds
        assert list(brick_wall_bounds(n)) + [int(m.sum())] == gm[f"n{n}"].tolist()


def test_stable_time_step_matches_reference_values():
    """stable_time_step (torch_cfd/equations.py:35-64): same signature / defaults / positional order, and
    the values the reference returns (computed with the reference itself in the authoring container)."""
    import inspect
    import math
    import torch_cfd_b200 as T
    sig = [(p.name, p.default) for p in inspect.signature(T.stable_time_step).parameters.values()]
    assert sig == [("dx", None), ("dt", None), ("max_velocity", 1.0), ("max_courant_number", 0.5),
                   ("viscosity", 1e-3), ("implicit_diffusion", True), ("ndim", 2)]
    f = T.stable_time_step
    assert f(dx=2 * math.pi / 64, dt=1e-3, viscosity=1e-3, max_velocity=7.0) == 0.001          # C1
    assert f(dx=1 / 256, dt=1.0, max_velocity=2.0) == 0.0009765625
    assert f(dx=0.01) == 0.005
    assert f(dx=1 / 2048, dt=1e-3, viscosity=1e-3, max_velocity=7.0, implicit_diffusion=False) == 3.487723214285714e-05
    assert f(dx=0.05, dt=None, max_velocity=3.0, max_courant_number=0.9, viscosity=0.5, implicit_diffusion=False,
             ndim=3) == 0.0006250000000000001
    assert f(0.1, 0.2, 4.0) == 0.0125


def test_module_refuses_cpu_and_bad_inputs():
    from _common import build_module
    with default_dtype(torch.float32):
        ns = build_module(64, torch.float32)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ns(torch.zeros(1, 64, 33, dtype=torch.complex64), 1e-3)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ns.explicit_terms(torch.zeros(64, 33, dtype=torch.complex64))


def test_drop_in_signatures_match_reference():
    """every shim takes the reference's parameters -- same names, order, kinds and defaults
    (tests/golden/signatures.json, written from the reference by make_golden.py); a shim may only APPEND
    optional parameters."""
    import inspect
    import json
    import torch_cfd_b200 as T
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "signatures.json")))

    def enc(d):
        if d is inspect._empty:
            return "<required>"
        return d if isinstance(d, (int, float, bool, str, type(None))) else repr(d)

    for name, want in ref.items():
        obj = T
        for part in name.split("."):
            obj = getattr(obj, part)
        got = [[p.name, enc(p.default), p.kind.name] for p in inspect.signature(obj).parameters.values()]
        var_kw = [g for g in got if g[2] == "VAR_KEYWORD"]
        core = [g for g in got if g[2] != "VAR_KEYWORD"]
        want_core = [w for w in want if w[2] != "VAR_KEYWORD"]
        for w, g in zip(want_core, core):
            if w[0] == "postprocess":  # upstream default is an nn.Identity() instance; None means the same here
                assert g[0] == "postprocess"
                continue
            assert w == g, (name, w, g)
        assert len(core) >= len(want_core), name
        for extra in core[len(want_core):]:
            assert extra[1] != "<required>", (name, extra)
        if any(w[2] == "VAR_KEYWORD" for w in want):
            assert var_kw, name
