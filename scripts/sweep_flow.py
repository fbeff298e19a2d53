#!/usr/bin/env python
"""Schedule sweep of hot path A on one GPU: two-launch schedule vs the persistent dataflow schedule at
several window sizes W and kernel variants (configs "flow:W[:G]" with G = "GR.GC.MINB" -> TCFD_FLOW_G, only
honoured by -DTCFD_FLOW_VARIANTS builds; TCFD_FLOW / TCFD_FLOW_W are read when a plan is created).  Prints
one line per configuration and checks that every schedule returns bit-identical states."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--multi", type=int, default=0, help="also time one call with this many steps")
    ap.add_argument("--configs", default="0:0,1:4,1:6,1:8,1:10,1:12,1:16,1:24,1:64")
    a = ap.parse_args()
    import torch_cfd_b200 as T
    from bench import make_state, VISC, DRAG, DT
    dev = torch.device("cuda", 0)
    torch.set_default_dtype(torch.float32)
    n, B = a.n, a.batch
    diam = 2 * torch.pi
    w0 = make_state(n, B, torch.float32, 0).to(dev)
    ref = None
    for cfg in a.configs.split(","):
        f = cfg.split(":")  # flow : W [: G]
        flow, W = f[0], f[1]
        G = f[2].replace(".", ",") if len(f) > 2 else ""
        os.environ["TCFD_FLOW"] = flow
        os.environ["TCFD_FLOW_W"] = W if int(W) > 0 else "64"
        if G:
            os.environ["TCFD_FLOW_G"] = G
        else:
            os.environ.pop("TCFD_FLOW_G", None)
        grid = T.Grid(shape=(n, n), domain=((0, diam), (0, diam)))
        forcing = T.KolmogorovForcing(diam=diam, wave_number=1, grid=grid, scale=1, vorticity=True)
        ns = T.NavierStokes2DSpectral(viscosity=VISC, grid=grid, drag=DRAG, smooth=True, forcing_fn=forcing,
                                      solver=T.RK4CrankNicolsonStepper())
        w = w0.clone()
        for _ in range(3):
            w, _ = ns(w, DT, steps=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            w, dw = ns(w, DT, steps=1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        rec = {"flow": int(flow), "W": int(W), "G": G, "n": n, "batch": B, "steps_per_s": a.steps / (ms * 1e-3),
               "launches_per_step": ns._plans[0].last_launch_count}
        if a.multi:
            w2, _ = ns(w0, DT, steps=a.multi)  # warm
            torch.cuda.synchronize()
            e0.record()
            w2, _ = ns(w0, DT, steps=a.multi)
            e1.record()
            torch.cuda.synchronize()
            rec["multi_steps_per_s"] = a.multi / (e0.elapsed_time(e1) * 1e-3)
        if ref is None:
            ref = (w.clone(), dw.clone())
            rec["bit_equal_to_first"] = True
        else:
            rec["bit_equal_to_first"] = bool(torch.equal(w, ref[0]) and torch.equal(dw, ref[1]))
        rec["finite"] = bool(torch.isfinite(torch.view_as_real(w)).all().item())
        if os.environ.get("TCFD_FLOW_PROF"):  # region cycle attribution (profiling kernel variant only)
            import ctypes
            plan = ns._plans[0]
            out = (ctypes.c_ulonglong * 16)()
            fn = plan.lib.c.tcfd_ns2d_flow_profile
            fn.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong)]
            if fn(plan._h, out) == 0:
                tot = float(sum(out)) or 1.0
                names = ["control", "dep_wait", "stage_wait", "fft", "cols_math", "rows_unit_head", "rows_fields", "item_barrier",
                         "rows_prologue_loads", "rows_update", "rows_post_update"]
                rec["cycles_share"] = {nm: round(out[i] / tot, 4) for i, nm in enumerate(names)}
        print(json.dumps(rec), flush=True)
        ns.invalidate_plan()
        del ns
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
