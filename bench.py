#!/usr/bin/env python
"""Benchmark of hot path A: RK4+CN pseudo-spectral vorticity steps per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--n 512] [--batch 64] [--dtype fp32|fp64]

Workload (BASELINE.json north_star target, SURVEY.md 8d row "T"): Kolmogorov-forced 2-D vorticity,
512 x 512 grid, batch 64 PER GPU, fp32, nu = 1e-3, drag 0.1, dt = 1e-3, 2/3 de-aliasing,
Carpenter-Kennedy RK4 + Crank-Nicolson.  One "step" = one full RK4 step (5 substages) of the
whole batch.  Weak scaling: every rank steps its own 64 samples, no data-path collective.

Prints ONE JSON line (rank 0).  `value` = steps/s with the state resident in HBM (CUDA events,
max over ranks); `e2e` = the same through the host-buffer API call (pinned H2D of the state,
step, D2H of both results inside the timed region); `roofline` = algorithmic bytes of the step
(18 * S * B, S = one half spectrum) over its device time against the measured HBM copy
bandwidth; `cpu_baseline` = the oracle port of the reference's CPU path on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

VISC, DRAG, DT = 1e-3, 0.1, 1e-3
# ONE unit string for both arms (the driver pairs the arms on metric / unit / direction); what a step
# is -- batch, grid, per-GPU -- lives in `config`
UNIT = "steps/s"


def workload_name(n, batch, dtype):
    """The same string in both arms."""
    return f"Kolmogorov-forced 2D vorticity RK4+CN, {n}x{n}, batch {batch} per GPU, {dtype}"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ns2d", choices=["ns2d", "sconv_c4", "fno3d_c5"],
                    help="ns2d: the headline RK4+CN step (default); sconv_c4 / fno3d_c5: BASELINE configs[3] / [4] of hot path B")
    ap.add_argument("--width", type=int, default=20, help="path B: channel width")
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU (ns2d: 64, sconv_c4: 32); fno3d_c5: GLOBAL batch (128)")
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=2)
    a = ap.parse_args()
    if a.batch is None:
        a.batch = {"ns2d": 64, "sconv_c4": 32, "fno3d_c5": 128}[a.workload]
    return a


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_state(n, batch, dtype, first_index):
    """Seeded synthetic periodic-box vorticity spectra (CPU generator => device independent).
    4 seeded fields per rank, rescaled per sample so all samples differ (cheap for large batches)."""
    from oracle import ns2d_oracle as O
    base = O.synthetic_vorticity_hat(n, 4, 1000, dtype, first_index=0)
    w = torch.empty(batch, n, n // 2 + 1, dtype=base.dtype)
    for i in range(batch):
        w[i] = base[(first_index + i) % 4] * (1.0 + 0.002 * ((first_index + i) % 97))
    return w


REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def reference_kind():
    """"reference": the UNMODIFIED reference modules copied by oracle/make_ref.sh are present (oracle/_ref, git-ignored,
    shipped by gpurun) and are what the CPU legs time; "port": the oracle restatement (pinned bit-exactly to the
    reference by tests/test_oracle_cpu.py)."""
    if os.path.isdir(os.path.join(REF_DIR, "torch_cfd")) and os.environ.get("TCFD_BENCH_PORT") != "1":
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        sys.dont_write_bytecode = True
        return "reference"
    return "port"


def cpu_reference_steps_per_s(n, batch, dtype, steps, threads):
    """The reference's CPU path on the host cores; 1 warm-up step + `steps` timed.  With oracle/_ref present this is
    the reference's own NavierStokes2DSpectral + RK4CrankNicolsonStepper (torch_cfd/equations.py), else the oracle port."""
    from oracle import ns2d_oracle as O
    torch.set_num_threads(threads)
    diam = 2 * torch.pi
    w = make_state(n, batch, dtype, 0)
    if reference_kind() == "reference":
        prev = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        try:
            from torch_cfd.grids import Grid
            from torch_cfd.equations import NavierStokes2DSpectral, RK4CrankNicolsonStepper
            from torch_cfd.forcings import KolmogorovForcing
            grid = Grid(shape=(n, n), domain=((0, diam), (0, diam)))
            ns = NavierStokes2DSpectral(viscosity=VISC, grid=grid, drag=DRAG, smooth=True,
                                        forcing_fn=KolmogorovForcing(diam=diam, wave_number=1, grid=grid, scale=1, vorticity=True),
                                        solver=RK4CrankNicolsonStepper())
            with torch.no_grad():
                w, _ = ns(w, DT, steps=1)
                t0 = time.perf_counter()
                w, _ = ns(w, DT, steps=steps)
                el = time.perf_counter() - t0
        finally:
            torch.set_default_dtype(prev)
        return steps / el, el
    tb = O.make_tables(n, diam, VISC, DRAG, True, ("vorticity", O.kolmogorov_forcing_vorticity(n, diam, dtype)), dtype)
    with torch.no_grad():
        w, _ = O.forward(tb, w, DT, 1)
        t0 = time.perf_counter()
        w, _ = O.forward(tb, w, DT, steps)
        el = time.perf_counter() - t0
    return steps / el, el


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dtype = torch.float32 if a.dtype == "fp32" else torch.float64
    threads = os.cpu_count() or 1
    k = max(1, a.steps)
    # bounded: the run must end within minutes -> cap the timed steps by a ~60 s budget
    per, el1 = cpu_reference_steps_per_s(a.n, a.batch, dtype, 1, threads)
    k = max(1, min(k, int(60.0 * per)))
    v, el = cpu_reference_steps_per_s(a.n, a.batch, dtype, k, threads)
    kind = reference_kind()
    sample = (f"{k} full steps of the {a.batch} x {a.n}^2 batch after 1 warm-up step, "
              f"{'reference modules (oracle/_ref)' if kind == 'reference' else 'oracle port'}, torch CPU, {threads} threads")
    unit = UNIT
    _emit(json.dumps({
        "impl": "reference", "metric": "rk4_spectral_steps_per_sec", "value": v, "unit": unit,
        "n_gpus": a.gpus, "steps": k, "warmup": 1, "ms_per_step": 1e3 / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if a.dtype == "fp32" else "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(a.n, a.batch, a.dtype),
                   "step": f"one RK4+CN step (5 substages) of {a.batch} x {a.n}^2 samples",
                   "note": "CPU arm: rank 0 only, one batch regardless of --gpus"},
        "cpu_baseline": {"value": v, "unit": unit, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(a):
    import torch.distributed as dist
    import torch_cfd_b200 as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float32 if a.dtype == "fp32" else torch.float64
    torch.set_default_dtype(dtype)
    n, B, K, W = a.n, a.batch, a.steps, max(3, a.warmup)
    diam = 2 * torch.pi
    grid = T.Grid(shape=(n, n), domain=((0, diam), (0, diam)))
    forcing = T.KolmogorovForcing(diam=diam, wave_number=1, grid=grid, scale=1, vorticity=True)
    ns = T.NavierStokes2DSpectral(viscosity=VISC, grid=grid, drag=DRAG, smooth=True, forcing_fn=forcing,
                                  solver=T.RK4CrankNicolsonStepper())
    w_host = make_state(n, B, dtype, rank * B).pin_memory()
    w = w_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return x

    # ---------------- device-resident: K steps, one step per API call
    for _ in range(W):
        w, _ = ns(w, DT, steps=1)
    plan = ns._plans[local]
    sampler = ClockSampler(local)
    launches = 0
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        w, _ = ns(w, DT, steps=1)
        launches += plan.last_launch_count
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    assert torch.isfinite(torch.view_as_real(w)).all().item(), "state blew up"
    value = world * K / (ms * 1e-3)

    # ---------------- per-kernel device times (separate instrumented pass, same workload): every
    # launch bracketed by CUDA events on the launching stream (tcfd_ns2d_step_timed).  Under the
    # dataflow schedule a call is ONE launch, so single-step calls are timed (the bench's own call shape).
    flow = plan.dataflow
    if flow:
        acc = None
        reps = 10
        for _ in range(reps):
            k1 = plan.kernel_times(w, DT, ns.solver, steps=1)
            if acc is None:
                acc = k1
            else:
                for key, v in k1.items():
                    if isinstance(v, dict):
                        acc[key]["launches"] += v["launches"]
                        acc[key]["ms_total"] += v["ms_total"]
        for v in acc.values():
            if isinstance(v, dict):
                v["us_per_launch"] = 1e3 * v["ms_total"] / v["launches"] if v["launches"] else None
        acc["steps"] = reps
        kt = acc
    else:
        kt = plan.kernel_times(w, DT, ns.solver, steps=3)

    # ---------------- end to end: host (pinned) buffers through the public host API
    e2e = None
    if not a.no_e2e:
        Ke = max(3, min(K, 20))
        bufs = [torch.empty_like(w_host).pin_memory() for _ in range(3)]  # state ping-pong + dw/dt
        cur, nxt = w_host, bufs[0]
        for _ in range(3):
            ns.forward_host(cur, DT, steps=1, out=nxt, dvdt_out=bufs[2])
        barrier()
        t0 = time.perf_counter()
        for i in range(Ke):
            ns.forward_host(cur, DT, steps=1, out=nxt, dvdt_out=bufs[2])
            cur, nxt = nxt, (bufs[1] if nxt is bufs[0] else bufs[0])
        barrier()
        el = max_over_ranks(time.perf_counter() - t0)
        sb = w_host.numel() * w_host.element_size()
        e2e = {"value": world * Ke / el, "unit": None, "h2d_bytes_per_step": sb, "d2h_bytes_per_step": 2 * sb,
               "steps": Ke}

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    es = 4 if dtype == torch.float32 else 8
    S = n * (n // 2 + 1) * 2 * es
    alg_bytes_step = 18 * S * B
    peak, peak_src = peaks()
    achieved = alg_bytes_step * K / (ms * 1e-3) / 1e9  # per GPU
    if flow:
        # dominant (only) kernel = the persistent dataflow launch of one step: it moves the step's whole
        # algorithmic traffic, 18 S per sample (DESIGN.md "Roofline accounting")
        dom_key, dom_name = "flow_call", "ns2d_flow_kernel (one persistent launch = one RK4+CN step of the batch)"
        dom_alg = 18.0 * S * B
        dom_note = ("one launch per step: prologue + 5 x (cols, rows) phases as ticketed work items with "
                    "per-sample dependency counters; H / advt are intermediate traffic, not algorithmic bytes")
    else:
        # dominant kernel = the substage rows kernel (forward y-FFT + RK/CN update + inverse y-FFTs): it is
        # the launch that moves the substage's algorithmic bytes (reads w, h; writes w, h: 4 S per sample,
        # 3 S in the first substage of a step) -- DESIGN.md "Roofline accounting"
        dom_key, dom_name = "rows_fwd_inv", "ns2d_rows3_kernel (substage: rows fwd + update + rows inv)"
        dom_alg = 3.75 * S * B
        dom_note = ("the cols kernel of the same substage moves no algorithmic bytes (its H/advt traffic is "
                    "intermediate); the honest whole-step figure is roofline_step")
    dom = kt.get(dom_key) or {}
    dom_us = dom.get("us_per_launch")
    dom_achieved = dom_alg / (dom_us * 1e-6) / 1e9 if dom_us else None
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tj["workload"] == {"n": n, "batch": B, "dtype": a.dtype}:
            traffic = tj["dram_bytes_per_launch"][dom_key]
    except (OSError, KeyError, ValueError):
        pass
    step_us = sum(v["ms_total"] for v in kt.values() if isinstance(v, dict)) * 1e3 / kt["steps"]
    unit = UNIT
    if e2e:
        e2e["unit"] = unit
    out = {
        "metric": "rk4_spectral_steps_per_sec", "value": value, "unit": unit, "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if dtype == torch.float32 else "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, B, a.dtype),
                   "step": f"one RK4+CN step (5 substages) of {B} x {n}^2 samples on every GPU",
                   "aggregate": "value and e2e.value are whole-job: steps/s summed over the GPUs (weak scaling, "
                                "every GPU steps its own batch); the reference arm steps ONE batch on the host cores",
                   "global_batch": B * world, "viscosity": VISC, "drag": DRAG, "dt": DT,
                   "l2": f"working set (state w+h {2 * S * B / 1e6:.0f} MB + workspace) exceeds the 126 MB L2; no flush",
                   "schedule": "dataflow (1 launch per call)" if flow else "two launches per substage",
                   "sample_steps_per_s": value * B, "cell_steps_per_s": value * B * n * n},
        "clocks": clocks, "gpu_launches": launches, "e2e": e2e,
        "roofline": {"bound": "hbm", "kernel": dom_name,
                     "achieved": dom_achieved, "peak": peak, "unit": "GB/s",
                     "frac": dom_achieved / peak if dom_achieved else None, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_alg,
                     "us_per_launch": dom_us, "share_of_step": (dom["ms_total"] * 1e3 / kt["steps"]) / step_us if dom_us else None,
                     "note": dom_note},
        "roofline_step": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                          "algorithmic_bytes_per_step": alg_bytes_step,
                          "note": "18 S B bytes per step over the device time of ALL launches of the step"},
        "kernels": kt,
    }
    if not a.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        v, el = cpu_reference_steps_per_s(n, B, dtype, a.cpu_steps, threads)
        kind = reference_kind()
        out["cpu_baseline"] = {"value": v, "unit": unit, "cores": threads, "kind": kind,
                               "sample": f"{a.cpu_steps} full steps of the {B} x {n}^2 batch after 1 warm-up, "
                                         f"{'reference modules (oracle/_ref)' if kind == 'reference' else 'oracle port of the reference'} "
                                         f"(torch CPU), {el:.1f} s"}
    _emit(json.dumps(out))


# ================================================================================================
# hot path B workloads (BASELINE.json configs[3] "C4" and configs[4] "C5"), same JSON schema
def _dist_setup():
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return x

    def finish():
        if world > 1:
            dist.destroy_process_group()
    return world, rank, local, dev, barrier, max_over_ranks, finish


def _timed(fn, K, W, barrier, max_over_ranks, local):
    for _ in range(W):
        fn()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        fn()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    return ms, sampler.stop()


C4 = dict(X=256, Y=256, T=10, modes=(20, 20, 8))  # SpectralConvT(temporal_padding=True): 10 -> 20 samples, 11 t-frequencies


def sconv_workload_name(b, C):
    return (f"SFNO spectral conv forward+backward, SpectralConvT(temporal_padding) x = ({b}, {C}, 256, 256, 10) per GPU, "
            f"modes (20, 20, 8), fp32")


def sconv_alg_bytes(b, C):
    act = b * C * C4["X"] * C4["Y"] * C4["T"] * 4
    wbytes = 4 * C * C * 20 * 20 * 8 * 8
    return 2 * act + wbytes, 4 * act + 3 * wbytes  # forward, forward + backward


def sconv_cpu(b, C, threads, batch_sample):
    """Oracle (reference arithmetic, torch CPU) forward + backward of a reduced batch, scaled to b."""
    from oracle import sconv_oracle as SO
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    x = torch.randn(batch_sample, C, C4["X"], C4["Y"], C4["T"], requires_grad=True)
    cot = torch.randn(batch_sample, C, C4["X"], C4["Y"], C4["T"])
    if reference_kind() == "reference":
        from fno.sfno import SpectralConvT as RefConvT
        m = RefConvT(C, C, 20, 20, 8, out_steps=C4["T"], temporal_padding=True, bias=False)

        def step():
            y = m(x)
            y.backward(cot)
            x.grad = None
            for p_ in m.parameters():
                p_.grad = None
    else:
        w = [(0.5 / (C * C) * torch.rand(C, C, 20, 20, 8, 2)).requires_grad_() for _ in range(4)]

        def step():
            y = SO.spectral_conv_t(x, w, 20, 20, 8, C4["T"], None, 0.1, True)
            y.backward(cot)
            x.grad = None
    step()
    t0 = time.perf_counter()
    step()
    el = (time.perf_counter() - t0) * b / batch_sample
    return 1.0 / el, el


def run_sconv(a):
    from torch_cfd_b200.fno import SpectralConvT
    world, rank, local, dev, barrier, max_over_ranks, finish = _dist_setup()
    b, C, K, W = a.batch, a.width, a.steps, max(3, a.warmup)
    torch.manual_seed(rank)
    m = SpectralConvT(C, C, 20, 20, 8, out_steps=C4["T"], temporal_padding=True, bias=False).to(dev)
    x_host = torch.randn(b, C, C4["X"], C4["Y"], C4["T"]).pin_memory()
    cot_host = torch.randn(b, C, C4["X"], C4["Y"], C4["T"]).pin_memory()
    x = x_host.to(dev).requires_grad_()
    cot = cot_host.to(dev)
    plan_launches = [0]

    def fwdbwd():
        y = m(x)
        plan_launches[0] += sum(p.last_launch_count for p in m._cache._plans.values())  # the forward call (5 kernels)
        y.backward(cot)
        plan_launches[0] += sum(p.last_launch_count for p in m._cache._plans.values())  # the backward call (6 kernels)
        x.grad = None
        for p in m.parameters():
            p.grad = None

    def fwd():
        with torch.no_grad():
            m(x)
    ms, clocks = _timed(fwdbwd, K, W, barrier, max_over_ranks, local)
    launches = plan_launches[0] * K // (K + W)
    ms_f, _ = _timed(fwd, K, W, barrier, max_over_ranks, local)
    # end to end: pinned host x and cotangent up, forward + backward, y and grad_x down, every step
    y_host, gx_host = torch.empty_like(x_host), torch.empty_like(x_host)
    e2e = None
    if not a.no_e2e:
        Ke = max(3, min(K, 10))

        def step_host():
            xd = x_host.to(dev, non_blocking=True).requires_grad_()
            cd = cot_host.to(dev, non_blocking=True)
            y = m(xd)
            y.backward(cd)
            y_host.copy_(y.detach(), non_blocking=True)
            gx_host.copy_(xd.grad, non_blocking=True)
            for p in m.parameters():
                p.grad = None
            torch.cuda.current_stream().synchronize()
        for _ in range(2):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            step_host()
        barrier()
        el = max_over_ranks(time.perf_counter() - t0)
        nb = x_host.numel() * 4
        e2e = {"value": world * Ke / el, "unit": UNIT, "h2d_bytes_per_step": 2 * nb, "d2h_bytes_per_step": 2 * nb, "steps": Ke}
    finish()
    if rank != 0:
        return
    alg_f, alg_fb = sconv_alg_bytes(b, C)
    peak, peak_src = peaks()
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["sconv_c4"]["dram_bytes_fwd_bwd"]
    except (OSError, KeyError, ValueError):
        pass
    out = {
        "metric": "sconv_fwd_bwd_steps_per_sec", "value": world * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": sconv_workload_name(b, C), "step": "one forward + backward of the layer on every GPU",
                   "reading": "BASELINE configs[3] with modes_t = 8 needs >= 8 t-frequencies: T = 10 is zero-padded to 20 "
                              "(SpectralConvT temporal_padding, the SFNO OutConv path), SURVEY 8d",
                   "l2": f"activations {x_host.numel() * 4 / 1e6:.0f} MB per tensor exceed the 126 MB L2; no flush",
                   "ms_forward_only": ms_f / K},
        "clocks": clocks, "gpu_launches": launches, "e2e": e2e,
        "roofline": {"bound": "hbm", "kernel": "tcfd_sconv3d forward + backward (all launches of the step)",
                     "achieved": alg_fb / (ms / K * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg_fb / (ms / K * 1e-3) / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_step": alg_fb,
                     "forward_only": {"achieved": alg_f / (ms_f / K * 1e-3) / 1e9, "frac": alg_f / (ms_f / K * 1e-3) / 1e9 / peak,
                                      "algorithmic_bytes": alg_f}},
    }
    if not a.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        v, el = sconv_cpu(b, C, threads, 1)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": reference_kind(),
                               "sample": f"forward + backward of 1 sample of the batch ({reference_kind()}, torch CPU), scaled x{b}"}
    _emit(json.dumps(out))


def run_sconv_reference(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    nb = max(1, min(a.batch, 2))
    v, el = sconv_cpu(a.batch, a.width, threads, nb)
    sample = f"forward + backward of {nb} samples of the batch ({reference_kind()}, torch CPU, {threads} threads), scaled to {a.batch}"
    _emit(json.dumps({
        "impl": "reference", "metric": "sconv_fwd_bwd_steps_per_sec", "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": 1,
        "warmup": 1, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": sconv_workload_name(a.batch, a.width), "note": "CPU arm: rank 0 only"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": reference_kind(), "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


C5 = dict(X=128, Y=128, T=10, modes=(8, 8, 5), in_ch=13)


def fno3d_workload_name(batch, C):
    return f"FNO3d(8, 8, 5, width={C}) full forward (4 spectral layers), x = ({batch}, 13, 128, 128, 10) fp32 GLOBAL batch"


def fno3d_cpu(batch, C, threads, nb):
    """The reference's layer sequence with torch CPU ops (spectral conv = oracle) on nb samples, scaled."""
    from oracle import sconv_oracle as SO
    import torch.nn as nn
    import torch.nn.functional as F
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    x = torch.randn(nb, C5["in_ch"], C5["X"], C5["Y"], C5["T"])
    if reference_kind() == "reference":
        from fno.fno3d import FNO3d as RefFNO3d
        m = RefFNO3d(8, 8, 5, C, input_channel=10).eval()
        with torch.no_grad():
            m(x)
            t0 = time.perf_counter()
            m(x)
            el = (time.perf_counter() - t0) * batch / nb
        return 1.0 / el, el
    p = nn.Conv3d(13, C, 1)
    layers = [([torch.rand(C, C, 8, 8, 5, dtype=torch.cfloat) / (C * C) for _ in range(4)],
               nn.Conv3d(C, C, 1), nn.Conv3d(C, C, 1), nn.Conv3d(C, C, 1)) for _ in range(4)]
    q1, q2 = nn.Conv3d(C, 128, 1), nn.Conv3d(128, 1, 1)

    def fwd():
        with torch.no_grad():
            h = p(x)
            for k, (w, m1, m2, ww) in enumerate(layers):
                h2 = m2(F.gelu(m1(SO.spectral_conv3d(h, w, 8, 8, 5)))) + ww(h)
                h = F.gelu(h2) if k < 3 else h2
            return q2(q1(h))
    fwd()
    t0 = time.perf_counter()
    fwd()
    el = (time.perf_counter() - t0) * batch / nb
    return 1.0 / el, el


def run_fno3d(a):
    from torch_cfd_b200.fno import FNO3d
    world, rank, local, dev, barrier, max_over_ranks, finish = _dist_setup()
    B, C, K, W = a.batch, a.width, a.steps, max(3, a.warmup)
    b = B // world
    torch.manual_seed(0)
    m = FNO3d(8, 8, 5, C, input_channel=10).to(dev).eval()
    x_host = torch.randn(b, 13, C5["X"], C5["Y"], C5["T"]).pin_memory()
    x = x_host.to(dev)

    def fwd():
        with torch.no_grad():
            m(x)
    ms, clocks = _timed(fwd, K, W, barrier, max_over_ranks, local)
    e2e = None
    if not a.no_e2e:
        Ke = max(3, min(K, 10))
        y_host = torch.empty(b, C5["X"], C5["Y"], C5["T"]).pin_memory()

        def step_host():
            with torch.no_grad():
                y, _ = m(x_host.to(dev, non_blocking=True))
                y_host.copy_(y, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        for _ in range(2):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            step_host()
        barrier()
        el = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": Ke / el, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": y_host.numel() * 4,
               "steps": Ke}
    finish()
    if rank != 0:
        return
    peak, peak_src = peaks()
    act = b * C * C5["X"] * C5["Y"] * C5["T"] * 4
    # algorithmic bytes of ONE forward per GPU with the glue fused: lifting reads 13/C + writes 1, every layer reads h
    # twice (conv, skip) + conv output written and read once + layer output written, projection reads 1 (+ 1/C out)
    alg = act * (13.0 / C + 1) + 4 * act * 5 + act * (1 + 1.0 / C)
    # launches per forward: lifting 1 + 4 x (5 spectral-conv kernels + 1 glue) + projection 1
    traffic = None
    try:
        if world == 1 and B == 128 and C == 20:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["fno3d_c5"]["dram_bytes_forward"]
    except Exception:
        pass
    out = {
        "metric": "fno3d_forward_steps_per_sec", "value": K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": fno3d_workload_name(B, C), "step": f"one forward of the GLOBAL batch {B}, sharded {b} per GPU",
                   "samples_per_s": B * K / (ms * 1e-3),
                   "l2": f"activations {act / 1e6:.0f} MB per tensor per GPU exceed the 126 MB L2 up to 4 GPUs; no flush"},
        "clocks": clocks, "gpu_launches": K * (1 + 4 * 6 + 1), "e2e": e2e,
        "roofline": {"bound": "hbm", "kernel": "whole forward (lifting, 4 x [spectral conv + fused layer glue], projection)",
                     "achieved": alg / (ms / K * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (ms / K * 1e-3) / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_step": alg,
                     "note": "bytes of the fused design: every activation tensor read / written once per consumer"},
    }
    if not a.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        v, el = fno3d_cpu(B, C, threads, 2)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": reference_kind(),
                               "sample": f"forward of 2 samples ({reference_kind()}: FNO3d layer sequence, torch CPU), scaled x{B // 2}"}
    _emit(json.dumps(out))


def run_fno3d_reference(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    v, el = fno3d_cpu(a.batch, a.width, threads, 4)
    sample = f"forward of 4 samples ({reference_kind()}: FNO3d layer sequence, torch CPU, {threads} threads), scaled to {a.batch}"
    _emit(json.dumps({
        "impl": "reference", "metric": "fno3d_forward_steps_per_sec", "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": 1,
        "warmup": 1, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": fno3d_workload_name(a.batch, a.width), "note": "CPU arm: rank 0 only"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": reference_kind(), "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def _emit(line: str):
    os.write(_REAL_STDOUT, (line + "\n").encode())


if __name__ == "__main__":
    # the contract is ONE JSON line on stdout: anything native libraries print there (NCCL prints its
    # version banner to stdout) is sent to stderr instead
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    table = {("ns2d", "ours"): run_ours, ("ns2d", "reference"): run_reference,
             ("sconv_c4", "ours"): run_sconv, ("sconv_c4", "reference"): run_sconv_reference,
             ("fno3d_c5", "ours"): run_fno3d, ("fno3d_c5", "reference"): run_fno3d_reference}
    table[(args.workload, args.impl)](args)
