"""CPU tier for get_trajectory_imex: the recording kernel on the host-emulation
build against the reference-generated trajectory fixture, the generic (oracle-backed) code path, and
the batch-sharded multi-process form over gloo with world_size 2."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from _common import O, default_dtype, emu_plan, load_golden, oracle_tables, rel_l2, substage_scalars


class OracleEquation:
    """Stand-in with the ImplicitExplicitODE surface get_trajectory_imex uses, backed by the oracle
    (test infrastructure: lets the CPU tests drive the host-side logic without a GPU)."""

    def __init__(self, tb):
        self.tb = tb

    def forward(self, w, dt, steps=1):
        return O.forward(self.tb, w, dt, steps)

    def residual(self, w, dwdt):
        return O.residual(self.tb, w, dwdt)

    def stream_function(self, w):
        return O.vorticity_to_velocity(self.tb, w)[1]


def test_emu_record_kernel_vs_reference_trajectory():
    g = load_golden("ns2d_c1_fp64")
    dtype = torch.float64
    tb = oracle_tables(int(g["n"]), dtype, float(g["viscosity"]), float(g["drag"]), str(g["forcing"]), True, float(g["diam"]))
    w = torch.from_numpy(g["w0_hat"]).reshape(-1, tb.n, tb.n // 2 + 1)
    num_steps, every, dt = int(g["traj_num_steps"]), int(g["traj_every"]), float(g["dt"])
    plan = emu_plan(tb, w.shape[0], dtype)
    beta, gdt, mu = substage_scalars(dtype, dt)
    rec = list(range(0, num_steps, every))
    snaps = {k: torch.empty(w.shape[0], len(rec), tb.n, tb.n // 2 + 1, dtype=torch.complex64)
             for k in ("vorticity", "stream", "vort_t", "residual")}
    done = -1
    for it, t in enumerate(rec):
        gap = t - done - 1
        if gap:
            nxt = torch.empty_like(w)
            plan.step(w, nxt, None, gap, beta, gdt, mu, 1 / (gap * dt))
            w = nxt
        nxt, dw, res = torch.empty_like(w), torch.empty_like(w), torch.empty_like(w)
        plan.step(w, nxt, dw, 1, beta, gdt, mu, 1 / dt)
        w = nxt
        plan.residual(w, dw, res)
        plan.record(w, dw, res, snaps, it)
        done = t
    for k, v in snaps.items():
        ref = torch.from_numpy(g[f"traj_{k}"])
        scale = 1.0 if k != "residual" else None
        if k == "residual":  # a cancellation: measured on the scale of dw/dt
            err = (torch.linalg.norm(v - ref) / torch.linalg.norm(torch.from_numpy(g["traj_vort_t"]))).item()
        else:
            err = rel_l2(v, ref)
        assert err < 2e-7, (k, err)  # complex64 rounding of an fp64 run


def test_generic_path_matches_oracle_trajectory():
    from torch_cfd_b200.solvers import get_trajectory_imex
    dtype = torch.float64
    tb = oracle_tables(32, dtype, 1e-3, 0.1, "vorticity")
    w0 = O.synthetic_vorticity_hat(32, 3, 2, dtype)
    out = get_trajectory_imex(OracleEquation(tb), w0, 1e-3, num_steps=7, record_every_steps=3)
    ref = O.trajectory(tb, w0, 1e-3, 7, 3)
    assert set(out) == {"vorticity", "stream", "vort_t", "residual"}
    for k in out:
        assert out[k].shape == ref[k].shape == (3, 3, 32, 17) and out[k].dtype == torch.complex64
        assert torch.equal(out[k], ref[k]), k
    one = get_trajectory_imex(OracleEquation(tb), w0[0], 1e-3, num_steps=2)  # un-batched input
    assert one["vorticity"].shape == (2, 32, 17)
    with pytest.raises(NotImplementedError):
        get_trajectory_imex(OracleEquation(tb), w0, 1e-3, require_grad=True)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from torch_cfd_b200.solvers import get_trajectory_imex_sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dtype = torch.float64
        tb = oracle_tables(32, dtype, 1e-3, 0.1, "vorticity")
        w0 = O.synthetic_vorticity_hat(32, 4, 9, dtype)  # global batch, same on every rank (seeded)
        per = w0.shape[0] // world
        out = get_trajectory_imex_sharded(OracleEquation(tb), w0[rank * per:(rank + 1) * per], 1e-3, num_steps=5,
                                          record_every_steps=2, fields=("vorticity", "stream"))
        ref = O.trajectory(tb, w0, 1e-3, 5, 2)
        ok = all(torch.equal(out[k], ref[k]) for k in ("vorticity", "stream")) and set(out) == {"vorticity", "stream"}
        # gather_to = None: every rank keeps its shard; gather_to = 1: only rank 1 receives the global result
        mine = get_trajectory_imex_sharded(OracleEquation(tb), w0[rank * per:(rank + 1) * per], 1e-3, num_steps=5,
                                           record_every_steps=2, fields=("vorticity",), gather_to=None)
        ok = ok and torch.equal(mine["vorticity"], ref["vorticity"][rank * per:(rank + 1) * per])
        one = get_trajectory_imex_sharded(OracleEquation(tb), w0[rank * per:(rank + 1) * per], 1e-3, num_steps=5,
                                          record_every_steps=2, fields=("vorticity",), gather_to=1)
        ok = ok and ((rank == 1 and torch.equal(one["vorticity"], ref["vorticity"])) or (rank != 1 and one == {}))
        q.put((rank, bool(ok), tuple(out["vorticity"].shape)))
    finally:
        dist.destroy_process_group()


def test_sharded_trajectory_gloo_world2():
    """Batch sharding + the one all-gather: every rank ends with the global trajectory, identical
    to the single-process result."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), res
    assert all(r[2] == (4, 3, 32, 17) for r in res)
