"""Generate the committed golden fixtures by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Writes tests/golden/*.npz (inputs + reference outputs, a few hundred KB in total).  The reference
is imported read-only (no bytecode is written); ``fno/data_gen/solvers.py`` is loaded by file path
with ``tqdm`` injected (it forgets to import it), and the packages with
import-time side effects (fno.pipeline, fno.utils, fno.data_gen) are never imported.
"""
import importlib.util
import os
import sys

sys.dont_write_bytecode = True
REF = os.environ.get("TORCH_CFD_REF", "/root/reference")
sys.path.insert(0, REF)

import numpy as np
import torch
import torch.fft as fft

HERE = os.path.dirname(os.path.abspath(__file__))


def _np(t):
    return t.detach().cpu().numpy()


def load_solvers():
    spec = importlib.util.spec_from_file_location(
        "ref_solvers", os.path.join(REF, "fno", "data_gen", "solvers.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from tqdm import tqdm
    mod.tqdm = tqdm
    return mod


def ns2d_case(name, n, batch, dtype, viscosity, drag, forcing, steps_list, dt=1e-3, low_storage=True,
              traj=None, diam=2 * torch.pi):
    """forcing: None | 'vorticity' | 'velocity'."""
    torch.set_default_dtype(dtype)
    from torch_cfd.grids import Grid
    from torch_cfd.equations import NavierStokes2DSpectral, RK4CrankNicolsonStepper
    from torch_cfd.forcings import KolmogorovForcing
    from torch_cfd.initial_conditions import vorticity_field

    grid = Grid(shape=(n, n), domain=((0, diam), (0, diam)))
    if forcing is None:
        forcing_fn = None
    else:
        # k=4 is swallowed by **kwargs upstream: effective wave number 1
        forcing_fn = KolmogorovForcing(diam=diam, wave_number=1, grid=grid, k=4, scale=1,
                                       vorticity=(forcing == "vorticity"))
    ns = NavierStokes2DSpectral(viscosity=viscosity, grid=grid, drag=drag, smooth=True,
                                forcing_fn=forcing_fn,
                                solver=RK4CrankNicolsonStepper(low_storage=low_storage))
    w0 = torch.stack([vorticity_field(grid, 4, random_state=s).data for s in range(batch)])
    w0_hat = fft.rfft2(w0)
    if batch == 1 and name.endswith("nobatch"):
        w0_hat = w0_hat[0]
    out = dict(n=n, batch=batch, dt=dt, viscosity=viscosity, drag=drag, diam=float(diam),
               forcing=str(forcing), low_storage=int(low_storage),
               kx=_np(ns.kx), ky=_np(ns.ky), laplace=_np(ns.laplace),
               linear_term=_np(ns.linear_term), filter=_np(ns.filter),
               w0_hat=_np(w0_hat))
    with torch.no_grad():
        out["F0"] = _np(ns.explicit_terms(w0_hat))
        out["G0"] = _np(ns.implicit_terms(w0_hat))
        out["solve0"] = _np(ns.implicit_solve(w0_hat, 0.5 * dt))
        for s in steps_list:
            w, dwdt = ns(w0_hat, dt, steps=s)
            out[f"w_{s}"] = _np(w)
            out[f"dwdt_{s}"] = _np(dwdt)
        out["res_1"] = _np(ns.residual(torch.from_numpy(out["w_1"]), torch.from_numpy(out["dwdt_1"])))
        if traj is not None:
            solvers = load_solvers()
            num_steps, every = traj
            res = solvers.get_trajectory_imex(ns, w0_hat, dt, num_steps=num_steps,
                                              record_every_steps=every, pbar=False)
            for k, v in res.items():
                out[f"traj_{k}"] = _np(v)
            out["traj_num_steps"] = num_steps
            out["traj_every"] = every
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, {k: getattr(v, "shape", v) for k, v in out.items() if k.startswith("w_")})


def imex_case():
    """IMEXStepper (torch_cfd/equations.py:110-246): forward-backward Euler (order 1, alpha = 1), standard
    IMEX Crank-Nicolson (order 1.5) and RK2-CN (order 2, Heun and Ralston weights), 64^2 fp64, forced."""
    torch.set_default_dtype(torch.float64)
    from torch_cfd.grids import Grid
    from torch_cfd.equations import NavierStokes2DSpectral, IMEXStepper
    from torch_cfd.forcings import KolmogorovForcing
    from torch_cfd.initial_conditions import vorticity_field
    n, diam, dt = 64, 2 * torch.pi, 1e-3
    grid = Grid(shape=(n, n), domain=((0, diam), (0, diam)))
    forcing_fn = KolmogorovForcing(diam=diam, wave_number=1, grid=grid, scale=1, vorticity=True)
    w0 = torch.stack([vorticity_field(grid, 4, random_state=s).data for s in range(2)])
    w0_hat = fft.rfft2(w0)
    out = dict(n=n, dt=dt, viscosity=1e-3, drag=0.1, diam=float(diam), forcing="vorticity", w0_hat=_np(w0_hat))
    cases = {"o1": dict(order=1, alpha=1.0), "o15": dict(order=1.5), "o2": dict(order=2),
             "o2r": dict(order=2, alpha=2.0 / 3.0)}
    with torch.no_grad():
        for tag, kw in cases.items():
            ns = NavierStokes2DSpectral(viscosity=1e-3, grid=grid, drag=0.1, smooth=True, forcing_fn=forcing_fn,
                                        solver=IMEXStepper(**kw))
            for s in (1, 3):
                w, dwdt = ns(w0_hat, dt, steps=s)
                out[f"{tag}_w_{s}"] = _np(w)
                out[f"{tag}_dwdt_{s}"] = _np(dwdt)
    np.savez_compressed(os.path.join(HERE, "ns2d_imex.npz"), **out)
    print("imex", [k for k in out if k.endswith("_w_3")])
    torch.set_default_dtype(torch.float32)


def signature_case():
    """Names, defaults and order of the parameters of every reference callable the shims stand in for
    (the drop-in boundary of SURVEY 8b) -> tests/golden/signatures.json."""
    import inspect
    import json
    import torch_cfd.equations as RE, torch_cfd.grids as RG, torch_cfd.forcings as RF, torch_cfd.spectral as RS
    import fno.fno3d as F3, fno.sfno as SF
    sol = load_solvers()
    targets = {
        "stable_time_step": RE.stable_time_step,
        "NavierStokes2DSpectral.__init__": RE.NavierStokes2DSpectral.__init__,
        "NavierStokes2DSpectral.forward": RE.NavierStokes2DSpectral.forward,
        "NavierStokes2DSpectral.explicit_terms": RE.NavierStokes2DSpectral.explicit_terms,
        "NavierStokes2DSpectral.implicit_terms": RE.NavierStokes2DSpectral.implicit_terms,
        "NavierStokes2DSpectral.implicit_solve": RE.NavierStokes2DSpectral.implicit_solve,
        "NavierStokes2DSpectral.residual": RE.NavierStokes2DSpectral.residual,
        "RK4CrankNicolsonStepper.__init__": RE.RK4CrankNicolsonStepper.__init__,
        "RK4CrankNicolsonStepper.forward": RE.RK4CrankNicolsonStepper.forward,
        "IMEXStepper.__init__": RE.IMEXStepper.__init__,
        "IMEXStepper.forward": RE.IMEXStepper.forward,
        "get_trajectory_imex": sol.get_trajectory_imex,
        "Grid.__init__": RG.Grid.__init__,
        "KolmogorovForcing.__init__": RF.KolmogorovForcing.__init__,
        "brick_wall_filter_2d": RS.brick_wall_filter_2d,
        "vorticity_to_velocity": RS.vorticity_to_velocity,
        "fno.SpectralConv3d.__init__": F3.SpectralConv3d.__init__,
        "fno.SpectralConv3d.forward": F3.SpectralConv3d.forward,
        "fno.FNO3d.__init__": F3.FNO3d.__init__,
        "fno.FNO3d.forward": F3.FNO3d.forward,
        "fno.SpectralConvS.__init__": SF.SpectralConvS.__init__,
        "fno.SpectralConvT.__init__": SF.SpectralConvT.__init__,
        "fno.SpectralConvT.forward": SF.SpectralConvT.forward,
    }

    def enc(d):
        if d is inspect._empty:
            return "<required>"
        return d if isinstance(d, (int, float, bool, str, type(None))) else repr(d)

    out = {k: [[p.name, enc(p.default), p.kind.name] for p in inspect.signature(f).parameters.values()]
           for k, f in targets.items()}
    with open(os.path.join(HERE, "signatures.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("signatures", len(out))


def mask_case():
    torch.set_default_dtype(torch.float32)
    from torch_cfd.grids import Grid
    from torch_cfd.spectral import brick_wall_filter_2d
    out = {}
    for n in [16, 32, 64, 128, 256, 512, 1024, 2048]:
        m = brick_wall_filter_2d(Grid(shape=(n, n), domain=((0, 1), (0, 1))))
        rows = (m.sum(1) > 0).nonzero().flatten()
        cols = (m.sum(0) > 0).nonzero().flatten()
        r_lo = int((rows < n // 2).sum())
        r_hi = int((rows >= n // 2).sum())
        out[f"n{n}"] = np.array([r_lo, r_hi, len(cols), int(m.sum())])
        if n <= 64:
            out[f"mask{n}"] = _np(m).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "mask_bounds.npz"), **out)
    print("mask", {k: v.tolist() for k, v in out.items() if k.startswith("n")})


def sconv_cases(grid32=False):
    """grid32=False: tiny grids (oracle pinning); grid32=True: 32-point grids, the smallest the CUDA
    kernels serve, written to sconv32.npz so the kernel path is checked against the reference itself."""
    torch.set_default_dtype(torch.float32)
    from fno.fno3d import SpectralConv3d
    from fno.sfno import SpectralConvS, SpectralConvT
    out = {}

    def run(tag, mod, x, **fw):
        x = x.clone().requires_grad_(True)
        y = mod(x, **fw)
        g = torch.Generator().manual_seed(123)
        cot = torch.randn(y.shape, generator=g)
        (y * cot).sum().backward()
        out[f"{tag}_x"] = _np(x)
        out[f"{tag}_y"] = _np(y)
        out[f"{tag}_cot"] = _np(cot)
        out[f"{tag}_gx"] = _np(x.grad)
        for pname, p in mod.named_parameters():
            key = pname.replace(".", "_")
            out[f"{tag}_p_{key}"] = _np(torch.view_as_real(p) if p.is_complex() else p)
            gp = p.grad
            out[f"{tag}_g_{key}"] = _np(torch.view_as_real(gp) if gp.is_complex() else gp)
        print(tag, tuple(x.shape), "->", tuple(y.shape))

    if grid32:
        torch.manual_seed(1)
        m = SpectralConv3d(3, 2, 6, 5, 4)
        run("c3d_32", m, torch.randn(2, 3, 32, 32, 10))
        m = SpectralConvS(2, 3, 8, 16, 3, bias=True, delta=0.5)
        with torch.no_grad():
            for b in m.bias:
                b.copy_(torch.randn_like(b))
        run("cs_32_bias", m, torch.randn(1, 2, 64, 32, 6))
        m = SpectralConvS(2, 2, 4, 4, 4, norm="ortho")
        run("cs_32_outT", m, torch.randn(1, 2, 32, 32, 8), out_mesh_size=[32, 32, 11])
        m = SpectralConvT(2, 2, 5, 4, 6, out_steps=7, temporal_padding=True, bias=True)
        with torch.no_grad():
            for b in m.bias:
                b.copy_(torch.randn_like(b))
        run("ct_32_pad", m, torch.randn(2, 2, 32, 32, 5))
        np.savez_compressed(os.path.join(HERE, "sconv32.npz"), **out)
        return
    torch.manual_seed(0)
    # SpectralConv3d: (b, C, X, Y, T)
    m = SpectralConv3d(3, 4, 4, 3, 3)
    run("c3d_a", m, torch.randn(2, 3, 16, 8, 6))
    m = SpectralConv3d(2, 2, 5, 5, 4)
    run("c3d_b", m, torch.randn(1, 2, 16, 16, 10))
    # SpectralConvS (real (..,2) weights), no bias
    m = SpectralConvS(3, 2, 4, 4, 3)
    run("cs_a", m, torch.randn(2, 3, 8, 16, 8))
    # SpectralConvS with bias (bias is zero-initialised upstream: randomise it so it is exercised)
    m = SpectralConvS(2, 3, 3, 4, 2, bias=True, delta=0.5)
    with torch.no_grad():
        for b in m.bias:
            b.copy_(torch.randn_like(b))
    run("cs_bias", m, torch.randn(2, 2, 8, 8, 6))
    # SpectralConvS with a different output mesh
    m = SpectralConvS(2, 2, 3, 3, 3)
    run("cs_outmesh", m, torch.randn(1, 2, 8, 8, 8), out_mesh_size=[16, 16, 12])
    # SpectralConvT: temporal padding + out_steps != in steps, bias on
    m = SpectralConvT(2, 2, 4, 4, 5, out_steps=12, temporal_padding=True, bias=True)
    with torch.no_grad():
        for b in m.bias:
            b.copy_(torch.randn_like(b))
    run("ct_pad", m, torch.randn(2, 2, 8, 8, 10))
    # SpectralConvT: no padding, explicit out_steps at call time
    m = SpectralConvT(3, 2, 3, 3, 3, bias=True, temporal_padding=False)
    run("ct_nopad", m, torch.randn(1, 3, 8, 16, 8), out_steps=5)
    np.savez_compressed(os.path.join(HERE, "sconv.npz"), **out)


def ic_case():
    """ic.npz: McWilliams vorticity fields of the reference generator (torch_cfd/initial_conditions.py:170-199)."""
    torch.set_default_dtype(torch.float32)
    from torch_cfd.grids import Grid
    from torch_cfd.initial_conditions import vorticity_field
    out = {}
    for n, pk, seed in ((64, 4, 0), (128, 6, 3)):
        grid = Grid(shape=(n, n), domain=((0, 2 * torch.pi), (0, 2 * torch.pi)))
        out[f"w_{n}_{pk}_{seed}"] = _np(vorticity_field(grid, pk, random_state=seed).data)
    np.savez_compressed(os.path.join(HERE, "ic.npz"), **out)
    print("ic", {k: v.shape for k, v in out.items()})


def legacy_case():
    """legacy_cn.npz: the first-order IMEX / Crank-Nicolson path of fno/data_gen/solvers.py -- one step with a
    per-sample forcing, update_residual, and a short get_trajectory_imex_crank_nicolson run with subsampling."""
    torch.set_default_dtype(torch.float32)
    S = load_solvers()
    n, bsz, visc, dt, diam = 64, 2, 1e-3, 1e-3, 1.0
    g = torch.Generator().manual_seed(11)
    ax = torch.fft.fftfreq(n, d=1.0 / n)
    kx, ky = torch.meshgrid(ax, ax[: n // 2 + 1], indexing="ij")
    env = 1.0 / (1 + (torch.sqrt(kx**2 + ky**2) / 4.0) ** 4)
    w0 = torch.fft.irfft2(torch.fft.rfft2(torch.randn(bsz, n, n, generator=g)) * env, s=(n, n))
    w0 = 4 * w0 / w0.abs().max()
    x = torch.linspace(0, 1, n + 1)[:-1]
    X, Y = torch.meshgrid(x, x, indexing="ij")
    f = 0.1 * (torch.sin(2 * torch.pi * (X + Y)) + torch.cos(2 * torch.pi * (X + Y)))
    out = {"w0": _np(w0), "f": _np(f), "visc": visc, "dt": dt, "diam": diam}
    w_h = fft.rfft2(w0)
    f_b = torch.stack([f, 0.5 * f.flip(0)])            # per-sample forcing
    for tag, ff in (("shared", fft.rfft2(f)), ("batched", fft.rfft2(f_b))):
        w_next, dwdt, _, psi_h, res_h, (kxr, kyr), lap, filt = S.imex_crank_nicolson_step(
            w_h, ff, visc, dt, diam=diam, dealias=True, output_rfft=True)
        r2 = S.update_residual(w_next, dwdt, ff if ff.ndim == 3 else ff.unsqueeze(0), visc, (kxr, kyr), lap,
                               dealias_filter=filt, dealias=True)
        for k, v in (("w_next", w_next), ("dwdt", dwdt), ("psi", psi_h), ("res", res_h), ("res_next", r2)):
            out[f"step_{tag}_{k}"] = _np(v)
    out["f_b"] = _np(f_b)
    res = S.get_trajectory_imex_crank_nicolson(w0, f, visc=visc, T=0.02, delta_t=dt, record_steps=4, diam=diam,
                                               dealias=True, subsample=2, pbar=False)
    for k, v in res.items():
        out[f"traj_{k}"] = _np(v)
    print("legacy", {k: tuple(v.shape) for k, v in res.items()})
    np.savez_compressed(os.path.join(HERE, "legacy_cn.npz"), **out)


def sfno_case():
    """sfno.npz: the x / y change of mesh of SpectralConv.forward (fno/base.py:229-237), SpectralConvT with the
    HelmholtzProjection post-process (fno/sfno.py:116-193, :452) -- values and all gradients -- and a small SFNO
    (fno/sfno.py:460-620): its state_dict, an input and the outputs for two output lengths."""
    torch.set_default_dtype(torch.float32)
    from fno.sfno import SFNO, HelmholtzProjection, SpectralConvS, SpectralConvT
    out = {}

    def run(tag, mod, x, **fw):
        x = x.clone().requires_grad_(True)
        y = mod(x, **fw)
        cot = torch.randn(y.shape, generator=torch.Generator().manual_seed(123))
        (y * cot).sum().backward()
        for k, v in (("x", x), ("y", y), ("cot", cot), ("gx", x.grad)):
            out[f"{tag}_{k}"] = _np(v)
        for pname, p in mod.named_parameters():
            key = pname.replace(".", "_")
            out[f"{tag}_p_{key}"] = _np(p)
            out[f"{tag}_g_{key}"] = _np(p.grad)
        print(tag, tuple(x.shape), "->", tuple(y.shape))

    torch.manual_seed(5)
    m = SpectralConvS(2, 2, 4, 4, 3)
    run("cs_up", m, torch.randn(1, 2, 32, 32, 8), out_mesh_size=[64, 64, 12])
    m = SpectralConvS(2, 3, 6, 5, 3, norm="ortho")
    run("cs_down", m, torch.randn(2, 2, 64, 64, 8), out_mesh_size=[32, 32, 6])
    m = SpectralConvT(2, 2, 4, 4, 3, out_steps=6, temporal_padding=True, bias=True,
                      postprocess=HelmholtzProjection(n_grid=32, diam=1))
    with torch.no_grad():
        for b in m.bias:
            b.copy_(torch.randn_like(b))
    run("ct_helm", m, torch.randn(2, 2, 32, 32, 5))
    torch.manual_seed(6)
    model = SFNO(4, 4, 3, 8, num_spectral_layers=3, latent_steps=5)
    with torch.no_grad():
        for b in model.output_operator.conv.bias:
            b.copy_(0.1 * torch.randn_like(b))
    x = torch.randn(2, 32, 32, 6)
    with torch.no_grad():
        out["sfno_x"] = _np(x)
        out["sfno_y"] = _np(model(x))
        out["sfno_y9"] = _np(model(x, out_steps=9))
    for k, v in model.state_dict().items():
        out[f"sfno_sd_{k}"] = _np(v)
    print("sfno", tuple(x.shape), "->", out["sfno_y"].shape, out["sfno_y9"].shape, len(model.state_dict()), "state_dict entries")
    np.savez_compressed(os.path.join(HERE, "sfno.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sfno":
        sfno_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ic":
        ic_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "legacy":
        legacy_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "signatures":
        signature_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "imex":  # only tests/golden/ns2d_imex.npz
        imex_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sconv32":
        sconv_cases(grid32=True)
        sys.exit(0)
    # C1 (BASELINE.json configs[0]): 64^2, batch 1, fp64, Kolmogorov vorticity forcing, drag 0.1
    ns2d_case("ns2d_c1_fp64", 64, 1, torch.float64, 1e-3, 0.1, "vorticity", [1, 2, 100], traj=(6, 2))
    # C2-like small: fp32, batch 3, unforced, no drag
    ns2d_case("ns2d_fp32_unforced", 64, 3, torch.float32, 1e-3, 0.0, None, [1, 10, 50])
    # velocity-form forcing (data_gen_Kolmogorov2d.py form), fp64, n=32
    ns2d_case("ns2d_fp64_velforce", 32, 2, torch.float64, 1e-3, 0.1, "velocity", [1, 20])
    # (low_storage=False cannot be generated: upstream builds an integer `betas` tensor and
    #  nn.Parameter rejects it -- RuntimeError at torch_cfd/equations.py:326,167.)
    # un-batched (n, nh) input, n=128, fp32
    ns2d_case("ns2d_fp32_n128_nobatch", 128, 1, torch.float32, 1e-3, 0.1, "vorticity", [1, 5],
              traj=(4, 1))
    mask_case()
    signature_case()
    imex_case()
    sconv_cases()
    sconv_cases(grid32=True)
    sfno_case()
    legacy_case()
    ic_case()
