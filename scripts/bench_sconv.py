#!/usr/bin/env python
"""Secondary benchmark (BASELINE.json configs[3], SURVEY 8d "C4"): SFNO spectral convolution
forward + backward.  Reading of the ambiguous config that the reference can actually run:
SpectralConvT(temporal_padding=True): x = (32, 20, 256, 256, 10) fp32, padded 10 -> 20 in time
(11 t-frequencies), modes (20, 20, 8), out_steps 10, width 20.  Also times SpectralConv3d-style
(no padding, T = 16).  Prints one JSON line per case: ms fwd / fwd+bwd (CUDA events), achieved GB/s
against the ALGORITHMIC bytes (read x + write y [+ read g + write gx] + weights), the oracle
(torch CPU) on a reduced batch."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--width", type=int, default=20)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--cpu-batch", type=int, default=1)
a = ap.parse_args()
from torch_cfd_b200.fno import SpectralConvT, SpectralConvS
from oracle import sconv_oracle as SO
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda", 0)
torch.manual_seed(0)
for name, T, pad, mt in [("SpectralConvT pad 10->20, modes (20,20,8)", 10, True, 8), ("SpectralConvS T=16, modes (20,20,8)", 16, False, 8)]:
    C, b = a.width, a.batch
    if pad:
        m = SpectralConvT(C, C, 20, 20, mt, out_steps=T, temporal_padding=True, bias=False).to(dev)
    else:
        m = SpectralConvS(C, C, 20, 20, mt).to(dev)
    x = torch.randn(b, C, 256, 256, T, device=dev, requires_grad=True)
    cot = torch.randn(b, C, 256, 256, T, device=dev)
    def fwd():
        with torch.no_grad():
            return m(x)
    def fwdbwd():
        y = m(x)
        y.backward(cot)
        x.grad = None
        for p in m.parameters():
            p.grad = None
    res = {}
    for label, fn in (("fwd", fwd), ("fwd_bwd", fwdbwd)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[label] = e0.elapsed_time(e1) / a.iters
    act = b * C * 256 * 256 * T * 4
    wbytes = 4 * C * C * 20 * 20 * mt * 8
    alg_f, alg_fb = 2 * act + wbytes, 4 * act + 3 * wbytes
    # CPU oracle on a reduced batch, scaled
    xc = torch.randn(a.cpu_batch, C, 256, 256, T)
    wr = [w.detach().cpu() for w in m.weight]
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        f = (lambda: SO.spectral_conv_t(xc, wr, 20, 20, mt, T, None, 0.1, True)) if pad else (lambda: SO.spectral_conv_s(xc, wr, 20, 20, mt))
        f()
        t0 = time.perf_counter(); f(); cpu_ms = (time.perf_counter() - t0) * 1e3 * b / a.cpu_batch
    print(json.dumps({"case": name, "batch": b, "width": C, "ms_fwd": res["fwd"], "ms_fwd_bwd": res["fwd_bwd"],
                      "alg_GB_fwd": alg_f / 1e9, "GBps_fwd": alg_f / res["fwd"] / 1e6, "frac_fwd": alg_f / res["fwd"] / 1e6 / peak,
                      "GBps_fwd_bwd": alg_fb / res["fwd_bwd"] / 1e6, "frac_fwd_bwd": alg_fb / res["fwd_bwd"] / 1e6 / peak,
                      "cpu_oracle_ms_fwd_scaled": cpu_ms, "cpu_cores": os.cpu_count(),
                      "cpu_sample": f"forward of batch {a.cpu_batch}, scaled x{b // a.cpu_batch}"}))
