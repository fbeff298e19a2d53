"""ctypes binding of the C ABI in include/tcfd.h (libtcfd.so, hand-written sm_100a kernels).

PyTorch tensors cross this boundary only as raw device pointers (``tensor.data_ptr()``) plus
sizes; the current CUDA stream is passed as an integer handle.  There is no CPU fallback: if the
shared library is missing or no CUDA device is usable, every entry point raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# TCFD_LIB: development override (an alternative build of the same CUDA library, e.g. a tuning variant)
_LIB_PATH = os.environ.get("TCFD_LIB") or os.path.join(_HERE, "libtcfd.so")


class _Desc(ctypes.Structure):
    _fields_ = [
        ("n", ctypes.c_int),
        ("prec", ctypes.c_int),
        ("max_batch", ctypes.c_int),
        ("kappa_x", ctypes.c_void_p),
        ("kappa_y", ctypes.c_void_p),
        ("neg_inv_lap", ctypes.c_void_p),
        ("linear_term", ctypes.c_void_p),
        ("filter", ctypes.c_void_p),
        ("f_hat", ctypes.c_void_p),
    ]


class _SconvDesc(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int) for k in
                ("X", "Y", "T_in", "t_pad", "T_out", "Ci", "Co", "mx", "my", "mt", "norm", "max_batch")]


class TcfdLibrary:
    """A loaded libtcfd with typed entry points."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise RuntimeError(
                f"torch-cfd_b200: CUDA library not found at {path}. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C torch-cfd_b200/csrc`. "
                "There is no CPU fallback."
            )
        self.path = path
        self.c = ctypes.CDLL(path)
        c = self.c
        vp, ci, dp = ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double)
        c.tcfd_last_error.restype = ctypes.c_char_p
        c.tcfd_version.restype = ctypes.c_char_p
        c.tcfd_ns2d_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(_Desc)]
        c.tcfd_ns2d_destroy.argtypes = [vp]
        c.tcfd_ns2d_set_forcing.argtypes = [vp, vp]
        c.tcfd_ns2d_workspace_bytes.argtypes = [vp]
        c.tcfd_ns2d_workspace_bytes.restype = ctypes.c_size_t
        c.tcfd_ns2d_last_launch_count.argtypes = [vp]
        c.tcfd_ns2d_check.argtypes = [vp]
        c.tcfd_ns2d_schedule.argtypes = [vp]
        c.tcfd_ns2d_schedule.restype = ctypes.c_int
        c.tcfd_ns2d_check.restype = ctypes.c_int
        c.tcfd_ns2d_step.argtypes = [vp, vp, vp, vp, ci, ci, ci, dp, dp, dp, ctypes.c_double, vp]
        c.tcfd_ns2d_step_host.argtypes = [vp, vp, vp, vp, ci, ci, ci, dp, dp, dp, ctypes.c_double, vp]
        c.tcfd_ns2d_step_timed.argtypes = [vp, vp, vp, ci, ci, ci, dp, dp, dp, vp,
                                           ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)]
        c.tcfd_ns2d_explicit_terms.argtypes = [vp, vp, vp, ci, vp]
        c.tcfd_ns2d_residual.argtypes = [vp, vp, vp, vp, ci, vp]
        c.tcfd_ns2d_record.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, vp]
        pp = ctypes.POINTER(vp)
        c.tcfd_sconv3d_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(_SconvDesc)]
        c.tcfd_sconv3d_destroy.argtypes = [vp]
        c.tcfd_sconv3d_workspace_bytes.argtypes = [vp]
        c.tcfd_sconv3d_workspace_bytes.restype = ctypes.c_size_t
        c.tcfd_sconv3d_xhat_elems.argtypes = [vp, ci]
        c.tcfd_sconv3d_xhat_elems.restype = ctypes.c_size_t
        c.tcfd_sconv3d_last_launch_count.argtypes = [vp]
        c.tcfd_sconv3d_forward.argtypes = [vp, vp, pp, pp, ctypes.c_float, vp, vp, ci, vp]
        c.tcfd_sconv3d_backward.argtypes = [vp, vp, vp, pp, vp, pp, pp, ctypes.c_float, ci, vp]
        c.tcfd_sconv3d_yhat_elems.argtypes = [vp, ci]
        c.tcfd_sconv3d_yhat_elems.restype = ctypes.c_size_t
        c.tcfd_sconv3d_analysis.argtypes = [vp, vp, pp, pp, ctypes.c_float, vp, vp, ci, vp]
        c.tcfd_sconv3d_synthesis.argtypes = [vp, vp, vp, ci, vp]
        c.tcfd_sconv3d_synthesis_backward.argtypes = [vp, vp, vp, ci, vp]
        c.tcfd_sconv3d_analysis_backward.argtypes = [vp, vp, vp, pp, vp, pp, pp, ctypes.c_float, ci, vp]
        c.tcfd_fft2_create.argtypes = [ctypes.POINTER(vp), ci, ci]
        c.tcfd_fft2_destroy.argtypes = [vp]
        c.tcfd_fft2_last_launch_count.argtypes = [vp]
        c.tcfd_fft2_irfft2.argtypes = [vp, vp, vp, ci, vp]
        c.tcfd_fft2_rfft2.argtypes = [vp, vp, vp, ci, vp]
        c.tcfd_resample_bilinear.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp]
        sz = ctypes.c_size_t
        c.tcfd_fno_pointwise_linear.argtypes = [vp, vp, vp, vp, ci, ci, ci, sz, vp]
        c.tcfd_fno_layer_glue.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, sz, vp]
        c.tcfd_fno_project.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, sz, vp]

    def check(self, rc: int, what: str):
        if rc != 0:
            msg = self.c.tcfd_last_error().decode("utf-8", "replace")
            raise RuntimeError(f"torch-cfd_b200: {what} failed ({rc}): {msg}")

    def version(self) -> str:
        return self.c.tcfd_version().decode()


_LIB: Optional[TcfdLibrary] = None


def load_library() -> TcfdLibrary:
    """The product library.  Raises loudly when it has not been built."""
    global _LIB
    if _LIB is None:
        _LIB = TcfdLibrary(_LIB_PATH)
    return _LIB


class HandleStore(dict):
    """Per-module store of native handles (plans).  Handles wrap ctypes pointers, which can be neither copied
    nor pickled: ``copy.deepcopy(module)`` / ``torch.save(module)`` see an EMPTY store instead (the copy builds
    its own handles on first use)."""

    def __deepcopy__(self, memo):
        return HandleStore()

    def __reduce__(self):
        return (HandleStore, ())


def _darr(vals: Sequence[float]):
    return (ctypes.c_double * len(vals))(*[float(v) for v in vals])


def _stream_handle(t: torch.Tensor) -> int:
    if t.is_cuda:
        return torch.cuda.current_stream(t.device).cuda_stream
    return 0


class NS2DPlan:
    """Owns one ``tcfd_ns2d_t`` handle: tables + workspace for (n, precision, max_batch) on the
    current device.  Host tensors in, raw pointers out."""

    def __init__(self, lib: TcfdLibrary, n: int, dtype: torch.dtype, max_batch: int,
                 kappa_x: torch.Tensor, kappa_y: torch.Tensor, neg_inv_lap: torch.Tensor,
                 linear_term: torch.Tensor, filter: Optional[torch.Tensor], f_hat: Optional[torch.Tensor]):
        assert dtype in (torch.float32, torch.float64)
        self.lib, self.n, self.nh, self.dtype, self.max_batch = lib, n, n // 2 + 1, dtype, max_batch
        self.cdtype = torch.complex64 if dtype == torch.float32 else torch.complex128
        self._h = ctypes.c_void_p()

        def host(t, dt, shape):
            t = t.detach().to("cpu", dt).contiguous()
            assert tuple(t.shape) == tuple(shape), (tuple(t.shape), shape)
            return t

        keep = [host(kappa_x, dtype, (n,)), host(kappa_y, dtype, (self.nh,)),
                host(neg_inv_lap, dtype, (n, self.nh)), host(linear_term, dtype, (n, self.nh))]
        filt = None if filter is None else host(filter, dtype, (n, self.nh))
        fh = None if f_hat is None else host(f_hat, self.cdtype, (n, self.nh))
        d = _Desc(n, 32 if dtype == torch.float32 else 64, max_batch,
                  keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr(), keep[3].data_ptr(),
                  None if filt is None else filt.data_ptr(), None if fh is None else fh.data_ptr())
        lib.check(lib.c.tcfd_ns2d_create(ctypes.byref(self._h), ctypes.byref(d)), "tcfd_ns2d_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.c.tcfd_ns2d_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.c.tcfd_ns2d_workspace_bytes(self._h))

    @property
    def last_launch_count(self) -> int:
        return int(self.lib.c.tcfd_ns2d_last_launch_count(self._h))

    @property
    def dataflow(self) -> bool:
        """True when tcfd_ns2d_step runs as ONE persistent dataflow launch per call."""
        return bool(self.lib.c.tcfd_ns2d_schedule(self._h))

    def check(self):
        """Raise if a kernel of an earlier asynchronous call reported a failure (tcfd_ns2d_check);
        meaningful after the stream was synchronised."""
        self.lib.check(self.lib.c.tcfd_ns2d_check(self._h), "tcfd_ns2d_check")

    def set_forcing(self, f_hat: Optional[torch.Tensor]):
        if f_hat is None:
            self.lib.check(self.lib.c.tcfd_ns2d_set_forcing(self._h, None), "tcfd_ns2d_set_forcing")
            return
        fh = f_hat.detach().to("cpu", self.cdtype).contiguous()
        assert tuple(fh.shape) == (self.n, self.nh)
        self.lib.check(self.lib.c.tcfd_ns2d_set_forcing(self._h, fh.data_ptr()), "tcfd_ns2d_set_forcing")

    def _check_state(self, t: torch.Tensor, name: str):
        if t.dtype != self.cdtype or not t.is_contiguous() or t.dim() != 3 or tuple(t.shape[1:]) != (self.n, self.nh):
            raise ValueError(f"{name}: expected contiguous {self.cdtype} tensor (B, {self.n}, {self.nh}), "
                             f"got {t.dtype} {tuple(t.shape)} contiguous={t.is_contiguous()}")

    def step(self, w_in: torch.Tensor, w_out: torch.Tensor, dwdt: Optional[torch.Tensor], steps: int,
             beta: Sequence[float], gdt: Sequence[float], mu: Sequence[float], inv_total_dt: float,
             host: bool = False):
        self._check_state(w_in, "w_in")
        self._check_state(w_out, "w_out")
        if dwdt is not None:
            self._check_state(dwdt, "dwdt")
        fn = self.lib.c.tcfd_ns2d_step_host if host else self.lib.c.tcfd_ns2d_step
        stream = torch.cuda.current_stream().cuda_stream if host and torch.cuda.is_available() else _stream_handle(w_in)
        rc = fn(self._h, w_in.data_ptr(), w_out.data_ptr(), None if dwdt is None else dwdt.data_ptr(),
                w_in.shape[0], int(steps), len(beta), _darr(beta), _darr(gdt), _darr(mu),
                float(inv_total_dt), stream)
        self.lib.check(rc, "tcfd_ns2d_step")

    KERNEL_KINDS = ("rows_inv", "rows_fwd_inv", "rows_fwd", "cols")

    def kernel_times(self, w_in: torch.Tensor, dt: float, solver, steps: int = 1):
        """Instrumented pass (tcfd_ns2d_step_timed): per-kernel-kind device time of `steps` steps.
        Returns {kind: {"launches": c, "ms_total": t, "us_per_launch": ...}}."""
        self._check_state(w_in, "w_in")
        beta, gdt, mu = solver.substage_scalars(dt)
        out = torch.empty_like(w_in)
        ms = (ctypes.c_float * 4)()
        cnt = (ctypes.c_int * 4)()
        rc = self.lib.c.tcfd_ns2d_step_timed(self._h, w_in.data_ptr(), out.data_ptr(), w_in.shape[0], int(steps),
                                             len(beta), _darr(beta), _darr(gdt), _darr(mu), _stream_handle(w_in),
                                             ms, cnt)
        self.lib.check(rc, "tcfd_ns2d_step_timed")
        res = {"steps": steps}
        kinds = ("rows_inv", "flow_call", "rows_fwd", "cols") if self.dataflow else self.KERNEL_KINDS
        for i, k in enumerate(kinds):
            res[k] = {"launches": cnt[i], "ms_total": ms[i],
                      "us_per_launch": (1e3 * ms[i] / cnt[i]) if cnt[i] else None}
        return res

    def record(self, w, dwdt, res, snaps, it: int):
        """snaps: dict with (B, n_t, n, nh) complex tensors (or None) under vorticity/stream/vort_t/residual."""
        self._check_state(w, "w")
        first = next(v for v in snaps.values() if v is not None)
        n_t = first.shape[1]
        out_prec = 32 if first.dtype == torch.complex64 else 64
        ptr = lambda t: None if t is None else t.data_ptr()
        rc = self.lib.c.tcfd_ns2d_record(self._h, w.data_ptr(), ptr(dwdt), ptr(res), ptr(snaps.get("vorticity")),
                                         ptr(snaps.get("stream")), ptr(snaps.get("vort_t")), ptr(snaps.get("residual")),
                                         w.shape[0], n_t, int(it), out_prec, _stream_handle(w))
        self.lib.check(rc, "tcfd_ns2d_record")

    def explicit_terms(self, w_in: torch.Tensor, out: torch.Tensor):
        self._check_state(w_in, "w_in")
        self._check_state(out, "out")
        self.lib.check(self.lib.c.tcfd_ns2d_explicit_terms(self._h, w_in.data_ptr(), out.data_ptr(),
                                                           w_in.shape[0], _stream_handle(w_in)),
                       "tcfd_ns2d_explicit_terms")

    def residual(self, w_in: torch.Tensor, wt_in: torch.Tensor, out: torch.Tensor):
        for t, nm in ((w_in, "w_in"), (wt_in, "wt_in"), (out, "out")):
            self._check_state(t, nm)
        self.lib.check(self.lib.c.tcfd_ns2d_residual(self._h, w_in.data_ptr(), wt_in.data_ptr(), out.data_ptr(),
                                                     w_in.shape[0], _stream_handle(w_in)),
                       "tcfd_ns2d_residual")


class FFT2Plan:
    """Owns one ``tcfd_fft2_t`` handle (twiddles + scratch) for (n, precision) on the current device: batched
    rfft2 / irfft2 in the reference's layouts and the bilinear resampling of the data-generation scripts."""

    def __init__(self, lib: TcfdLibrary, n: int, dtype: torch.dtype):
        assert dtype in (torch.float32, torch.float64)
        self.lib, self.n, self.nh, self.dtype = lib, n, n // 2 + 1, dtype
        self.cdtype = torch.complex64 if dtype == torch.float32 else torch.complex128
        self._h = ctypes.c_void_p()
        rc = lib.c.tcfd_fft2_create(ctypes.byref(self._h), n, 32 if dtype == torch.float32 else 64)
        if rc != 0:
            raise ValueError(f"torch-cfd_b200: tcfd_fft2_create failed ({rc}): "
                             + lib.c.tcfd_last_error().decode("utf-8", "replace"))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.c.tcfd_fft2_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def last_launch_count(self) -> int:
        return int(self.lib.c.tcfd_fft2_last_launch_count(self._h))

    def irfft2(self, x_hat: torch.Tensor) -> torch.Tensor:
        """(*, n, n//2+1) complex -> (*, n, n) real, == torch.fft.irfft2."""
        if x_hat.dtype != self.cdtype or tuple(x_hat.shape[-2:]) != (self.n, self.nh):
            raise ValueError(f"irfft2: expected {self.cdtype} (*, {self.n}, {self.nh}), got {x_hat.dtype} {tuple(x_hat.shape)}")
        xh = x_hat.contiguous()
        out = torch.empty(tuple(xh.shape[:-1]) + (self.n,), dtype=self.dtype, device=xh.device)
        count = max(1, xh.numel() // (self.n * self.nh))
        self.lib.check(self.lib.c.tcfd_fft2_irfft2(self._h, xh.data_ptr(), out.data_ptr(), count, _stream_handle(xh)),
                       "tcfd_fft2_irfft2")
        return out

    def rfft2(self, x: torch.Tensor) -> torch.Tensor:
        """(*, n, n) real -> (*, n, n//2+1) complex, == torch.fft.rfft2."""
        if x.dtype != self.dtype or tuple(x.shape[-2:]) != (self.n, self.n):
            raise ValueError(f"rfft2: expected {self.dtype} (*, {self.n}, {self.n}), got {x.dtype} {tuple(x.shape)}")
        xc = x.contiguous()
        out = torch.empty(tuple(xc.shape[:-1]) + (self.nh,), dtype=self.cdtype, device=xc.device)
        count = max(1, xc.numel() // (self.n * self.n))
        self.lib.check(self.lib.c.tcfd_fft2_rfft2(self._h, xc.data_ptr(), out.data_ptr(), count, _stream_handle(xc)),
                       "tcfd_fft2_rfft2")
        return out


def resample_bilinear(lib: TcfdLibrary, x: torch.Tensor, n_out: int, dtype: torch.dtype) -> torch.Tensor:
    """F.interpolate(x.to(dtype), size=(n_out, n_out), mode="bilinear") for square (*, n, n) real fields."""
    if x.dtype not in (torch.float32, torch.float64) or dtype not in (torch.float32, torch.float64):
        raise TypeError("resample_bilinear: float32 / float64 only")
    if x.shape[-1] != x.shape[-2]:
        raise ValueError("resample_bilinear: square fields only")
    xc = x.contiguous()
    n_in = xc.shape[-1]
    out = torch.empty(tuple(xc.shape[:-2]) + (n_out, n_out), dtype=dtype, device=xc.device)
    count = max(1, xc.numel() // (n_in * n_in))
    prec = lambda d: 32 if d == torch.float32 else 64
    lib.check(lib.c.tcfd_resample_bilinear(xc.data_ptr(), out.data_ptr(), prec(xc.dtype), prec(dtype), count, n_in, n_out,
                                           _stream_handle(xc)), "tcfd_resample_bilinear")
    return out


_NORMS = {"backward": 0, None: 0, "ortho": 1, "forward": 2}


def _ptr_array(tensors):
    if tensors is None:
        return None
    return (ctypes.c_void_p * 4)(*[t.data_ptr() for t in tensors])


class SConv3dPlan:
    """Owns one ``tcfd_sconv3d_t`` handle (t-axis tables, twiddles, spectral workspace) for one
    layer geometry on the current device."""

    def __init__(self, lib: TcfdLibrary, X, Y, T_in, t_pad, T_out, Ci, Co, mx, my, mt, norm, max_batch):
        self.lib = lib
        self.geom = (X, Y, T_in, t_pad, T_out, Ci, Co, mx, my, mt, norm)
        self.max_batch = max_batch
        self._h = ctypes.c_void_p()
        d = _SconvDesc(X, Y, T_in, t_pad, T_out, Ci, Co, mx, my, mt, _NORMS[norm], max_batch)
        rc = lib.c.tcfd_sconv3d_create(ctypes.byref(self._h), ctypes.byref(d))
        if rc != 0:
            msg = lib.c.tcfd_last_error().decode("utf-8", "replace")
            raise ValueError(f"torch-cfd_b200: tcfd_sconv3d_create failed ({rc}): {msg}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.c.tcfd_sconv3d_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def last_launch_count(self) -> int:
        return int(self.lib.c.tcfd_sconv3d_last_launch_count(self._h))

    def xhat_elems(self, batch: int) -> int:
        return int(self.lib.c.tcfd_sconv3d_xhat_elems(self._h, batch))

    def forward(self, x, w, bias, delta, y, xhat):
        rc = self.lib.c.tcfd_sconv3d_forward(self._h, x.data_ptr(), _ptr_array(w), _ptr_array(bias), float(delta),
                                             y.data_ptr(), None if xhat is None else xhat.data_ptr(),
                                             x.shape[0], _stream_handle(x))
        self.lib.check(rc, "tcfd_sconv3d_forward")

    def backward(self, gy, xhat, w, gx, gw, gbias, delta):
        rc = self.lib.c.tcfd_sconv3d_backward(self._h, gy.data_ptr(), xhat.data_ptr(), _ptr_array(w),
                                              None if gx is None else gx.data_ptr(), _ptr_array(gw),
                                              _ptr_array(gbias), float(delta), gy.shape[0], _stream_handle(gy))
        self.lib.check(rc, "tcfd_sconv3d_backward")

    # the two halves as separate calls (tcfd_sconv3d_analysis / synthesis and their adjoints)
    def yhat_shape(self, batch: int):
        X, Y, T_in, t_pad, T_out, Ci, Co, mx, my, mt, norm = self.geom
        return (batch, Co, 2 * mx, 2 * my, mt)

    def analysis(self, x, w, bias, delta, yhat, xhat):
        rc = self.lib.c.tcfd_sconv3d_analysis(self._h, x.data_ptr(), _ptr_array(w), _ptr_array(bias), float(delta),
                                              yhat.data_ptr(), None if xhat is None else xhat.data_ptr(),
                                              x.shape[0], _stream_handle(x))
        self.lib.check(rc, "tcfd_sconv3d_analysis")

    def synthesis(self, yhat, y):
        rc = self.lib.c.tcfd_sconv3d_synthesis(self._h, yhat.data_ptr(), y.data_ptr(), yhat.shape[0], _stream_handle(yhat))
        self.lib.check(rc, "tcfd_sconv3d_synthesis")

    def synthesis_backward(self, gy, gyhat):
        rc = self.lib.c.tcfd_sconv3d_synthesis_backward(self._h, gy.data_ptr(), gyhat.data_ptr(), gy.shape[0],
                                                        _stream_handle(gy))
        self.lib.check(rc, "tcfd_sconv3d_synthesis_backward")

    def analysis_backward(self, gyhat, xhat, w, gx, gw, gbias, delta):
        rc = self.lib.c.tcfd_sconv3d_analysis_backward(self._h, gyhat.data_ptr(), xhat.data_ptr(), _ptr_array(w),
                                                       None if gx is None else gx.data_ptr(), _ptr_array(gw),
                                                       _ptr_array(gbias), float(delta), gyhat.shape[0],
                                                       _stream_handle(gyhat))
        self.lib.check(rc, "tcfd_sconv3d_analysis_backward")


# ------------------------------------------------------------------------------------------------
# FNO3d layer glue (tcfd_fno_*): thin wrappers over raw pointers; every tensor fp32, contiguous, on one
# device (CUDA for the product library, CPU for the host-emulation build used by the tests)
def _fno_check(*ts):
    dev = ts[0].device
    for t in ts:
        if t is None:
            continue
        if t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
            raise ValueError("tcfd_fno_*: expected contiguous float32 tensors on one device")


def _ptr(t):
    return None if t is None else t.data_ptr()


def fno_pointwise_linear(lib: TcfdLibrary, x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]):
    """y = Conv3d(kernel 1)(x): x (b, Ci, X, Y, T), weight (Co, Ci, 1, 1, 1) or (Co, Ci)."""
    w = weight.reshape(weight.shape[0], -1).contiguous()
    _fno_check(x, w, bias)
    b, Ci = x.shape[:2]
    Co, npts = w.shape[0], x[0, 0].numel()
    y = torch.empty((b, Co) + tuple(x.shape[2:]), dtype=x.dtype, device=x.device)
    lib.check(lib.c.tcfd_fno_pointwise_linear(x.data_ptr(), y.data_ptr(), w.data_ptr(), _ptr(bias), b, Ci, Co, npts,
                                              _stream_handle(x)), "tcfd_fno_pointwise_linear")
    return y


def fno_glue_host_weights(w1, b1, w2, b2, ww, bw):
    """The layer's weights as the HOST tensors tcfd_fno_layer_glue takes (they travel as a kernel
    parameter): one device->host copy, to be cached by the caller while the parameters do not change."""
    C = w1.shape[0]
    host = lambda t, shape: None if t is None else t.detach().to("cpu", torch.float32).reshape(shape).contiguous()
    return (host(w1, (C, C)), host(b1, (C,)), host(w2, (C, C)), host(b2, (C,)), host(ww, (C, C)), host(bw, (C,)))


def fno_layer_glue(lib: TcfdLibrary, conv_out: torch.Tensor, x: torch.Tensor, host_weights, act: bool):
    """y = act( mlp2(gelu(mlp1(conv_out))) + w(x) ), all channel counts equal; host_weights from
    fno_glue_host_weights()."""
    C = x.shape[1]
    _fno_check(conv_out, x)
    for t in host_weights:
        if t is not None and (t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous()):
            raise ValueError("layer glue: weights must be contiguous float32 HOST tensors")
    if conv_out.shape != x.shape:
        raise ValueError("layer glue: conv_out and x must have the same shape")
    w1, b1, w2, b2, ww, bw = host_weights
    if tuple(w1.shape) != (C, C) or tuple(w2.shape) != (C, C) or tuple(ww.shape) != (C, C):
        raise ValueError("layer glue: weight matrices must be (C, C)")
    y = torch.empty_like(x)
    lib.check(lib.c.tcfd_fno_layer_glue(conv_out.data_ptr(), x.data_ptr(), y.data_ptr(), w1.data_ptr(), _ptr(b1),
                                        w2.data_ptr(), _ptr(b2), ww.data_ptr(), _ptr(bw), 1 if act else 0,
                                        x.shape[0], C, x[0, 0].numel(), _stream_handle(x)), "tcfd_fno_layer_glue")
    return y


def fno_project(lib: TcfdLibrary, x: torch.Tensor, w1, b1, w2, b2, act: bool):
    """y = mlp2(act(mlp1(x))) with one output channel: x (b, C, X, Y, T) -> (b, 1, X, Y, T)."""
    C = x.shape[1]
    w1m = w1.reshape(w1.shape[0], C).contiguous()
    M = w1m.shape[0]
    w2m = w2.reshape(-1).contiguous()
    if w2m.numel() != M:
        raise ValueError("project: mlp2 must have exactly one output channel")
    _fno_check(x, w1m, w2m, b1, b2)
    y = torch.empty((x.shape[0], 1) + tuple(x.shape[2:]), dtype=x.dtype, device=x.device)
    lib.check(lib.c.tcfd_fno_project(x.data_ptr(), y.data_ptr(), w1m.data_ptr(), _ptr(b1), w2m.data_ptr(), _ptr(b2),
                                     1 if act else 0, x.shape[0], C, M, x[0, 0].numel(), _stream_handle(x)),
              "tcfd_fno_project")
    return y
