"""Definition-level restatement of the transforms hot path A and B delegate to ``torch.fft``.

TEST INFRASTRUCTURE ONLY (see oracle/ns2d_oracle.py header).  O(n^2) DFT matrices in numpy
float64/complex128; used by tests/test_oracle_cpu.py to pin, at small n, the conventions of the
third-party library the reference calls (``torch.fft.rfft2 / irfft2 / rfftn / irfftn``, torch
2.11.0, "backward" normalisation):

* forward:  X[k] = sum_j x[j] exp(-2 pi i j k / n)            (un-normalised)
* inverse:  x[j] = (1/n) sum_k X[k] exp(+2 pi i j k / n)
* rfft keeps bins 0..n//2 of the LAST transformed axis,
* irfft (C2R) reconstructs the dropped bins by Hermitian symmetry of that axis only, which means
  the imaginary parts of bin 0 and (n even) bin n/2 are ignored AFTER the other axes have been
  inverse-transformed (reference call sites: torch_cfd/equations.py:415,419,422).
"""
import numpy as np


def dft_matrix(n: int, sign: float = -1.0) -> np.ndarray:
    j = np.arange(n)
    return np.exp(sign * 2j * np.pi * np.outer(j, j) / n)


def rfft2_naive(x: np.ndarray) -> np.ndarray:
    """x: (..., n0, n1) real -> (..., n0, n1//2+1) complex."""
    n0, n1 = x.shape[-2:]
    F0, F1 = dft_matrix(n0), dft_matrix(n1)[:, : n1 // 2 + 1]
    return np.einsum("ka,...ab,bl->...kl", F0, x.astype(np.complex128), F1)


def irfft2_naive(X: np.ndarray, n1: int | None = None) -> np.ndarray:
    """X: (..., n0, n1//2+1) complex -> (..., n0, n1) real, torch/numpy C2R semantics."""
    n0, nh = X.shape[-2:]
    n1 = 2 * (nh - 1) if n1 is None else n1
    # inverse along axis -2 (full complex)
    Y = np.einsum("xk,...kl->...xl", dft_matrix(n0, +1.0) / n0, X.astype(np.complex128))
    # C2R along the last axis: out[y] = (1/n1) [Re Y0 + (-1)^y Re Y_{n1/2} + 2 Re sum_{0<l<n1/2} Y_l e^{+2 pi i l y/n1}]
    y = np.arange(n1)
    out = np.zeros(Y.shape[:-1] + (n1,))
    for l in range(nh):
        c = 1.0 if (l == 0 or (n1 % 2 == 0 and l == n1 // 2)) else 2.0
        out += c * (Y[..., l : l + 1] * np.exp(2j * np.pi * l * y / n1)).real
    return out / n1


def rfftn3_naive(x: np.ndarray) -> np.ndarray:
    """(..., X, Y, T) real -> (..., X, Y, T//2+1)."""
    nx, ny, nt = x.shape[-3:]
    return np.einsum("ka,lb,...abc,cm->...klm", dft_matrix(nx), dft_matrix(ny),
                     x.astype(np.complex128), dft_matrix(nt)[:, : nt // 2 + 1])


def irfftn3_naive(X: np.ndarray, s) -> np.ndarray:
    """(..., X, Y, Th) complex -> (..., s0, s1, s2) real with torch's trim/zero-pad-to-s semantics."""
    sx, sy, st = s

    def fit(a, axis, size):
        cur = a.shape[axis]
        if cur >= size:
            sl = [slice(None)] * a.ndim
            sl[axis] = slice(0, size)
            return a[tuple(sl)]
        pad = [(0, 0)] * a.ndim
        pad[axis] = (0, size - cur)
        return np.pad(a, pad)

    X = fit(fit(fit(X.astype(np.complex128), -3, sx), -2, sy), -1, st // 2 + 1)
    Y = np.einsum("xk,yl,...klm->...xym", dft_matrix(sx, +1.0) / sx, dft_matrix(sy, +1.0) / sy, X)
    t = np.arange(st)
    out = np.zeros(Y.shape[:-1] + (st,))
    for m in range(X.shape[-1]):
        c = 1.0 if (m == 0 or (st % 2 == 0 and m == st // 2)) else 2.0
        out += c * (Y[..., m : m + 1] * np.exp(2j * np.pi * m * t / st)).real
    return out / st
