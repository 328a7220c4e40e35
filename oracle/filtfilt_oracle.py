"""CPU oracle for the zero-phase Bessel filters (LPF / BPF).

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import this; the product never does.

Reference call sites: ``opticomlib/devices.py:814-823`` (BPF:
``bessel(n, BW/2, 'low', fs, 'sos', norm='mag')`` + ``sosfiltfilt(axis=-1)`` on
signal and noise separately) and ``devices.py:1363-1368`` (LPF: ``Wn=BW``,
``.real`` of the result).  The arithmetic lives in SciPy (not under
/root/reference; pinned scipy==1.12.0, image 1.18.1), so this file restates the
published ``sosfiltfilt`` algorithm:

    ntaps = 2S+1 - min(#(b2==0), #(a2==0));  edge = 3*ntaps
    ext   = odd extension of x by `edge` samples on both sides
    zi    = per-section steady state of the unit step response, scaled by the
            DC gain of the sections before it (sosfilt_zi)
    forward  sosfilt(ext,  zi*ext[0]);  backward the same on the reversed
    output with zi*y[-1];  reverse;  strip `edge` samples.

Parity status: PINNED against ``scipy.signal.sosfiltfilt`` itself and against
outputs of the reference LPF/BPF (``tests/golden/filters_*.npz``).
The inner recurrence runs in C (``filtfilt_oracle.c``) when the helper library
is built, else in a (slow) Python loop -- same arithmetic either way.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liboracle.so")
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            lib.oracle_sosfilt_f64.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                               ctypes.c_long, ctypes.c_long, ctypes.c_void_p]
            lib.oracle_sosfilt_f64.restype = None
            _LIB = lib
        else:
            _LIB = False
    return _LIB


def bessel_sos(n: int, wn_hz: float, fs: float) -> np.ndarray:
    """Filter design exactly as the reference requests it (devices.py:814, 1363)."""
    from scipy import signal as sg

    return sg.bessel(N=n, Wn=wn_hz, btype="low", fs=fs, output="sos", norm="mag")


def pad_edge(sos: np.ndarray) -> int:
    S = sos.shape[0]
    ntaps = 2 * S + 1 - min(int((sos[:, 2] == 0).sum()), int((sos[:, 5] == 0).sum()))
    return 3 * ntaps


def steady_state(sos: np.ndarray) -> np.ndarray:
    """sosfilt_zi: S x 2 initial state for a unit step input."""
    S = sos.shape[0]
    zi = np.empty((S, 2))
    scale = 1.0
    for s in range(S):
        b, a = sos[s, :3], sos[s, 3:]
        m = np.array([[1.0 + a[1], -1.0], [a[2], 1.0]])
        rhs = b[1:] - a[1:] * b[0]
        zi[s] = scale * np.linalg.solve(m, rhs)
        scale *= b.sum() / a.sum()
    return zi


def _sosfilt_real(sos, x, zi):
    """In-place cascade over a 1-D float64 array; zi (S,2) is updated."""
    lib = _lib()
    if lib:
        assert x.dtype == np.float64 and zi.dtype == np.float64 and zi.flags.c_contiguous
        sosc = np.ascontiguousarray(sos, dtype=np.float64)
        lib.oracle_sosfilt_f64(sosc.ctypes.data, sos.shape[0], x.ctypes.data, x.shape[0],
                               x.strides[0] // 8, zi.ctypes.data)
        return
    for i in range(x.shape[0]):
        v = x[i]
        for s in range(sos.shape[0]):
            b0, b1, b2, _, a1, a2 = sos[s]
            y = b0 * v + zi[s, 0]
            zi[s, 0] = b1 * v - a1 * y + zi[s, 1]
            zi[s, 1] = b2 * v - a2 * y
            v = y
        x[i] = v


def _filtfilt_1d_real(sos, x, edge, zi):
    n = x.shape[0]
    ext = np.concatenate((2 * x[0] - x[edge:0:-1], x, 2 * x[-1] - x[-2:-edge - 2:-1])).astype(np.float64)
    st = zi * ext[0]
    _sosfilt_real(sos, ext, st)
    rev = ext[::-1].copy()
    st = zi * rev[0]
    _sosfilt_real(sos, rev, st)
    return rev[::-1][edge:edge + n].copy()


def oracle_sosfiltfilt(sos: np.ndarray, x: np.ndarray) -> np.ndarray:
    """Zero-phase filtering along the last axis; real or complex input."""
    sos = np.asarray(sos, dtype=np.float64)
    x = np.asarray(x)
    edge = pad_edge(sos)
    if x.shape[-1] <= edge:
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % edge)
    zi = steady_state(sos)
    flat = x.reshape(-1, x.shape[-1])
    cplx = np.iscomplexobj(x)
    out = np.empty(flat.shape, dtype=np.complex128 if cplx else np.float64)
    for r in range(flat.shape[0]):
        if cplx:
            re = _filtfilt_1d_real(sos, np.ascontiguousarray(flat[r].real, dtype=np.float64), edge, zi)
            im = _filtfilt_1d_real(sos, np.ascontiguousarray(flat[r].imag, dtype=np.float64), edge, zi)
            out[r] = re + 1j * im
        else:
            out[r] = _filtfilt_1d_real(sos, np.ascontiguousarray(flat[r], dtype=np.float64), edge, zi)
    return out.reshape(x.shape)


def oracle_lpf(signal, noise, bw, fs, n=4):
    """LPF (devices.py:1363-1368): real part of filtfilt; noise filtered separately."""
    sos = bessel_sos(n, bw, fs)
    s = oracle_sosfiltfilt(sos, signal).real
    nz = None if noise is None else oracle_sosfiltfilt(sos, noise).real
    return s, nz


def oracle_bpf(signal, noise, bw, fs, n=4):
    """BPF (devices.py:814-823): low-pass of the complex envelope at BW/2."""
    sos = bessel_sos(n, bw / 2, fs)
    s = oracle_sosfiltfilt(sos, signal)
    nz = None if noise is None else oracle_sosfiltfilt(sos, noise)
    return s, nz
