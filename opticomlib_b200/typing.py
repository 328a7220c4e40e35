"""Signal containers on the hot path: the surface of ``opticomlib/typing.py`` that
``FIBER`` / ``DBP`` / ``LPF`` / ``BPF`` touch (last paragraph).

Only what the path needs is here -- constructors (typing.py:1111-1165, 2124-2196), ``signal`` /
``noise`` / ``n_pol`` / ``size`` / ``ndim`` / ``type``, ``to_numpy`` (1593-1597), ``w`` (1628-1644),
slicing copy (1366-1371, 2278-2291), ``execution_time`` and the ``NULL`` noise marker (56-93), and
the global sampling parameters ``gv`` (105-388: ``sps``, ``R``, ``fs``, ``dt``, ``N``, ``f0``).
Plotting, PSD, eye diagrams and operator sugar are out of scope.  The device functions also accept
the reference's own objects by duck typing (see ``devices._kind``).
"""
from __future__ import annotations

import numpy as np

__all__ = ["NULL", "gv", "electrical_signal", "optical_signal"]


class _NullType:
    """Absent noise: ``x + NULL -> x`` (typing.py:56-93)."""
    _inst = None

    def __new__(cls):
        if cls._inst is None:
            cls._inst = super().__new__(cls)
        return cls._inst

    def __add__(self, other):
        return other

    __radd__ = __add__

    def __mul__(self, other):
        return self

    __rmul__ = __mul__

    def __neg__(self):
        return self

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method == "__call__" and ufunc in (np.add, np.subtract) and inputs[1] is self:
            return inputs[0]
        return self

    def __repr__(self):
        return "NULL"


NULL = _NullType()


class _GlobalVars:
    """Sampling parameters shared by all devices (typing.py:105-388); read at call time."""

    def __init__(self):
        self.default()

    def default(self):
        self.sps, self.R = 16, 1e9
        self.fs = self.R * self.sps
        self.dt = 1 / self.fs
        self.wavelength = 1550e-9
        self.f0 = 299792458.0 / self.wavelength
        self.N = 128
        return self

    def __call__(self, sps=None, R=None, fs=None, wavelength=1550e-9, N=None, **custom):
        if sps:
            self.sps = int(np.round(sps))
            if R:
                self.R, self.fs = R, R * self.sps
            elif fs:
                self.fs, self.R = fs, fs / self.sps
            else:
                self.fs = self.R * self.sps
        elif R:
            self.R = R
            if fs:
                self.fs, self.sps = fs, int(np.round(fs / R))
            else:
                self.fs = R * self.sps
        elif fs:
            self.fs, self.sps = fs, int(np.round(fs / self.R))
        self.dt = 1 / self.fs
        self.N = N if N is not None else self.N
        self.wavelength = wavelength
        self.f0 = 299792458.0 / wavelength
        for k, v in custom.items():
            setattr(self, k, v)
        return self

    @property
    def t(self):
        return np.linspace(0, self.N * self.sps / self.fs, self.N * self.sps, endpoint=True)


gv = _GlobalVars()


def _as_arrays(signal, noise, dtype):
    sig = np.array(signal)
    if noise is NULL:
        return (sig.astype(dtype) if dtype is not None else sig), NULL
    noi = np.array(noise)
    common = np.result_type(sig, noi) if dtype is None else dtype
    sig, noi = sig.astype(common), noi.astype(common)
    if sig.shape != noi.shape:
        raise ValueError(f"`signal` and `noise` must have the same shape, mismatch shapes {sig.shape} and {noi.shape}!")
    return sig, noi


class electrical_signal:
    """1-D electrical waveform: ``signal`` plus optional ``noise`` (typing.py:1022-1165)."""

    def __init__(self, signal, noise=NULL, dtype=None):
        if type(self) is electrical_signal:
            if isinstance(signal, electrical_signal):
                extra = noise
                signal, noise = signal.signal, signal.noise
                if extra is not NULL:
                    noise = noise + np.array(extra)
            sig, noi = _as_arrays(signal, noise, dtype)
            if sig.ndim > 1 or sig.size < 1:
                raise ValueError(f"Signal must be scalar or 1D array for electrical_signal, invalid shape {sig.shape}")
            if sig.ndim == 0:
                sig = sig[np.newaxis]
                noi = noi if noi is NULL else noi[np.newaxis]
            signal, noise = sig, noi
        self.signal = signal
        self.noise = noise
        self.execution_time = 0.0

    # -- what the devices read ---------------------------------------------------------------
    @property
    def type(self):
        return type(self)

    @property
    def size(self):
        return self.signal.size

    @property
    def ndim(self):
        return self.signal.ndim

    @property
    def shape(self):
        return self.signal.shape

    @property
    def dtype(self):
        return self.to_numpy().dtype

    def to_numpy(self, dtype=None, copy=False):
        return np.array(self.signal + self.noise, dtype=dtype, copy=copy)

    def __array__(self, dtype=None, copy=None):
        arr = self.signal + self.noise
        return arr if dtype is None else arr.astype(dtype)

    def __len__(self):
        return self.size

    def w(self, shift=False):
        w = np.fft.fftfreq(self.size, gv.dt) * 2 * np.pi
        return np.fft.fftshift(w, axes=-1) if shift else w

    def __getitem__(self, key):
        if isinstance(key, slice):
            return self.__class__(self.signal[key]) if self.noise is NULL else self.__class__(self.signal[key], self.noise[key])
        if self.noise is NULL:
            return self.signal[key]
        return self.__class__(self.signal[key], self.noise[key])

    def __repr__(self):
        return f"{type(self).__name__}(signal={self.signal!r}, noise={self.noise!r})"


class optical_signal(electrical_signal):
    """Optical field, one or two polarisation rows (typing.py:2103-2196)."""

    def __init__(self, signal, noise=NULL, n_pol=None, dtype=None):
        if isinstance(signal, electrical_signal):
            extra = noise
            signal, noise = signal.signal, signal.noise
            if extra is not NULL:
                noise = noise + np.array(extra)
        sig, noi = _as_arrays(signal, noise, dtype)
        if sig.ndim > 2 or (sig.ndim > 1 and sig.shape[0] > 2) or sig.size < 1:
            raise ValueError(f"Signal must be a scalar, 1D or 2D array for optical_signal, invalid shape {sig.shape}")
        if n_pol is not None and n_pol not in (1, 2):
            raise ValueError("n_pol must be either 1 or 2")

        def both(f):
            return f(sig), (noi if noi is NULL else f(noi))

        if sig.ndim == 0:
            if n_pol in (None, 1):
                sig, noi = both(lambda a: a[np.newaxis]); n_pol = 1
            else:
                sig, noi = both(lambda a: np.array([[a], [a]]))
        elif sig.ndim == 1:
            if n_pol in (None, 1):
                n_pol = 1
            else:
                sig, noi = both(lambda a: np.array([a, a]))
        elif sig.shape[0] == 1:
            if n_pol in (None, 2):
                sig, noi = both(lambda a: np.tile(a, (2, 1))); n_pol = 2
            else:
                sig, noi = both(lambda a: a[0])
        else:
            if n_pol in (None, 2):
                n_pol = 2
            else:
                sig, noi = both(lambda a: a[0])
        self.n_pol = n_pol
        super().__init__(sig, noi)

    @property
    def size(self):
        return self.signal.size if self.n_pol == 1 else self.signal[0].size

    def __getitem__(self, key):
        if isinstance(key, slice):
            pick = (lambda a: a[key]) if self.n_pol == 1 else (lambda a: a[:, key])
            noi = NULL if self.noise is NULL else pick(self.noise)
            return self.__class__(pick(self.signal), noi, n_pol=self.n_pol)
        sig = self.signal[key]
        if self.noise is NULL:
            return sig if self.n_pol == 1 else self.__class__(sig, NULL, n_pol=1)
        return self.__class__(sig, self.noise[key], n_pol=1 if sig.ndim != 2 else self.n_pol)
