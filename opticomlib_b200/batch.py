"""Batch-aware signal containers: ``[B, (P,) N]`` fields that stay on the device between devices.

The reference moves one ``optical_signal`` / ``electrical_signal`` (NumPy, host) from device function to device function
(typing.py:1111-1165, 2124-2196).  For Monte-Carlo batches that means a host round trip per stage; these two containers hold
the same pair (signal, noise) for B independent rows as CUDA tensors and expose the hot-path stages as methods, so a chain

    optical_batch.from_signal(tx).edfa(G=10, NF=5, rows=4096).fiber(length=50, ...).pd(BW=7.5e9, sample_stride=gv.sps)

touches the host only at its two ends.  Arithmetic stays in the kernels behind the C-ABI (engine / devices); the containers
only carry tensors.  ``to_signals()`` hands back ordinary per-row signal objects.
"""
from __future__ import annotations

import numpy as np

from . import devices, engine
from .typing import electrical_signal, gv, optical_signal


def _cuda_c128(a, device=None):
    torch = engine._torch()
    dev = engine.require_cuda(a.device if torch.is_tensor(a) and a.is_cuda else device)
    t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
    return t.to(dev).to(torch.complex128).contiguous()


class optical_batch:
    """B optical fields ``signal[B, N]`` or ``signal[B, P, N]`` (complex, CUDA) with optional ``noise`` of the same shape."""

    def __init__(self, signal, noise=None, *, device=None):
        self.signal = _cuda_c128(signal, device)
        if self.signal.ndim not in (2, 3):
            raise ValueError("signal must have shape [B, N] or [B, P, N]")
        self.noise = None if noise is None else _cuda_c128(noise, self.signal.device)
        if self.noise is not None and self.noise.shape != self.signal.shape:
            raise ValueError("noise must have the shape of signal")
        self.ssfm_info = None

    # ---- shape -----------------------------------------------------------------------------------------
    @property
    def batch(self):
        return self.signal.shape[0]

    @property
    def n_pol(self):
        return 1 if self.signal.ndim == 2 else self.signal.shape[1]

    @property
    def size(self):
        return self.signal.shape[-1]

    @classmethod
    def from_signal(cls, sig, rows=1, device=None):
        """One reference-style ``optical_signal`` repeated ``rows`` times."""
        s = np.asarray(sig.signal)[None].repeat(rows, axis=0)
        nz = None if devices._is_null(sig.noise) else np.asarray(sig.noise)[None].repeat(rows, axis=0)
        return cls(s, nz, device=device)

    def field(self):
        """signal + noise: what FIBER propagates (typing.py:1593-1597)."""
        return self.signal if self.noise is None else self.signal + self.noise

    def to_signals(self):
        s = self.signal.cpu().numpy()
        nz = None if self.noise is None else self.noise.cpu().numpy()
        return [optical_signal(s[b]) if nz is None else optical_signal(s[b], nz[b]) for b in range(self.batch)]

    # ---- stages ----------------------------------------------------------------------------------------
    def edfa(self, G, NF, rows=None, n_pol_out=None, seed=0):
        """EDFA gain + ASE (devices.py:921-936): the generated ASE joins ``noise``; a batch of one waveform can be amplified
        into ``rows`` independent realisations.  ``n_pol_out`` defaults to the input's polarisation count (the reference
        always returns two: pass 2)."""
        torch = engine._torch()
        rows = self.batch if rows is None else int(rows)
        if self.batch not in (1, rows):
            raise ValueError("rows must equal the batch size unless the batch holds one waveform")
        P, N = self.n_pol, self.size
        n_pol_out = P if n_pol_out is None else int(n_pol_out)
        g = float(np.sqrt(10 ** (G / 10)))
        src = (self.signal[0] if self.batch == 1 else self.signal).contiguous()
        amp = devices.edfa_batch(src, rows, G, NF, n_pol_out=n_pol_out, seed=seed).reshape(rows, n_pol_out, N)
        sig = torch.zeros_like(amp)
        sig[:, :P] = self.signal.reshape(self.batch, P, N) * g          # (broadcasts a batch of one)
        noise = amp - sig                                               # the generated ASE ...
        if self.noise is not None:
            noise[:, :P] += self.noise.reshape(self.batch, P, N) * g    # ... plus the amplified input noise
        shape = (rows, N) if (self.signal.ndim == 2 and n_pol_out == 1) else (rows, n_pol_out, N)
        return optical_batch(sig.reshape(shape), noise.reshape(shape))

    def fiber(self, length, alpha=0.0, beta_2=0.0, beta_3=0.0, gamma=0.0, phi_max=0.01, h=None, precision="fp64", dt=None):
        """FIBER (devices.py:1038-1206) on every row; the result has no separate noise, like the reference's."""
        out, info = devices.fiber_batch(self.field(), gv.dt if dt is None else dt, length, alpha, beta_2, beta_3, gamma, phi_max, h,
                                        precision=precision)
        res = optical_batch(out)
        res.ssfm_info = info
        return res

    def dbp(self, length, alpha=0.0, beta_2=0.0, beta_3=0.0, gamma=0.0, phi_max=0.01, h=None, precision="fp64", dt=None):
        return self.fiber(length, -alpha, -beta_2, -beta_3, -gamma, phi_max, h, precision, dt)

    def bpf(self, BW, n=4, fs=None):
        """BPF (devices.py:788-826): signal and noise filtered separately."""
        sos = devices._bessel_sos(n, BW / 2, gv.fs if fs is None else fs)
        return optical_batch(engine.filtfilt_sos(self.signal, sos), None if self.noise is None else engine.filtfilt_sos(self.noise, sos))

    def pd(self, BW, r=1.0, R_load=50.0, i_dark=0.0, extra_noise=None, sample_offset=0, sample_stride=1, n=4, fs=None):
        """PD square law -> LPF -> SAMPLER in one pass (devices.py:1514-1552, 1871-1891); ``extra_noise[B, N]`` = thermal + shot
        current samples (``engine.gaussian_noise``), added to the noise row."""
        sos = devices._bessel_sos(n, BW, gv.fs if fs is None else fs)
        s, nz = engine.pd_lpf(self.signal, sos, self.noise, extra_noise, r, R_load, i_dark, sample_offset, sample_stride)
        return electrical_batch(s, nz)

    def psd(self, nperseg=None, fs=None):
        """(f, psd[B, (P,) nperseg]) like utils.get_psd (utils.py:2048-2079) for every row."""
        nper = min(2048, self.size) if nperseg is None else nperseg
        f = np.fft.fftshift(np.fft.fftfreq(nper, 1.0 / (gv.fs if fs is None else fs)))
        return f, engine.welch_psd(self.signal, nper)


class electrical_batch:
    """B electrical signals ``signal[B, N]`` (float64, CUDA) with optional ``noise``."""

    def __init__(self, signal, noise=None, *, device=None):
        torch = engine._torch()
        dev = engine.require_cuda(signal.device if torch.is_tensor(signal) and signal.is_cuda else device)
        to = lambda a: (a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))).to(dev).to(torch.float64).contiguous()
        self.signal = to(signal)
        self.noise = None if noise is None else to(noise)

    @property
    def batch(self):
        return self.signal.shape[0]

    @property
    def size(self):
        return self.signal.shape[-1]

    def to_signals(self):
        s = self.signal.cpu().numpy()
        nz = None if self.noise is None else self.noise.cpu().numpy()
        return [electrical_signal(s[b]) if nz is None else electrical_signal(s[b], nz[b]) for b in range(self.batch)]

    def lpf(self, BW, n=4, fs=None):
        """LPF (devices.py:1286-1375): signal and noise travel as re / im of one complex row, filtered separately."""
        torch = engine._torch()
        sos = devices._bessel_sos(n, BW, gv.fs if fs is None else fs)
        packed = torch.complex(self.signal, self.noise if self.noise is not None else torch.zeros_like(self.signal))
        y = engine.filtfilt_sos(packed, sos)
        return electrical_batch(y.real.contiguous(), None if self.noise is None else y.imag.contiguous())

    def psd(self, nperseg=None, fs=None):
        torch = engine._torch()
        nper = min(2048, self.size) if nperseg is None else nperseg
        f = np.fft.fftshift(np.fft.fftfreq(nper, 1.0 / (gv.fs if fs is None else fs)))
        return f, engine.welch_psd(self.signal.to(torch.complex128), nper)
