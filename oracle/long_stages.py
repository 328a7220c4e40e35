"""NumPy model of the long-waveform STAGES (outer N0-point stage, inner N_l-point stage, controller).

TEST INFRASTRUCTURE ONLY (see oracle/ssfm_oracle.py): it is injected into
``opticomlib_b200.longwave.LongPlan`` by the CPU tests so that the sequencing, the layout exchange
(``all_to_all_single``) and the max all-reduce of the N>1 path run under gloo, and it is an independent
restatement of the N = N0 x N_l decomposition (natural-order numpy FFTs, no transposed spectra) that the
CUDA stages are compared with.  The statements split here are opticomlib/devices.py:1155-1196:

    sample n = na N_l + nb,   bin k = ka + N0 kb
    outer "open" : A *= exp(j (h/2) gamma |A|^2)   (1175-1177);  N0-point DFT over na;  A *= W_N^{nb ka}
    inner        : N_l-point DFT over nb;  A *= exp(j imag(D~(w_k)) h)   (1179);  inverse (unscaled)
    outer "close": A *= conj W;  inverse N0-point DFT over ka;  A *= exp(-alpha/2 h) / N  (1179-1180);
                   A *= exp(j (h/2) gamma |A_start|^2)   (1181);  max |A|^2  (1194)
"""
from __future__ import annotations

import math

import numpy as np


class _State:
    def __init__(self, steps, z, h, done, log):
        self.steps = np.array([steps], np.int32); self.z = np.array([z], np.float64)
        self.h_next = np.array([h], np.float64); self.done = np.array([done], bool); self.h_log = log


class NumpyStages:
    def __init__(self, n_global, n_outer, ranks, rank, real):
        self.N, self.N0, self.G, self.g = int(n_global), int(n_outer), int(ranks), int(rank)
        self.Nl = self.N // self.N0
        self.cols, self.rows = self.Nl // self.G, self.N0 // self.G
        self.R = np.float32 if real in (np.float32, "fp32") else np.float64
        self.C = np.complex64 if self.R is np.float32 else np.complex128
        nb = np.arange(self.g * self.cols, (self.g + 1) * self.cols)            # my global columns
        ka = np.arange(self.N0)
        self.tw = np.exp(-2j * np.pi * ((ka[:, None] * nb[None, :]) % self.N) / self.N).astype(self.C)

    # ---- the C-ABI surface of CudaStages --------------------------------------------------------
    def begin(self, field, prm):
        R = self.R
        self.prm = prm
        self.a_lin = R(prm.alpha_db_km / 4.343)
        self.b2, self.b3, self.gm = R(prm.beta2_ps2_km), R(prm.beta3_ps3_km), R(prm.gamma_w_km)
        self.L, self.phi = R(prm.length_km), R(prm.phi_max_rad)
        self.fixed = not math.isnan(prm.h_km)
        self.single = (not self.fixed) and ((self.b2 == 0 and self.b3 == 0) or self.gm == 0)
        self._pmax = R(0)
        if not self.fixed and not self.single:
            self._pmax = R((np.abs(field.numpy()) ** 2).max())
        # my bins: rows ka = g*rows .. ; k = ka + N0 kb  -> imag(D~) on the fftfreq grid (typing.py:1641, devices.py:1144-1145)
        ka = np.arange(self.g * self.rows, (self.g + 1) * self.rows)
        k = ka[:, None] + self.N0 * np.arange(self.Nl)[None, :]
        k = np.where(k < self.N // 2, k, k - self.N)
        w = (k * (((1.0 / (self.N * prm.dt_s)) * 2.0) * np.pi * 1e-12)).astype(R)
        self.dim = (R(0.5) * self.b2) * w ** 2 + (R(1.0 / 6.0) * self.b3) * w ** 3
        self.z, self.h, self.steps, self.done, self.log = R(0), R(0), 0, False, []

    def pmax(self, value=None):
        if value is None:
            return float(self._pmax)
        self._pmax = self.R(value)
        return float(value)

    def ctrl(self, init):
        R = self.R
        with np.errstate(all="ignore"):
            if init:
                h = R(self.prm.h_km) if self.fixed else (self.L if self.single else R(self.phi / (np.abs(self.gm) * self._pmax)))
                self.h = self.L if self.L < h else h
                self.z, self.steps, self.done = R(0), 0, not (R(0) < self.L)
            else:
                self.log.append(float(self.h))
                self.z = R(self.z + self.h)
                hn = self.h if self.fixed else R(self.phi / (np.abs(self.gm) * self._pmax))
                rem = R(self.L - self.z)
                self.h = rem if rem < hn else hn
                self.steps += 1
                self.done = not (self.z < self.L)
            self._pmax = R(0)

    def outer(self, field, stage):
        A = field.numpy()
        if stage in (1, 2):                                                      # close
            # numpy's ifft divides by N0; the inner stage left its inverse unscaled, so N0/N remains
            A[:] = (np.fft.ifft(A * np.conj(self.tw), axis=0) * (self.N0 / self.N) * np.exp(-self.a_lin / 2 * self.h)).astype(self.C)
            A *= np.exp(1j * self.stash).astype(self.C)
            self._pmax = max(self._pmax, self.R((np.abs(A) ** 2).max()))
            if stage == 1:
                self.ctrl(False)                                                 # fixed step: the controller advances alone
                if self.done:
                    return
        if stage in (0, 1):                                                      # open
            self.stash = ((self.h / 2) * (self.gm * np.abs(A) ** 2)).astype(self.R)
            A *= np.exp(1j * self.stash).astype(self.C)
            A[:] = (np.fft.fft(A, axis=0) * self.tw).astype(self.C)

    def inner(self, rows):
        X = rows.numpy()
        X[:] = (np.fft.ifft(np.fft.fft(X, axis=1) * np.exp(1j * (self.dim * self.h)), axis=1) * self.Nl).astype(self.C)

    def sync(self):
        pass

    def state(self, want_log=False):
        log = np.array([self.log]) if want_log and self.log else None
        return _State(self.steps, float(self.z), float(self.h), self.done, log)

    def close(self):
        pass
