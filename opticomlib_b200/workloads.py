"""Synthetic PRBS-driven OOK waveforms for the BASELINE.json configurations.

Host-side NumPy input builders used by ``bench.py`` and the tests.  They are
*input generators*, not part of the accelerated path: the reference builds the
same inputs with ``PRBS -> DAC -> LASER -> MZM`` (reference
``examples/ook_transmission_fiber_simulation.py:27-39``); those devices are out
of scope, so the chain is restated here in a few lines so that
inputs can be regenerated on a GPU box where the reference is absent.
``tests/test_oracle_vs_reference.py`` checks these builders against the real
reference chain when it is importable.

Reference statements followed:
  PRBS   devices.py:134-175 (Fibonacci LFSR, taps table)
  DAC    devices.py:296-345 + utils.py:1918-1921 (gaussian pulse), utils.py:1975-1980 (upfir)
  LASER  devices.py:483 (CW field sqrt(P0))
  MZM    devices.py:762-768
  EDFA   devices.py:920-936 (gain and ASE power; x polarisation only here)
"""
from __future__ import annotations

import numpy as np
from scipy import signal as _sg

_TAPS = {7: (7, 6), 9: (9, 5), 11: (11, 9), 15: (15, 14), 20: (20, 3), 23: (23, 18), 31: (31, 28)}

PLANCK = 6.62607015e-34
C_LIGHT = 299792458.0


def prbs(order: int, length: int | None = None, seed: int | None = None) -> np.ndarray:
    """PRBS bits as uint8 (same LFSR convention as the reference PRBS)."""
    if order not in _TAPS:
        raise ValueError("order must be one of %s" % (sorted(_TAPS),))
    state = (1 << order) - 1 if seed is None else seed % (1 << order)
    if state == 0:
        state = 1
    n = (1 << order) - 1 if length is None else int(length)
    t1, t2 = _TAPS[order][0] - 1, _TAPS[order][1] - 1
    mask = (1 << order) - 1
    out = np.empty(n, dtype=np.uint8)
    for i in range(n):
        out[i] = state & 1
        fb = ((state >> t1) ^ (state >> t2)) & 1
        state = ((state << 1) | fb) & mask
    return out


def gaussian_drive(bits: np.ndarray, sps: int, vpp: float = 5.0, offset: float = -2.5) -> np.ndarray:
    """DAC(..., pulse_shape='gaussian', Vpp, offset): complex128 drive voltage."""
    nb = len(bits)
    span = max(4, nb - 4)
    t = np.linspace(-span / 2, span / 2, span * sps + 1)
    a = 2 * np.sqrt(np.log(2)) / 1
    pulse = np.exp(-(a * (1 + 1j * 0.0) * t) ** (2 * 1))
    up = np.zeros(nb * sps)
    up[sps // 2::sps] = bits
    x = _sg.fftconvolve(up, pulse, mode="same")
    return x * vpp + offset


def nrz_drive(bits: np.ndarray, sps: int, vpp: float = 5.0, offset: float = -2.5) -> np.ndarray:
    """Rectangular NRZ drive built directly (no FIR): used for the 2^26 config."""
    return np.repeat(bits.astype(np.float64), sps) * vpp + offset


def mzm_field(drive: np.ndarray, p0_dbm: float, bias: float = -2.5, vpi: float = 5.0,
              loss_db: float = 3.0, er_db: float = 26.0) -> np.ndarray:
    """CW laser of ``p0_dbm`` through a push-pull MZM driven by ``drive``."""
    carrier = np.sqrt(10 ** (p0_dbm / 10) * 1e-3)
    loss = 10 ** (-loss_db / 10)
    eta = 2 * (10 ** (-er_db / 10)) ** 0.5
    g = np.pi / 2 / vpi * (drive + bias)
    return carrier * (loss ** 0.5 * (np.cos(g) + 1j * eta / 2 * np.sin(g)))


def ook_field(order: int, nbits: int, sps: int, p0_dbm: float, shape: str = "gaussian") -> np.ndarray:
    """complex128[nbits*sps] OOK field at the fibre input."""
    bits = prbs(order, nbits)
    drive = gaussian_drive(bits, sps) if shape == "gaussian" else nrz_drive(bits, sps)
    return np.asarray(mzm_field(drive, p0_dbm), dtype=np.complex128)


def ase_rows(base: np.ndarray, rows, fs: float, gain_db: float = 10.0, nf_db: float = 5.0,
             wavelength: float = 1550e-9, seed0: int = 1000) -> np.ndarray:
    """Monte-Carlo rows sqrt(G)*base + n_b (config #3).

    ``rows`` is an iterable of row indices b; row b uses default_rng(seed0+b).
    """
    G = 10 ** (gain_db / 10)
    p_ase = 10 ** (nf_db / 10) * PLANCK * (C_LIGHT / wavelength) * (G - 1) * fs
    sig = np.sqrt(p_ase / 4)
    rows = list(rows)
    out = np.empty((len(rows), base.size), dtype=np.complex128)
    for i, b in enumerate(rows):
        rng = np.random.default_rng(seed0 + int(b))
        n = rng.standard_normal((2, base.size))
        out[i] = np.sqrt(G) * base + sig * (n[0] + 1j * n[1])
    return out


# ---- the five BASELINE.json configurations (fibre parameters + sampling) -------------

CONFIGS = {
    # name: dict(sps, R, nbits, prbs_order, p0_dbm, fibre kwargs)
    "cfg1": dict(sps=64, R=10e9, nbits=1024, order=7, p0_dbm=5.0,
                 fiber=dict(length=50.0, alpha=0.2, beta_2=-20.0, beta_3=0.0, gamma=2.0, phi_max=0.01)),
    "cfg2": dict(sps=64, R=10e9, nbits=16384, order=15, p0_dbm=20.0,
                 fiber=dict(length=100.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.01)),
    "cfg3": dict(sps=64, R=10e9, nbits=1024, order=7, p0_dbm=5.0, rows=4096, gain_db=10.0, nf_db=5.0,
                 fiber=dict(length=50.0, alpha=0.2, beta_2=-20.0, beta_3=0.0, gamma=2.0, phi_max=0.01)),
    "cfg4": dict(sps=64, R=10e9, nbits=4096, order=15, p0_dbm=0.0, rows=1024, spans=10,
                 fiber=dict(length=80.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, h=10.0)),
    "cfg5": dict(sps=64, R=10e9, nbits=2 ** 20, order=23, p0_dbm=0.0,
                 fiber=dict(length=100.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, h=1.0)),
}


def config_input(name: str, shape: str | None = None) -> tuple[np.ndarray, float, dict]:
    """(field complex128[N], dt seconds, fibre kwargs) for a BASELINE config."""
    c = CONFIGS[name]
    fs = c["R"] * c["sps"]
    if shape is None:
        shape = "nrz" if name == "cfg5" else "gaussian"
    field = ook_field(c["order"], c["nbits"], c["sps"], c["p0_dbm"], shape)
    return field, 1.0 / fs, dict(c["fiber"])


# ---- BASELINE config #4: received frames for digital back-propagation ------------------------------
def cfg4_link(frames, fields_fn, spans: int = 10, gain_db: float = 16.0, nf_db: float = 5.0, seed0: int = 2000):
    """Forward link of config #4: `spans` x [80 km SSMF -> 16 dB gain -> ASE], run by
    ``fields_fn(block, **fiber_kwargs) -> block`` (the accelerated engine; the CPU reference would need days at
    full size).  ``frames`` is a complex128 array [B, N] of transmitted frames; row b gets ASE from
    default_rng(seed0 + b).  Returns the received frames [B, N]."""
    c = CONFIGS["cfg4"]
    fs = c["R"] * c["sps"]
    fwd = dict(length=80.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, h=10.0)
    G = 10 ** (gain_db / 10)
    p_ase = 10 ** (nf_db / 10) * PLANCK * (C_LIGHT / 1550e-9) * (G - 1) * fs
    x = np.array(frames, dtype=np.complex128, copy=True)
    rngs = [np.random.default_rng(seed0 + b) for b in range(x.shape[0])]
    for _ in range(spans):
        x = np.asarray(fields_fn(x, **fwd)) * np.sqrt(G)
        for b, rng in enumerate(rngs):
            n = rng.standard_normal((2, x.shape[1]))
            x[b] += np.sqrt(p_ase / 4) * (n[0] + 1j * n[1])
    return x


CFG4_RX = dict(bpf_bw=40e9, lpf_bw=7.5e9, span_loss_db=16.0, spans=10,
               dbp=dict(length=80.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, h=10.0))
