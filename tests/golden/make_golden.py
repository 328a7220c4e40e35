"""Generate the golden fixtures by executing the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

Writes ``tests/golden/*.npz``.  Every array comes from the reference's own
``FIBER`` / ``DBP`` / ``LPF`` / ``BPF`` /
``animated_fiber_propagation_with_phase`` executed under this image's
NumPy 2.3.5 / SciPy 1.18.1 (the result is NumPy-version
dependent; these pins are for NumPy >= 2).  Inputs are stored with the outputs
so the fixtures are self-contained on a GPU box without the reference.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_shim import import_reference  # noqa: E402

import_reference()
from opticomlib import gv, optical_signal, electrical_signal  # noqa: E402
from opticomlib.devices import (  # noqa: E402
    FIBER, DBP, LPF, BPF, PRBS, DAC, LASER, MZM, animated_fiber_propagation_with_phase,
)


def tx(sps, nbits, order, p0):
    gv(sps=sps, R=10e9, N=nbits)
    bits = PRBS(order=order, len=nbits)
    v = DAC(bits, Vpp=5, offset=-2.5, pulse_shape="gaussian")
    return MZM(LASER(P0=p0), v, bias=-2.5, Vpi=5, loss_dB=3, ER_dB=26)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("%-28s %8.1f KiB" % (name, os.path.getsize(path) / 1024))


def fiber_case(name, sig, fn=FIBER, **kw):
    """Run the reference twice: plain call (output object) and return_steps (z log)."""
    out = fn(sig, **kw)
    z, traj = fn(sig, return_steps=True, **kw)
    assert np.array_equal(traj[-1], out.signal)
    save(name,
         x=sig.to_numpy(), dt=np.float64(gv.dt), out=out.signal, z=z,
         kw_names=np.array(sorted(kw)), kw_vals=np.array([np.nan if kw[k] is None else kw[k] for k in sorted(kw)], dtype=np.float64),
         is_dbp=np.bool_(fn is DBP))


def nonpow2():
    """Lengths that are not powers of two (the reference accepts any N; added later, own RNG stream)."""
    rng = np.random.default_rng(20261018)
    s = tx(10, 127, 7, 10.0)   # N = 1270
    fiber_case("fiber_n1270_adaptive", s, length=20.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0)
    s = tx(12, 250, 9, 8.0)    # N = 3000
    fiber_case("fiber_n3000_fixed", s, length=12.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, h=0.5)
    s = tx(8, 125, 7, 10.0)    # N = 1000, two polarisations + noise, odd-symmetric check of the bin map via DBP
    sig2 = np.stack([s.signal, 0.5j * s.signal[::-1]])
    noi2 = 1e-3 * (rng.standard_normal(sig2.shape) + 1j * rng.standard_normal(sig2.shape))
    fiber_case("dbp_2pol_n1000", optical_signal(sig2, noi2), fn=DBP, length=8.0, alpha=0.2, beta_2=-20.0, gamma=1.5)
    s = tx(9, 111, 7, 10.0)    # N = 999 (odd)
    fiber_case("fiber_n999_odd", s, length=10.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, phi_max=0.02)


def dm():
    """DM (devices.py:941-1035): two polarisations + noise, and a plain single-polarisation field."""
    from opticomlib.devices import DM
    rng = np.random.default_rng(20261019)
    s = tx(16, 256, 9, 10.0)   # N = 4096
    sig2 = np.stack([s.signal, 0.5j * s.signal[::-1]])
    noi2 = 1e-3 * (rng.standard_normal(sig2.shape) + 1j * rng.standard_normal(sig2.shape))
    o, H = DM(optical_signal(sig2, noi2), D=4000.0, retH=True)
    save("dm_2pol_noise_4096", x=sig2, xn=noi2, dt=np.float64(gv.dt), D=np.float64(4000.0), out=o.signal, outn=o.noise, H=H)
    o = DM(optical_signal(s.signal), D=-1700.0)
    save("dm_1pol_4096", x=s.signal, dt=np.float64(gv.dt), D=np.float64(-1700.0), out=o.signal)


def main():
    if "nonpow2" in sys.argv:
        return nonpow2()
    if "dm" in sys.argv:
        return dm()
    rng = np.random.default_rng(20261017)

    # --- FIBER / DBP, shipped float32 path ------------------------------------------------
    s = tx(16, 256, 7, 10.0)  # N = 4096
    fiber_case("fiber_adaptive_4096", s, length=20.0, alpha=0.2, beta_2=-20.0, gamma=2.0)
    fiber_case("fiber_beta3_4096", s, length=30.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.02)
    fiber_case("dbp_adaptive_4096", s, fn=DBP, length=10.0, alpha=0.2, beta_2=-20.0, gamma=2.0)
    fiber_case("fiber_gamma0_4096", s, length=40.0, alpha=0.2, beta_2=-20.0, gamma=0.0)
    fiber_case("fiber_nodisp_4096", s, length=5.0, alpha=0.2, gamma=2.0)
    fiber_case("fiber_alpha_only_4096", s, length=10.0, alpha=0.2)

    s = tx(16, 128, 7, 8.0)  # N = 2048
    fiber_case("fiber_fixed_h03_2048", s, length=50.0, alpha=0.2, beta_2=-20.0, gamma=2.0, h=0.3)
    s = tx(8, 128, 7, 3.0)  # N = 1024
    fiber_case("fiber_fixed_h01_1024", s, length=50.0, alpha=0.2, beta_2=-20.0, gamma=2.0, h=0.1)

    # two polarisations + noise merged into the field (typing.py:1596)
    s = tx(16, 256, 9, 10.0)
    sig2 = np.stack([s.signal, 0.5j * s.signal[::-1]])
    noi2 = 1e-3 * (rng.standard_normal(sig2.shape) + 1j * rng.standard_normal(sig2.shape))
    s2 = optical_signal(sig2, noi2)
    fiber_case("fiber_2pol_noise_4096", s2, length=15.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=1.5)
    fiber_case("dbp_2pol_fixed_4096", s2, fn=DBP, length=8.0, alpha=0.2, beta_2=-20.0, gamma=1.5, h=0.5)

    # BASELINE config #1 at full size: input is regenerated bit-exactly by
    # opticomlib_b200.workloads.config_input('cfg1'); keep a decimated output.
    s = tx(64, 1024, 7, 5.0)
    kw = dict(length=50.0, alpha=0.2, beta_2=-20.0, gamma=2.0)
    out = FIBER(s, **kw)
    z, _ = FIBER(s, return_steps=True, **kw)
    save("fiber_cfg1_65536", x_head=s.signal[:64], dt=np.float64(gv.dt), out_dec=out.signal[::16], z=z,
         out_abs2_sum=np.float64((np.abs(out.signal.astype(np.complex128)) ** 2).sum()))

    # --- float64: the reference's own float64 loop (devices.py:2440-2486) ----------------------
    s = tx(16, 256, 7, 10.0)
    kw = dict(length=20.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, phi_max=0.01)
    z, A_z, _, _ = animated_fiber_propagation_with_phase(s, plot=False, **kw)
    out64 = A_z[-1] * np.exp(-(kw["alpha"] / 4.343) * kw["length"] / 2)  # undo the display rescale (devices.py:2472)
    save("fiber_f64_4096", x=s.signal, dt=np.float64(gv.dt), out=out64, z=z,
         kw_names=np.array(sorted(kw)), kw_vals=np.array([kw[k] for k in sorted(kw)], dtype=np.float64))

    # --- LPF / BPF ---------------------------------------------------------------------
    gv(sps=16, R=1e9)
    e = electrical_signal(np.ones(100))  # reference tests/devices_test.py:279-284
    save("lpf_ones_100", x=e.signal, fs=np.float64(gv.fs), bw=np.float64(1e9), n=np.int64(4), out=LPF(e, BW=1e9).signal)

    gv(sps=64, R=10e9)
    for n in (3, 4, 5):
        xs = rng.standard_normal(4096) + 0.3
        xn = 0.1 * rng.standard_normal(4096)
        o = LPF(electrical_signal(xs, xn), BW=7.5e9, n=n)
        save("lpf_n%d_4096" % n, x=xs, xn=xn, fs=np.float64(gv.fs), bw=np.float64(7.5e9), n=np.int64(n),
             out=o.signal, outn=o.noise)
    xs = rng.standard_normal(1000) + 1j * rng.standard_normal(1000)  # complex in -> real out, N not a power of two
    o = LPF(electrical_signal(xs), BW=20e9)
    save("lpf_complex_1000", x=xs, fs=np.float64(gv.fs), bw=np.float64(20e9), n=np.int64(4), out=o.signal)

    xs = rng.standard_normal((2, 4096)) + 1j * rng.standard_normal((2, 4096))
    xn = 0.1 * (rng.standard_normal((2, 4096)) + 1j * rng.standard_normal((2, 4096)))
    o = BPF(optical_signal(xs, xn), BW=40e9)
    save("bpf_2pol_4096", x=xs, xn=xn, fs=np.float64(gv.fs), bw=np.float64(40e9), n=np.int64(4), out=o.signal, outn=o.noise)
    xs = rng.standard_normal(8192) + 1j * rng.standard_normal(8192)
    o = BPF(optical_signal(xs), BW=15e9, n=5)
    save("bpf_1pol_n5_8192", x=xs, fs=np.float64(gv.fs), bw=np.float64(15e9), n=np.int64(5), out=o.signal)


if __name__ == "__main__":
    main()
