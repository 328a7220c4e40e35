"""BASELINE config #4 receiver chain on the GPU path vs the CPU oracles (pytest -m gpu):

    BPF(40 GHz) -> 10 x [ x 10^(-16/20) ; DBP(80 km, h = 10 km) ] -> |.|^2 -> LPF(7.5 GHz)

on frames that went through 10 x (80 km + 16 dB + ASE).  Parity on a few frames at reduced length and on
one frame at the full 2^18 samples; every stage is compared, not only the end of the chain."""
import numpy as np
import pytest

from oracle.filtfilt_oracle import oracle_bpf, oracle_lpf
from oracle.ssfm_oracle import oracle_dbp, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import opticomlib_b200 as ob
    return ob


def _receiver_gpu(ob, rx, dt, fs, precision):
    from opticomlib_b200 import workloads as wl
    from oracle.filtfilt_oracle import bessel_sos
    c = wl.CFG4_RX
    y = ob.filtfilt_batch(rx, bessel_sos(4, c["bpf_bw"] / 2, fs))                  # BPF (devices.py:814: Wn = BW/2)
    stages = [y.copy()]
    for _ in range(c["spans"]):
        y, info = ob.dbp_batch(y * 10 ** (-c["span_loss_db"] / 20), dt, precision=precision, **c["dbp"])
        assert (info.steps == 8).all()
    stages.append(np.asarray(y))
    p = np.abs(np.asarray(y, dtype=np.complex128)) ** 2
    out = ob.filtfilt_batch(p, bessel_sos(4, c["lpf_bw"], fs))                        # LPF (devices.py:1363: Wn = BW)
    return stages, out


def _receiver_oracle(rx_row, dt, fs, real):
    from opticomlib_b200 import workloads as wl
    c = wl.CFG4_RX
    y, _ = oracle_bpf(rx_row, None, c["bpf_bw"], fs)
    stages = [y.copy()]
    for _ in range(c["spans"]):
        y = oracle_dbp(y * 10 ** (-c["span_loss_db"] / 20), dt, real=real, **c["dbp"])["out"]
    stages.append(y)
    out, _ = oracle_lpf(np.abs(y.astype(np.complex128)) ** 2, None, c["lpf_bw"], fs)
    return stages, out


@pytest.mark.parametrize("nbits,frames,precision", [(256, 3, "fp64"), (256, 3, "fp32"), (4096, 1, "fp64")])
def test_cfg4_receiver_chain(ob, nbits, frames, precision):
    from opticomlib_b200 import workloads as wl
    c = wl.CONFIGS["cfg4"]
    fs = c["R"] * c["sps"]
    dt = 1.0 / fs
    base = wl.ook_field(15, nbits, c["sps"], c["p0_dbm"])                             # N = 2^14 or 2^18
    tx = np.stack([np.roll(base, 977 * b) for b in range(frames)])
    rx = wl.cfg4_link(tx, lambda blk, **kw: ob.fiber_batch(blk, dt, precision="fp64", **kw)[0])
    assert rx.shape == tx.shape and np.isfinite(rx).all()
    st_g, out_g = _receiver_gpu(ob, rx, dt, fs, precision)
    real = np.float64 if precision == "fp64" else np.float32
    tol = 1e-10 if precision == "fp64" else 1e-4
    for b in range(frames):
        st_o, out_o = _receiver_oracle(rx[b], dt, fs, real)
        assert rel_l2(st_g[0][b], st_o[0]) <= 1e-10                                   # BPF
        assert rel_l2(st_g[1][b], st_o[1]) <= tol                                      # after 10 DBP spans
        assert rel_l2(out_g[b], out_o) <= (1e-10 if precision == "fp64" else 2e-4)     # detected + LPF
    # DBP undid most of the deterministic distortion: closer to the transmitted field than the raw reception
    g_tot = 10 ** ((16.0 - 0.2 * 80.0) * 10 / 20)
    err_rx = rel_l2(rx / g_tot, tx)
    err_dbp = rel_l2(np.asarray(st_g[1], dtype=np.complex128), tx)
    assert err_dbp < err_rx
