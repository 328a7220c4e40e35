"""Live check of the oracle and the input builders against the UNMODIFIED reference.

Runs only where the reference tree is present (the build container); skipped on
the GPU box, where tests/golden/*.npz carry the same pins.
"""
import numpy as np
import pytest

from oracle.ref_shim import reference_root

pytestmark = pytest.mark.skipif(reference_root() is None, reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    from oracle.ref_shim import import_reference

    import_reference()
    import opticomlib.devices as dv
    from opticomlib import gv, optical_signal, electrical_signal

    return dict(dv=dv, gv=gv, optical_signal=optical_signal, electrical_signal=electrical_signal)


def _tx(ref, sps, nbits, order, p0):
    dv, gv = ref["dv"], ref["gv"]
    gv(sps=sps, R=10e9, N=nbits)
    v = dv.DAC(dv.PRBS(order=order, len=nbits), Vpp=5, offset=-2.5, pulse_shape="gaussian")
    return dv.MZM(dv.LASER(P0=p0), v, bias=-2.5, Vpi=5, loss_dB=3, ER_dB=26)


def test_input_builder_is_the_reference_tx_chain(ref):
    from opticomlib_b200 import workloads as wl

    s = _tx(ref, 16, 256, 7, 5.0)
    mine = wl.ook_field(7, 256, 16, 5.0)
    assert np.array_equal(mine, s.signal)
    bits = ref["dv"].PRBS(order=15, len=500).data
    assert np.array_equal(wl.prbs(15, 500), bits)


@pytest.mark.parametrize("kw", [
    dict(length=20.0, alpha=0.2, beta_2=-20.0, gamma=2.0),
    dict(length=12.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.005),
    dict(length=50.0, alpha=0.2, beta_2=-20.0, gamma=2.0, h=0.7),
])
def test_f32_oracle_is_bit_identical_to_reference(ref, kw):
    from oracle.ssfm_oracle import oracle_fiber, oracle_dbp

    s = _tx(ref, 16, 128, 7, 9.0)
    for fn_ref, fn_or in ((ref["dv"].FIBER, oracle_fiber), (ref["dv"].DBP, oracle_dbp)):
        out = fn_ref(s, **kw)
        z, _ = fn_ref(s, return_steps=True, **kw)
        o = fn_or(s.signal, ref["gv"].dt, real=np.float32, **kw)
        assert np.array_equal(o["out"], out.signal)
        assert np.array_equal(o["z"].astype(np.float64), z[1:])


def test_filter_oracles_against_reference(ref):
    from oracle.filtfilt_oracle import oracle_lpf, oracle_bpf
    from oracle.ssfm_oracle import rel_l2

    gv = ref["gv"]
    gv(sps=32, R=10e9)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(3000)
    o = ref["dv"].LPF(ref["electrical_signal"](x), BW=9e9, n=4)
    s, _ = oracle_lpf(x, None, 9e9, gv.fs, 4)
    assert rel_l2(s, o.signal) < 1e-13
    xc = rng.standard_normal((2, 3000)) + 1j * rng.standard_normal((2, 3000))
    o = ref["dv"].BPF(ref["optical_signal"](xc), BW=30e9, n=3)
    s, _ = oracle_bpf(xc, None, 30e9, gv.fs, 3)
    assert rel_l2(s, o.signal) < 1e-13
