"""GPU drop-in test: the reference's OWN modules, objects and call chains, with `install()` redirecting the hot path.

Needs the unmodified reference next to the repo (baseline/_ref on the GPU box, /root/reference in the build container);
skipped otherwise.  `opticomlib_b200.install()` rebinds `opticomlib.devices.{FIBER,DBP,LPF,BPF,DM}` (and the import-time
copies in `opticomlib.ook` / `opticomlib.ppm`); everything else -- `gv`, `optical_signal`, `electrical_signal`, `PRBS`,
`DAC`, `LASER`, `MZM`, `PD`, `EDFA` -- is the reference's own code.  Every case runs the reference call twice, once with
the original functions (NumPy on the host) and once installed (B200 kernels), on the same reference objects, and
compares results, return types and the attributes the reference sets.

Reference call sites covered: FIBER devices.py:1038-1206, DBP 1209-1283, LPF 1286-1375, BPF 788-826, DM 941-1035,
in-module callers DAC->LPF (347), MZM->BPF (780), PD->LPF (1552), and the example chain
examples/ook_transmission_fiber_simulation.py:27-45 (PRBS -> DAC -> LASER -> MZM -> FIBER -> PD).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


@pytest.fixture(scope="module")
def ref():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle.ref_shim import reference_root, import_reference
    if reference_root() is None:
        pytest.skip("reference tree not present (baseline/_ref)")
    import_reference()
    import opticomlib
    import opticomlib.devices as rdv
    import opticomlib.ook  # noqa: F401  (holds an import-time binding of LPF)
    return opticomlib, rdv


@pytest.fixture()
def installed(ref):
    from opticomlib_b200 import devices as dv
    dv.install()
    yield dv
    dv.uninstall()
    dv.DEFAULT_PRECISION = "fp32"


def _tx(opticomlib, rdv, order=7, nbits=256, sps=16, p0=5.0):
    opticomlib.gv(sps=sps, R=10e9, wavelength=1550e-9, Vpi=5, N=nbits)
    bits = rdv.PRBS(order=order, len=nbits)
    v = rdv.DAC(bits, Vpp=5, offset=-2.5, pulse_shape="gaussian")
    return bits, rdv.MZM(rdv.LASER(P0=p0), v, bias=-2.5, Vpi=5, loss_dB=3, ER_dB=26)


def test_fiber_and_dbp_through_the_reference_module(ref):
    opticomlib, rdv = ref
    from opticomlib_b200 import devices as dv
    _, sig = _tx(opticomlib, rdv)
    kw = dict(length=30, alpha=0.2, beta_2=-20, beta_3=0.1, gamma=2)
    want = rdv.FIBER(sig, **kw)
    want_z, want_traj = rdv.FIBER(sig, return_steps=True, **kw)
    want_back = rdv.DBP(want, **kw)
    dv.install()
    try:
        assert rdv.FIBER is dv.FIBER and rdv.DBP is dv.DBP
        got = rdv.FIBER(sig, **kw)
        got_z, got_traj = rdv.FIBER(sig, return_steps=True, **kw)
        got_back = rdv.DBP(want, **kw)
    finally:
        dv.uninstall()
    assert type(got) is opticomlib.optical_signal and got.signal.dtype == want.signal.dtype == np.complex64
    assert got.noise is opticomlib.NULL or type(got.noise).__name__ == type(want.noise).__name__
    assert got.n_pol == want.n_pol and got.size == want.size and got.execution_time is not None
    assert rel_l2(got.signal, want.signal) <= 1e-4
    assert got_z.dtype == want_z.dtype and got_traj.shape == want_traj.shape
    assert got_z.shape == want_z.shape                      # identical step count; adaptive step sizes agree to float32 rounding
    np.testing.assert_allclose(got_z, want_z, rtol=2e-6)
    assert rel_l2(got_traj, want_traj) <= 1e-4
    assert type(got_back) is opticomlib.optical_signal
    assert rel_l2(got_back.signal, want_back.signal) <= 1e-4


def test_two_polarisations_with_noise(ref, installed):
    opticomlib, rdv = ref
    opticomlib.gv(sps=16, R=10e9, N=256)
    rng = np.random.default_rng(2)
    n = 4096
    s = np.sqrt(2e-3) * np.exp(2j * np.pi * rng.random((2, n)) * 0.05) * (0.6 + 0.4 * np.sign(np.sin(np.arange(n) / 40.0)))
    nz = 1e-4 * (rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n)))
    sig = opticomlib.optical_signal(s, nz)
    kw = dict(length=20, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.02)
    installed.uninstall()
    want = rdv.FIBER(sig, **kw)
    installed.install()
    got = rdv.FIBER(sig, **kw)
    assert got.signal.shape == (2, n) and got.n_pol == 2
    assert rel_l2(got.signal, want.signal) <= 1e-4


def test_filters_and_in_module_callers(ref):
    """BPF / LPF / DM directly, and DAC -> LPF, MZM -> BPF, PD -> LPF through the reference's own functions."""
    opticomlib, rdv = ref
    from opticomlib_b200 import devices as dv
    bits, sig = _tx(opticomlib, rdv, nbits=128, sps=32)
    rng = np.random.default_rng(4)
    noisy = opticomlib.optical_signal(sig.signal, 1e-3 * (rng.standard_normal(sig.size) + 1j * rng.standard_normal(sig.size)))
    el = opticomlib.electrical_signal(np.abs(sig.signal) ** 2, 1e-4 * rng.standard_normal(sig.size))

    def chain():
        out = {}
        out["bpf"] = rdv.BPF(noisy, BW=40e9)
        out["lpf"] = rdv.LPF(el, BW=7.5e9)
        out["lpf_h"] = rdv.LPF(el, BW=7.5e9, retH=True)[1]
        out["dm"] = rdv.DM(noisy, D=-300.0)
        out["dac"] = rdv.DAC(bits, Vpp=5, offset=-2.5, pulse_shape="nrz", BW=8e9)          # DAC -> LPF (devices.py:347)
        out["mzm"] = rdv.MZM(rdv.LASER(P0=3), out["dac"], bias=-2.5, Vpi=5, loss_dB=2, ER_dB=30, BW=30e9)   # MZM -> BPF (780)
        out["pd"] = rdv.PD(rdv.FIBER(sig, length=10, alpha=0.2, beta_2=-20, gamma=2), BW=7.5e9, r=1, include_noise="none")  # PD -> LPF (1552)
        return out

    want = chain()
    dv.install()
    try:
        assert rdv.LPF is dv.LPF and rdv.BPF is dv.BPF and rdv.DM is dv.DM
        import opticomlib.ook as rook
        assert rook.LPF is dv.LPF
        got = chain()
    finally:
        dv.uninstall()
    for key in ("bpf", "lpf", "dm", "dac", "mzm"):
        assert type(got[key]) is type(want[key]), key
        assert rel_l2(got[key].signal, want[key].signal) <= 1e-10, key
        if want[key].noise is not opticomlib.NULL:
            assert rel_l2(got[key].noise, want[key].noise) <= 1e-10, key
        else:
            assert got[key].noise is opticomlib.NULL, key
    np.testing.assert_allclose(got["lpf_h"], want["lpf_h"], rtol=1e-12, atol=1e-15)
    assert type(got["pd"]) is opticomlib.electrical_signal
    assert rel_l2(got["pd"].signal, want["pd"].signal) <= 2e-4     # FIBER in fp32 (as shipped) feeds the square law


def test_example_chain_ook_transmission(ref):
    """examples/ook_transmission_fiber_simulation.py:27-45 at its own sizes (2^16 samples), noise-free receiver."""
    opticomlib, rdv = ref
    from opticomlib_b200 import devices as dv
    opticomlib.gv(sps=64, R=10e9, wavelength=1550e-9, Vpi=5, N=2 ** 10)
    tx = rdv.PRBS(order=9, len=opticomlib.gv.N)
    v = rdv.DAC(tx, Vpp=opticomlib.gv.Vpi, offset=-opticomlib.gv.Vpi / 2, pulse_shape="gaussian")
    mod = rdv.MZM(rdv.LASER(P0=5), v, bias=-opticomlib.gv.Vpi / 2, Vpi=opticomlib.gv.Vpi, loss_dB=3, ER_dB=26)

    def rx():
        f = rdv.FIBER(mod, length=50, alpha=0.2, beta_2=-20, gamma=2)
        return f, rdv.PD(f, BW=opticomlib.gv.R * 0.75, r=1, include_noise="none")

    want_f, want_pd = rx()
    dv.install()
    try:
        got_f, got_pd = rx()
    finally:
        dv.uninstall()
    assert got_f.size == 2 ** 16 and int(got_f.ssfm_info.steps[0]) == 9        # 9 steps with PRBS9
    assert rel_l2(got_f.signal, want_f.signal) <= 1e-4
    assert abs(got_f.power("dBm") - want_f.power("dBm")) <= 1e-3
    assert rel_l2(got_pd.signal, want_pd.signal) <= 2e-4


@pytest.mark.parametrize("mode", ["all", "none", "ase-only", "thermal-shot", "ase-shot"])
def test_pd_drop_in_with_seeded_noise(ref, mode):
    """PD (devices.py:1377-1556) redirected as a whole: the thermal and shot samples come from np.random.normal in the
    reference's order, so with the same seed the two implementations must agree sample for sample (signal AND noise)."""
    opticomlib, rdv = ref
    from opticomlib_b200 import devices as dv
    _, sig = _tx(opticomlib, rdv, nbits=128, sps=32)
    rng = np.random.default_rng(11)
    for pols in (1, 2):
        s = sig.signal if pols == 1 else np.stack([sig.signal, 0.5j * sig.signal])
        nz = 2e-3 * np.abs(sig.signal).max() * (rng.standard_normal(s.shape) + 1j * rng.standard_normal(s.shape))
        x = opticomlib.optical_signal(s, nz)
        np.random.seed(5)
        want = rdv.PD(x, BW=7.5e9, r=0.9, include_noise=mode, Fn=3)
        dv.install()
        try:
            assert rdv.PD is dv.PD
            np.random.seed(5)
            got = rdv.PD(x, BW=7.5e9, r=0.9, include_noise=mode, Fn=3)
        finally:
            dv.uninstall()
        assert type(got) is opticomlib.electrical_signal and got.signal.shape == want.signal.shape
        assert rel_l2(got.signal, want.signal) <= 1e-10
        if mode == "none":
            assert got.noise is opticomlib.NULL and want.noise is opticomlib.NULL
        else:
            assert rel_l2(got.noise, want.noise) <= 1e-10
    with pytest.raises(ValueError):
        dv.PD(x, BW=7.5e9, include_noise="bogus")
    with pytest.raises(ValueError):
        dv.PD(x, BW=7.5e9, r=1.5)


def test_errors_match_the_reference(ref, installed):
    opticomlib, rdv = ref
    with pytest.raises(TypeError):
        rdv.FIBER(np.ones(16, complex), length=1)
    with pytest.raises(TypeError):
        rdv.BPF(opticomlib.electrical_signal(np.ones(64)), BW=1e9)
    with pytest.raises(ValueError):
        rdv.LPF(opticomlib.optical_signal(np.ones((2, 64), complex)), BW=1e9)
