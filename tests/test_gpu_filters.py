"""GPU parity tests for LPF / BPF (pytest -m gpu): CUDA filtfilt vs golden reference outputs,
vs the CPU oracle and vs scipy.signal.sosfiltfilt itself.  Tolerance 1e-10 rel-L2 (BASELINE.json)."""
import numpy as np
import pytest

from conftest import golden
from oracle.filtfilt_oracle import oracle_sosfiltfilt, bessel_sos
from oracle.ssfm_oracle import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def ob():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import opticomlib_b200 as ob
    return ob


def _set_fs(ob, fs):
    ob.gv.fs = float(fs)
    ob.gv.dt = 1.0 / float(fs)


@pytest.mark.parametrize("name", ["lpf_ones_100", "lpf_n3_4096", "lpf_n4_4096", "lpf_n5_4096", "lpf_complex_1000"])
def test_lpf_matches_reference(ob, name):
    g = golden(name)
    _set_fs(ob, g["fs"])
    has_noise = "xn" in g.files
    inp = ob.electrical_signal(g["x"], g["xn"]) if has_noise else ob.electrical_signal(g["x"])
    out = ob.LPF(inp, BW=float(g["bw"]), n=int(g["n"]))
    assert out.type is ob.electrical_signal and out.size == inp.size
    assert out.signal.dtype == np.float64
    assert rel_l2(out.signal, g["out"]) <= TOL
    if has_noise:
        assert rel_l2(out.noise, g["outn"]) <= TOL
    else:
        assert out.noise is ob.NULL
    assert out.execution_time > 0


@pytest.mark.parametrize("name", ["bpf_2pol_4096", "bpf_1pol_n5_8192"])
def test_bpf_matches_reference(ob, name):
    g = golden(name)
    _set_fs(ob, g["fs"])
    has_noise = "xn" in g.files
    inp = ob.optical_signal(g["x"], g["xn"]) if has_noise else ob.optical_signal(g["x"])
    out = ob.BPF(inp, BW=float(g["bw"]), n=int(g["n"]))
    assert out.type is ob.optical_signal and out.size == inp.size and out.n_pol == inp.n_pol
    assert rel_l2(out.signal, g["out"]) <= TOL
    if has_noise:
        assert rel_l2(out.noise, g["outn"]) <= TOL


def test_lpf_fs_argument_ndarray_input_and_retH(ob):
    from scipy import signal as sg
    _set_fs(ob, 1e9)
    x = np.random.default_rng(3).standard_normal(777)
    out, H = ob.LPF(x, BW=2e9, n=4, fs=40e9, retH=True)
    sos = bessel_sos(4, 2e9, 40e9)
    assert rel_l2(out.signal, sg.sosfiltfilt(sos, x)) <= TOL
    _, Href = sg.sosfreqz(sos, worN=777, fs=40e9, whole=True)
    assert np.allclose(H, np.fft.fftshift(Href))
    assert out.execution_time == 0.0                                    # reference returns before toc() (devices.py:1370-1372)


@pytest.mark.parametrize("n_samples,rows", [(1 << 16, 8), (1 << 18, 4), (100000, 3)])
def test_filtfilt_batch_large_rows_vs_oracle_and_scipy(ob, n_samples, rows):
    from scipy import signal as sg
    rng = np.random.default_rng(n_samples)
    x = rng.standard_normal((rows, n_samples)) + 1j * rng.standard_normal((rows, n_samples))
    x[0] += 3.0                                                         # non-zero mean exercises the zi*x0 start-up
    sos = bessel_sos(4, 20e9, 640e9)
    y = ob.filtfilt_batch(x, sos)
    assert rel_l2(y, sg.sosfiltfilt(sos, x, axis=-1)) <= TOL
    assert rel_l2(y[:2], oracle_sosfiltfilt(sos, x[:2])) <= TOL
    # linearity (size-independent property)
    y2 = ob.filtfilt_batch(2.5 * x + 1.0, sos)
    dc = ob.filtfilt_batch(np.ones((1, n_samples), complex), sos)
    assert rel_l2(y2, 2.5 * y + dc) <= 1e-9


def test_fft_path_and_sequential_path_agree(ob, monkeypatch):
    """2^16-sample rows take the FFT path (circular |H|^2 + exact end segments); forcing the sequential
    recursion must give the same samples, including the very first and last ones."""
    from scipy import signal as sg
    rng = np.random.default_rng(11)
    x = rng.standard_normal((3, 1 << 16)) + 1j * rng.standard_normal((3, 1 << 16)) + (1.5 - 0.5j)
    for order, bw in ((4, 7.5e9), (5, 30e9), (3, 12e9), (1, 10e9), (2, 10e9)):
        sos = bessel_sos(order, bw, 640e9)
        monkeypatch.delenv("SSFM_FILTFILT_SEQUENTIAL", raising=False)
        a = ob.filtfilt_batch(x, sos)
        monkeypatch.setenv("SSFM_FILTFILT_SEQUENTIAL", "1")
        b = ob.filtfilt_batch(x, sos)
        monkeypatch.delenv("SSFM_FILTFILT_SEQUENTIAL", raising=False)
        ref = sg.sosfiltfilt(sos, x, axis=-1)
        assert rel_l2(b, ref) <= 1e-13
        assert rel_l2(a, ref) <= 1e-12
        assert np.abs(a - ref)[:, :2000].max() <= 1e-11 and np.abs(a - ref)[:, -2000:].max() <= 1e-11


def test_apply_transfer_is_the_dm_operation(ob):
    """ssfm_apply_transfer: ifft(fft(x) * H) with H in numpy bin order (reference DM, devices.py:1025-1029)."""
    import ctypes, torch
    from opticomlib_b200 import engine, _lib
    rng = np.random.default_rng(5)
    n = 1 << 14
    x = rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))
    w = 2 * np.pi * np.fft.fftfreq(n, 1 / 640e9) * 1e-12
    H = np.exp(1j * w ** 2 * (-21.27 * 30) / 2)
    plan = engine.get_plan(n, 1, 2, torch.complex128, None)
    xt = torch.from_numpy(x).cuda(); ht = torch.from_numpy(H).cuda()
    _lib.check(plan.lib.ssfm_apply_transfer(plan.handle, xt.data_ptr(), ht.data_ptr(), None))
    ref = np.fft.ifft(np.fft.fft(x, axis=-1) * H, axis=-1)
    assert rel_l2(xt.cpu().numpy(), ref) <= 1e-13


def test_filter_errors_match_reference(ob):
    _set_fs(ob, 16e9)
    with pytest.raises(TypeError, match="optical_signal"):
        ob.BPF(ob.electrical_signal(np.ones(100)), BW=1e9)
    with pytest.raises(ValueError, match="1D-array"):
        ob.LPF(ob.optical_signal(np.ones((2, 100), complex)), BW=1e9)
    with pytest.raises(ValueError, match="padlen"):                     # scipy's message for N <= edge
        ob.LPF(ob.electrical_signal(np.ones(15)), BW=1e9)


def test_dm_matches_reference_output(ob):
    """DM drop-in (N1 of SURVEY.md section 8(f)) against outputs of the unmodified reference (tests/golden/dm_*.npz)."""
    from conftest import golden
    g = golden("dm_2pol_noise_4096")
    ob.gv.dt = float(g["dt"]); ob.gv.fs = 1 / float(g["dt"])
    out, H = ob.DM(ob.optical_signal(g["x"], g["xn"]), D=float(g["D"]), retH=True)
    assert isinstance(out, ob.optical_signal) and out.signal.shape == g["out"].shape and out.signal.dtype == np.complex128
    assert rel_l2(out.signal, g["out"]) <= 1e-12 and rel_l2(out.noise, g["outn"]) <= 1e-12
    assert np.array_equal(H, g["H"])
    g = golden("dm_1pol_4096")
    out = ob.DM(ob.optical_signal(g["x"]), D=float(g["D"]))
    assert out.noise is ob.NULL and rel_l2(out.signal, g["out"]) <= 1e-12 and out.execution_time > 0
    back = ob.DM(out, D=-float(g["D"]))                                  # the inverse medium restores the field
    assert rel_l2(back.signal, g["x"]) <= 1e-12
    with pytest.raises(TypeError, match="optical_signal"):
        ob.DM(ob.electrical_signal(np.ones(1024)), D=1.0)
    # any length, as numpy.fft accepts it (chirp-z transforms): against the reference's own statements
    rng = np.random.default_rng(5)
    for n in (1000, 1270, 4099):
        x = rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))
        out = ob.DM(ob.optical_signal(x), D=2500.0)
        w = np.fft.fftfreq(n, ob.gv.dt) * 2 * np.pi
        ref = np.fft.ifft(np.fft.fft(x, axis=-1) * np.exp(1j * w ** 2 * (2500.0 * 1e-12 ** 2) / 2), axis=-1)
        assert rel_l2(out.signal, ref) <= 1e-12


@pytest.mark.parametrize("n_samples", [1 << 16, 60001, 4096, 5000])
def test_overlap_save_blocks_match_scipy_and_the_team_kernels(ob, monkeypatch, n_samples):
    """Rows of >= 4096 samples are filtered block by block (k_ols: 4096-sample blocks with a halo of K samples, any row
    length): against SciPy, against the whole-row transform path (power-of-two rows) and in place; a filter too narrow for
    the blocks (K > 1024) must still take the older paths and agree."""
    import torch
    from scipy import signal as sg
    from opticomlib_b200 import engine
    rng = np.random.default_rng(n_samples)
    rows = 5
    x = rng.standard_normal((rows, n_samples)) + 1j * rng.standard_normal((rows, n_samples)) + (0.7 - 0.2j)
    for order, bw in ((4, 7.5e9), (4, 20e9), (5, 30e9), (1, 10e9)):
        sos = bessel_sos(order, bw, 640e9)
        ref = sg.sosfiltfilt(sos, x, axis=-1)
        monkeypatch.delenv("SSFM_FILTFILT_NO_OLS", raising=False)
        a = ob.filtfilt_batch(x, sos)
        assert rel_l2(a, ref) <= 1e-12
        assert np.abs(a - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max())
        monkeypatch.setenv("SSFM_FILTFILT_NO_OLS", "1")
        b = ob.filtfilt_batch(x, sos)
        monkeypatch.delenv("SSFM_FILTFILT_NO_OLS", raising=False)
        assert rel_l2(a, b) <= 1e-12
        xd = torch.from_numpy(x).cuda()
        engine.filtfilt_sos(xd, sos, out=xd)                              # in place: blocks go through a scratch chunk
        assert rel_l2(xd.cpu().numpy(), ref) <= 1e-12
    if n_samples >= 60001:
        sos = bessel_sos(4, 2.0e9, 640e9)                                 # K ~ 3000: no block path
        assert rel_l2(ob.filtfilt_batch(x, sos), sg.sosfiltfilt(sos, x, axis=-1)) <= 1e-11


def test_overlap_save_photodetector_odd_length_with_sampler(ob):
    """PD square law + noise beat terms + LPF + sampler through the block kernel on rows whose length is no power of two
    (the square law runs in the block load, the sampler in the block store)."""
    import torch
    from scipy import signal as sg
    from opticomlib_b200 import engine
    rng = np.random.default_rng(77)
    rows, n_pol, n = 3, 2, 50003
    e = 0.03 * (rng.standard_normal((rows, n_pol, n)) + 1j * rng.standard_normal((rows, n_pol, n))) + 0.05
    z = 0.002 * (rng.standard_normal((rows, n_pol, n)) + 1j * rng.standard_normal((rows, n_pol, n)))
    extra = 1e-5 * rng.standard_normal((rows, n))
    sos = bessel_sos(4, 7.5e9, 640e9)
    r, r_load, i_dark, offset, stride = 0.9, 50.0, 1e-8, 17, 64
    sig = r_load * (r * (np.abs(e) ** 2).sum(axis=1))
    noi = r_load * (r * (2 * np.real(e * np.conj(z)) + np.abs(z) ** 2).sum(axis=1) + extra + i_dark)
    want_s = sg.sosfiltfilt(sos, sig, axis=-1)[:, offset::stride]
    want_n = sg.sosfiltfilt(sos, noi, axis=-1)[:, offset::stride]
    got_s, got_n = engine.pd_lpf(torch.from_numpy(e).cuda(), sos, torch.from_numpy(z).cuda(), torch.from_numpy(extra).cuda(),
                                 r, r_load, i_dark, offset, stride)
    assert got_s.shape == want_s.shape
    assert rel_l2(got_s.cpu().numpy(), want_s) <= TOL
    assert rel_l2(got_n.cpu().numpy(), want_n) <= TOL
