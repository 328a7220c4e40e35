"""GPU tests of the receiver / amplifier stages next to the hot path:

* `ssfm_pd_lpf` -- photodetector square law + noise assembly + zero-phase low-pass + sampler in one device call
  (reference PD devices.py:1514-1552 -> LPF 1363-1368 -> SAMPLER 1871-1891) against a NumPy restatement of those statements
  with the filter oracle (oracle/filtfilt_oracle.py, pinned to SciPy and to reference outputs); tolerance 1e-10;
* the chunked L2-resident schedule of the zero-phase filter (several chunks, ragged last chunk) against the oracle;
* `ssfm_gaussian_noise` / `ssfm_edfa` -- Philox noise on the device (reference EDFA devices.py:921-936): the deterministic
  part (gain) bit-exact against NumPy, the noise validated statistically (moments, independence, reproducibility), because
  the reference draws from NumPy's global stream.
"""
import numpy as np
import pytest

from oracle.filtfilt_oracle import oracle_sosfiltfilt, bessel_sos
from oracle.ssfm_oracle import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-10
FS = 640e9


@pytest.fixture(scope="module")
def ob():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import opticomlib_b200 as ob
    return ob


def _field(rows, n_pol, n, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / n
    env = 0.55 + 0.45 * np.sign(np.sin(2 * np.pi * 11 * t + 0.2))
    env = np.convolve(env, np.ones(15) / 15, mode="same")
    f = np.sqrt(1e-3) * env * np.exp(2j * np.pi * rng.random((rows, n_pol, 1)))
    return f + 1e-3 * (rng.standard_normal((rows, n_pol, n)) + 1j * rng.standard_normal((rows, n_pol, n))) * np.sqrt(1e-3)


def _pd_reference(field, noise, extra, r, r_load, i_dark, sos, offset, stride):
    """devices.py:1514-1552 + 1363-1368 + 1889 restated: currents, packing into signal / noise, filtfilt of each, sampling."""
    sig = (r * (field * field.conj()).real).sum(axis=1) * r_load
    out_s = oracle_sosfiltfilt(sos, sig)[..., offset::stride]
    if noise is None and extra is None:
        return out_s, None
    i_n = np.zeros(sig.shape)
    if noise is not None:
        i_n = i_n + (r * (field * noise.conj() + noise * field.conj() + noise * noise.conj()).real).sum(axis=1)
    if extra is not None:
        i_n = i_n + extra
    i_n = (i_n + i_dark) * r_load
    return out_s, oracle_sosfiltfilt(sos, i_n)[..., offset::stride]


@pytest.mark.parametrize("rows,n_pol,n,with_noise,with_extra,offset,stride", [
    (3, 1, 4096, False, False, 0, 1),
    (2, 2, 4096, True, True, 0, 1),
    (70, 1, 1 << 16, True, True, 32, 64),          # three L2 chunks (32 rows each), ragged last one, sampler
    (5, 2, 3000, True, False, 7, 16),              # not a power of two: sequential recursion path
    (4, 1, 1 << 13, False, True, 8, 16),
], ids=["plain", "2pol_noise", "chunks_sampler", "seq_path", "extra_only"])
def test_pd_lpf_matches_restatement(ob, rows, n_pol, n, with_noise, with_extra, offset, stride):
    import torch
    from opticomlib_b200 import engine
    rng = np.random.default_rng(rows + n)
    field = _field(rows, n_pol, n, 3)
    noise = 0.05 * _field(rows, n_pol, n, 4) if with_noise else None
    extra = 1e-6 * rng.standard_normal((rows, n)) if with_extra else None
    sos = bessel_sos(4, 7.5e9, FS)
    r, r_load, i_dark = 0.8, 50.0, 10e-9
    want_s, want_n = _pd_reference(field, noise, extra, r, r_load, i_dark, sos, offset, stride)
    dev = torch.device("cuda", 0)
    cu = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    got_s, got_n = engine.pd_lpf(cu(field), sos, cu(noise), cu(extra), r, r_load, i_dark, offset, stride)
    assert tuple(got_s.shape) == want_s.shape and got_s.dtype == torch.float64
    assert rel_l2(got_s.cpu().numpy(), want_s) <= TOL
    if want_n is None:
        assert got_n is None
    else:
        assert rel_l2(got_n.cpu().numpy(), want_n) <= TOL


def test_filtfilt_many_chunks_in_place_and_out_of_place(ob):
    import torch
    from opticomlib_b200 import engine
    rows, n = 150, 1 << 15                                       # 64 rows per 32 MiB chunk: 3 chunks, last one ragged
    x = _field(rows, 1, n, 8)[:, 0, :]
    sos = bessel_sos(4, 20e9, FS)
    want = oracle_sosfiltfilt(sos, x)
    xd = torch.from_numpy(x).cuda()
    y = engine.filtfilt_sos(xd, sos)
    assert rel_l2(y.cpu().numpy(), want) <= TOL
    np.testing.assert_array_equal(xd.cpu().numpy(), x)           # the input is not modified
    y2 = engine.filtfilt_sos(xd, sos, out=xd)                    # in place
    assert y2.data_ptr() == xd.data_ptr()
    assert rel_l2(xd.cpu().numpy(), want) <= TOL


def test_pd_lpf_batch_numpy_in_numpy_out(ob):
    n = 1 << 12
    field = _field(2, 1, n, 5)[:, 0, :]
    sos = bessel_sos(4, 7.5e9, FS)
    s, nz = ob.pd_lpf_batch(field, sos, responsivity=1.0, r_load=1.0)
    assert isinstance(s, np.ndarray) and nz is None
    assert rel_l2(s, oracle_sosfiltfilt(sos, np.abs(field) ** 2)) <= TOL
    with pytest.raises(ValueError):
        ob.pd_lpf_batch(field, sos, sample_offset=n)


def test_gaussian_noise_statistics_and_reproducibility(ob):
    from opticomlib_b200 import engine
    n = 1 << 22
    a = engine.gaussian_noise((n,), 2.5, seed=123, substream=1, mean=0.5).cpu().numpy()
    b = engine.gaussian_noise((n,), 2.5, seed=123, substream=1, mean=0.5).cpu().numpy()
    c = engine.gaussian_noise((n,), 2.5, seed=123, substream=2, mean=0.5).cpu().numpy()
    d = engine.gaussian_noise((n,), 2.5, seed=124, substream=1, mean=0.5).cpu().numpy()
    np.testing.assert_array_equal(a, b)                           # a pure function of (seed, substream, index)
    z = (a - 0.5) / 2.5
    se = 1 / np.sqrt(n)
    assert abs(z.mean()) < 5 * se and abs(z.var() - 1) < 5 * np.sqrt(2) * se
    assert abs((z ** 3).mean()) < 5 * np.sqrt(15) * se and abs((z ** 4).mean() - 3) < 5 * np.sqrt(96) * se
    assert abs(np.mean(z > 1.0) - 0.15865525393145707) < 5 * 0.37 * se        # tail mass of N(0,1)
    for other in (c, d):                                          # other streams / seeds are independent
        assert abs(np.corrcoef(a, other)[0, 1]) < 5 * se
    assert abs(np.corrcoef(z[::2], z[1::2])[0, 1]) < 5 * np.sqrt(2) * se       # the two outputs of one Philox call
    assert abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 5 * se
    odd = engine.gaussian_noise((7,), 1.0, seed=1).cpu().numpy()               # odd count
    assert np.isfinite(odd).all() and odd.shape == (7,)


def test_edfa_gain_exact_and_ase_statistics(ob):
    import torch
    from opticomlib_b200 import engine
    ob.gv(sps=64, R=10e9)
    n, rows = 1 << 14, 64
    base = _field(1, 1, n, 6)[0, 0]
    G, NF = 10.0, 5.0
    p_ase = 10 ** (NF / 10) * 6.62607015e-34 * ob.gv.f0 * (10 ** (G / 10) - 1) * ob.gv.fs
    # (1) deterministic part: with zero ASE power the output is exactly sqrt(idb(G)) * E (devices.py:921), broadcast to the rows
    out0 = engine.edfa(torch.from_numpy(base).cuda(), rows, G, 0.0, seed=1)
    assert tuple(out0.shape) == (rows, n)
    np.testing.assert_array_equal(out0.cpu().numpy(), np.broadcast_to(np.sqrt(10 ** (G / 10)) * base, (rows, n)))
    # (2) the public entry point: formula of devices.py:930-934, one polarisation kept (config #3) and the reference's two
    out1 = ob.edfa_batch(base, rows, G, NF, n_pol_out=1, seed=7).cpu().numpy()
    ase = out1 - np.sqrt(10 ** (G / 10)) * base
    comp = np.concatenate([ase.real.ravel(), ase.imag.ravel()])
    se = 1 / np.sqrt(comp.size)
    assert abs(comp.var() / (p_ase / 4) - 1) < 5 * np.sqrt(2) * se and abs(comp.mean()) < 5 * np.sqrt(p_ase / 4) * se
    assert abs(np.corrcoef(ase.real.ravel(), ase.imag.ravel())[0, 1]) < 5 * np.sqrt(2) * se
    assert abs(np.corrcoef(ase[0].real, ase[1].real)[0, 1]) < 5 / np.sqrt(n)       # rows are independent realisations
    out2 = ob.edfa_batch(base, 8, G, NF, seed=7).cpu().numpy()                    # default: two polarisations like the reference
    assert out2.shape == (8, 2, n)
    y = out2[:, 1, :]                                                             # y polarisation: ASE only (devices.py:924)
    assert abs(np.concatenate([y.real.ravel(), y.imag.ravel()]).var() / (p_ase / 4) - 1) < 0.05
    np.testing.assert_array_equal(ob.edfa_batch(base, 8, G, NF, seed=7).cpu().numpy(), out2)     # reproducible
    assert not np.array_equal(ob.edfa_batch(base, 8, G, NF, seed=8).cpu().numpy(), out2)
    # (3) a batch in, a batch out
    batch = _field(5, 2, n, 9)
    out3 = ob.edfa_batch(batch, 5, G, NF, seed=3).cpu().numpy()
    assert out3.shape == (5, 2, n)
    assert rel_l2(out3, np.sqrt(10 ** (G / 10)) * batch) < 5e-2                    # ASE is ~-36 dB below this signal


def test_welch_psd_matches_scipy(ob):
    """typing.py:1899-1902 / utils.py:2074-2079: sg.welch(..., scaling='spectrum', return_onesided=False, detrend=False)."""
    from scipy import signal as sg
    ob.gv(sps=64, R=10e9)
    for rows, n, nper in ((3, 1 << 14, None), (1, 5000, None), (2, 1 << 16, 1024), (4, 2048, None), (2, 700, 512)):
        x = _field(rows, 1, n, n)[:, 0, :]
        f, psd = ob.psd_batch(x, nperseg=nper)
        m = min(2048, n) if nper is None else nper
        fr, pr = sg.welch(x, fs=ob.gv.fs, nperseg=m, scaling="spectrum", return_onesided=False, detrend=False)
        np.testing.assert_allclose(f, np.fft.fftshift(fr), rtol=1e-12)
        assert psd.shape == (rows, m)
        assert rel_l2(psd, np.fft.fftshift(pr, axes=-1)) <= TOL
    with pytest.raises(ValueError):
        ob.psd_batch(_field(1, 1, 3000, 1)[:, 0, :], nperseg=1000)


def test_batch_containers_chain_stays_on_device(ob):
    """optical_batch: EDFA realisations -> FIBER -> BPF -> PD/LPF/SAMPLER against the per-row host path of the same stages."""
    import torch
    from oracle.ssfm_oracle import oracle_fiber
    ob.gv(sps=16, R=10e9)
    n, rows = 1 << 12, 6
    base = _field(1, 1, n, 2)[0, 0]
    tx = ob.optical_batch(base[None, :])
    amp = tx.edfa(G=10.0, NF=5.0, rows=rows, seed=3)
    assert amp.signal.shape == (rows, n) and amp.noise.shape == (rows, n) and amp.signal.is_cuda
    np.testing.assert_allclose(amp.signal.cpu().numpy(), np.broadcast_to(np.sqrt(10.0) * base, (rows, n)), rtol=1e-15)
    kw = dict(length=10.0, alpha=0.2, beta_2=-20.0, gamma=2.0)
    out = amp.fiber(**kw)
    fld = amp.field().cpu().numpy()
    for b in (0, rows - 1):
        with np.errstate(all="ignore"):
            ref = oracle_fiber(fld[b], ob.gv.dt, real=np.float64, **kw)
        assert int(out.ssfm_info.steps[b]) == ref["steps"] and rel_l2(out.signal[b].cpu().numpy(), ref["out"]) <= TOL
    filt = out.bpf(BW=60e9)
    sos_b = bessel_sos(4, 30e9, ob.gv.fs)
    assert rel_l2(filt.signal.cpu().numpy(), oracle_sosfiltfilt(sos_b, out.signal.cpu().numpy())) <= TOL
    rx = filt.pd(BW=7.5e9, r=1.0, R_load=50.0, sample_offset=ob.gv.sps // 2, sample_stride=ob.gv.sps)
    assert rx.signal.shape == (rows, n // ob.gv.sps) and rx.noise is None
    sos_l = bessel_sos(4, 7.5e9, ob.gv.fs)
    want = oracle_sosfiltfilt(sos_l, 50.0 * np.abs(filt.signal.cpu().numpy()) ** 2)[:, ob.gv.sps // 2::ob.gv.sps]
    assert rel_l2(rx.signal.cpu().numpy(), want) <= TOL
    sigs = rx.to_signals()
    assert len(sigs) == rows and isinstance(sigs[0], ob.electrical_signal) and sigs[0].size == n // ob.gv.sps
    f, p = out.psd()
    assert p.shape == (rows, 2048) and f.shape == (2048,)
    lp = ob.electrical_batch(np.abs(fld) ** 2, 0.1 * np.abs(fld) ** 2).lpf(BW=7.5e9)
    assert rel_l2(lp.signal.cpu().numpy(), oracle_sosfiltfilt(sos_l, np.abs(fld) ** 2)) <= TOL
    assert rel_l2(lp.noise.cpu().numpy(), oracle_sosfiltfilt(sos_l, 0.1 * np.abs(fld) ** 2)) <= TOL


@pytest.mark.parametrize("single_launch", [True, False], ids=["one_launch", "launch_per_chunk"])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_generated_monte_carlo_pipeline_equals_monolithic(ob, monkeypatch, precision, single_launch):
    """edfa_fiber_batch (fp64: the batch generated on the device, ONE streamed launch, chunks copied out as the kernel finishes
    them; otherwise chunks generated, propagated and copied out on three streams) against edfa_batch + fiber_batch on the
    whole batch at once: the noise of a row depends on (seed, global row index) only, so the rows are bit-identical."""
    import torch
    from opticomlib_b200 import devices
    ob.gv(sps=16, R=10e9)
    n, rows = 1 << 13, 29
    base = _field(1, 1, n, 12)[0, 0] * 3.0
    kw = dict(length=8.0, alpha=0.2, beta_2=-20.0, gamma=2.0, phi_max=0.01)
    monkeypatch.setattr(devices, "HOST_CHUNK_BYTES", 4 * n * 16)              # 8 chunks, ragged last one
    monkeypatch.setattr(devices, "HOST_SINGLE_CHUNK_BYTES", 4 * n * 16)
    monkeypatch.setattr(devices, "HOST_SINGLE_LAUNCH", single_launch)
    got, info = ob.edfa_fiber_batch(base, rows, 10.0, 5.0, ob.gv.dt, seed=77, precision=precision, **kw)
    assert got.is_pinned() and tuple(got.shape) == (rows, n)
    whole = ob.edfa_batch(base, rows, 10.0, 5.0, n_pol_out=1, seed=77)
    want, iw = ob.fiber_batch(whole, ob.gv.dt, precision=precision, **kw)
    np.testing.assert_array_equal(info.steps, iw.steps)
    np.testing.assert_array_equal(got.numpy(), want.cpu().numpy())
    part, _ = ob.edfa_fiber_batch(base, 5, 10.0, 5.0, ob.gv.dt, seed=77, precision=precision, first_row=11, **kw)
    np.testing.assert_array_equal(part.numpy(), got.numpy()[11:16])            # rows 11..15 of the same global batch
