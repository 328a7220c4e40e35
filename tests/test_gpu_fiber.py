"""GPU parity tests for FIBER / DBP (run on the B200 box: pytest -m gpu).

Every comparison goes  CUDA kernels (through the C-ABI)  vs  CPU oracle / golden fixtures.
Tolerances are BASELINE.json's: rel-L2 <= 1e-4 in fp32 against the reference's shipped output,
<= 1e-10 in fp64 against the dtype-lifted oracle; identical step counts; fixed-h positions bit-exact.
"""
import ctypes
import math

import numpy as np
import pytest

from conftest import golden, fiber_kwargs
from oracle.ssfm_oracle import oracle_fiber, oracle_dbp, rel_l2

pytestmark = pytest.mark.gpu

TOL32, TOL64 = 1e-4, 1e-10

F32_CASES = [
    "fiber_adaptive_4096", "fiber_beta3_4096", "dbp_adaptive_4096", "fiber_gamma0_4096",
    "fiber_nodisp_4096", "fiber_alpha_only_4096", "fiber_fixed_h03_2048", "fiber_fixed_h01_1024",
    "fiber_2pol_noise_4096", "dbp_2pol_fixed_4096",
    # lengths that are not powers of two (the reference accepts any N): 1270, 3000, 1000 (two polarisations), 999 (odd)
    "fiber_n1270_adaptive", "fiber_n3000_fixed", "dbp_2pol_n1000", "fiber_n999_odd",
]


@pytest.fixture(scope="module")
def ob():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import opticomlib_b200 as ob
    return ob


def _set_dt(ob, dt):
    ob.gv.dt = float(dt)
    ob.gv.fs = 1.0 / float(dt)


def _run(ob, g, precision, **extra):
    kw = fiber_kwargs(g)
    _set_dt(ob, g["dt"])
    fn = ob.DBP if bool(g["is_dbp"]) else ob.FIBER
    return fn(ob.optical_signal(g["x"]), precision=precision, **kw, **extra), kw


@pytest.mark.parametrize("name", F32_CASES)
def test_fp32_matches_reference_output(ob, name):
    g = golden(name)
    out, kw = _run(ob, g, "fp32")
    assert isinstance(out, ob.optical_signal) and out.noise is ob.NULL
    assert out.signal.dtype == np.complex64 and out.signal.shape == g["out"].shape
    assert rel_l2(out.signal, g["out"]) <= TOL32
    info = out.ssfm_info
    assert int(info.steps[0]) == len(g["z"]) - 1                     # identical step count
    assert out.execution_time > 0


@pytest.mark.parametrize("name", ["fiber_fixed_h03_2048", "fiber_fixed_h01_1024", "dbp_2pol_fixed_4096",
                                  "fiber_adaptive_4096", "fiber_beta3_4096", "fiber_n3000_fixed", "fiber_n999_odd"])
def test_fp32_step_positions(ob, name):
    g = golden(name)
    kw = fiber_kwargs(g)
    _set_dt(ob, g["dt"])
    fn = ob.DBP if bool(g["is_dbp"]) else ob.FIBER
    z, traj = fn(ob.optical_signal(g["x"]), return_steps=True, **kw)
    assert z.dtype == np.float64 and len(z) == len(g["z"]) and z[0] == 0.0
    assert traj.shape == (len(z),) + g["out"].shape and traj.dtype == np.complex64
    if kw.get("h") is not None:
        assert np.array_equal(z, g["z"])                               # float32 bookkeeping, bit exact
    else:
        np.testing.assert_allclose(z, g["z"], rtol=1e-3)              # SURVEY.md §7 hard part 3
    assert rel_l2(traj[-1], g["out"]) <= TOL32
    assert rel_l2(traj[0], g["x"].astype(np.complex64)) == 0.0


@pytest.mark.parametrize("name", F32_CASES + ["fiber_f64_4096"])
def test_fp64_matches_lifted_oracle(ob, name):
    g = golden(name)
    kw = fiber_kwargs(g)
    is_dbp = bool(g["is_dbp"]) if "is_dbp" in g.files else False
    _set_dt(ob, g["dt"])
    fn, orc = (ob.DBP, oracle_dbp) if is_dbp else (ob.FIBER, oracle_fiber)
    with np.errstate(all="ignore"):
        ref = orc(g["x"], float(g["dt"]), real=np.float64, **kw)
    out = fn(ob.optical_signal(g["x"]), precision="fp64", **kw)
    assert out.signal.dtype == np.complex128
    assert rel_l2(out.signal, ref["out"]) <= TOL64
    info = out.ssfm_info
    assert int(info.steps[0]) == ref["steps"]
    np.testing.assert_allclose(info.z[0], ref["z"][-1], rtol=1e-12)
    if name == "fiber_f64_4096":                                       # the reference's own float64 loop
        assert rel_l2(out.signal, g["out"]) <= TOL64


def test_fp64_step_sizes_match_oracle(ob):
    g = golden("fiber_beta3_4096")
    kw = fiber_kwargs(g)
    ref = oracle_fiber(g["x"], float(g["dt"]), real=np.float64, **kw)
    out, info = ob.fiber_batch(g["x"][None, :], float(g["dt"]), precision="fp64", want_log=True, **kw)
    assert int(info.steps[0]) == ref["steps"]
    np.testing.assert_allclose(info.h_log[0, :ref["steps"]], ref["h"], rtol=1e-10)
    assert rel_l2(out[0], ref["out"]) <= TOL64


def test_cfg1_full_size(ob):
    from opticomlib_b200 import workloads as wl
    x, dt, kw = wl.config_input("cfg1")
    g = golden("fiber_cfg1_65536")
    ob.gv(sps=64, R=10e9)
    assert ob.gv.dt == dt
    out = ob.FIBER(ob.optical_signal(x), **kw)
    assert int(out.ssfm_info.steps[0]) == 8
    assert rel_l2(out.signal[::16], g["out_dec"]) <= TOL32           # vs the unmodified reference
    ref32 = oracle_fiber(x, dt, real=np.float32, **kw)
    assert rel_l2(out.signal, ref32["out"]) <= TOL32
    ref64 = oracle_fiber(x, dt, real=np.float64, **kw)
    out64 = ob.FIBER(ob.optical_signal(x), precision="fp64", **kw)
    assert rel_l2(out64.signal, ref64["out"]) <= TOL64
    assert int(out64.ssfm_info.steps[0]) == ref64["steps"]


@pytest.mark.parametrize("log2n", list(range(8, 23)))
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_every_supported_length(ob, log2n, precision):
    """All transform-size instantiations (N = 2^8 .. 2^22), two fixed steps, one row."""
    n = 1 << log2n
    rng = np.random.default_rng(log2n)
    t = np.arange(n) / n
    x = (np.sqrt(1e-3) * (1 + 0.5 * np.cos(2 * np.pi * 5 * t)) * np.exp(2j * np.pi * 3 * t)
         + 1e-3 * (rng.standard_normal(n) + 1j * rng.standard_normal(n)))
    dt = 1 / 160e9
    kw = dict(length=2.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, h=1.0)
    real = np.float64 if precision == "fp64" else np.float32
    ref = oracle_fiber(x, dt, real=real, **kw)
    out, info = ob.fiber_batch(x[None, :], dt, precision=precision, **kw)
    assert int(info.steps[0]) == ref["steps"] == 2
    assert rel_l2(out[0], ref["out"]) <= (TOL64 if precision == "fp64" else TOL32)


def test_batch_rows_have_their_own_step_sequences(ob):
    g = golden("fiber_adaptive_4096")
    kw = fiber_kwargs(g)
    dt = float(g["dt"])
    scales = [1.0, 0.5, 1.7, 0.05, 1.0, 2.5, 0.9]
    rows = np.stack([np.sqrt(s) * np.roll(g["x"], 37 * i) for i, s in enumerate(scales)])
    for precision, real, tol in (("fp64", np.float64, TOL64), ("fp32", np.float32, TOL32)):
        out, info = ob.fiber_batch(rows, dt, precision=precision, **kw)
        for chunk in (2, 3):                                            # chunked scheduling gives the same result
            out_c, info_c = ob.fiber_batch(rows, dt, precision=precision, chunk_waveforms=chunk, **kw)
            assert np.array_equal(out_c, out) and np.array_equal(info_c.steps, info.steps)
        counts = []
        for i in range(len(scales)):
            ref = oracle_fiber(rows[i], dt, real=real, **kw)
            assert int(info.steps[i]) == ref["steps"], (precision, i)
            assert rel_l2(out[i], ref["out"]) <= tol
            counts.append(ref["steps"])
        assert len(set(counts)) > 2                                     # rows really diverged
        assert info.done.all()


@pytest.mark.parametrize("precision,tol", [("fp64", 1e-13), ("fp32", 2e-6)])
def test_fused_and_unfused_schedules_agree(ob, precision, tol):
    """k_col_mid (2R+2W per step, one merged Kerr rotation) vs col_inv + col_fwd (3R+3W)."""
    g = golden("fiber_beta3_4096")
    kw = fiber_kwargs(g)
    rows = np.stack([g["x"], 0.7 * np.roll(g["x"], 100), 1.3 * g["x"][::-1]])
    for extra in ({}, {"h": 0.37}):
        k = {**kw, **extra}
        a, ia = ob.fiber_batch(rows, float(g["dt"]), precision=precision, fused=True, want_log=True, **k)
        b, ib = ob.fiber_batch(rows, float(g["dt"]), precision=precision, fused=False, want_log=True, **k)
        assert np.array_equal(ia.steps, ib.steps)
        assert rel_l2(a, b) <= tol
        if "h" in extra:
            assert np.array_equal(ia.h_log, ib.h_log) and np.array_equal(ia.z, ib.z)


@pytest.mark.parametrize("n", [2, 3, 100, 255, 257, 6561, 10000, 65537])
def test_arbitrary_lengths_against_the_oracle(ob, n):
    """Any N, as numpy.fft accepts it: chirp-z transforms on the power-of-two kernels (fp64 and fp32), batch of 3 rows with
    their own step sequences."""
    rng = np.random.default_rng(n)
    t = np.arange(n) / n
    base = np.sqrt(2e-3) * (1 + 0.5 * np.cos(2 * np.pi * 3 * t)) * np.exp(2j * np.pi * 2 * t) + 1e-3 * (
        rng.standard_normal(n) + 1j * rng.standard_normal(n))
    rows = np.stack([base, 0.6 * base[::-1], 1.5 * np.roll(base, n // 3)])
    dt = 1 / 160e9
    for kw in (dict(length=12.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, phi_max=0.02),
               dict(length=2.2, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, h=0.5)):
        for precision, real, tol in (("fp64", np.float64, TOL64), ("fp32", np.float32, TOL32)):
            out, info = ob.fiber_batch(rows, dt, precision=precision, want_log=True, **kw)
            for i in range(3):
                ref = oracle_fiber(rows[i], dt, real=real, **kw)
                assert int(info.steps[i]) == ref["steps"], (precision, i)
                assert rel_l2(out[i], ref["out"]) <= tol, (precision, i)
                np.testing.assert_allclose(info.h_log[i, :ref["steps"]], ref["h"], rtol=1e-3 if precision == "fp32" else 1e-10)


def test_two_polarisations_share_one_controller(ob):
    g = golden("fiber_2pol_noise_4096")
    kw = fiber_kwargs(g)
    dt = float(g["dt"])
    batch = np.stack([g["x"], 0.3 * g["x"][::-1]])                     # [B=2, P=2, N]
    out, info = ob.fiber_batch(batch, dt, precision="fp64", **kw)
    for i in range(2):
        ref = oracle_fiber(batch[i], dt, real=np.float64, **kw)
        assert int(info.steps[i]) == ref["steps"]
        assert rel_l2(out[i], ref["out"]) <= TOL64


def test_properties_at_full_size(ob):
    """Size-independent checks at BASELINE sizes (N = 2^18 and 2^20), fp64."""
    import torch
    rng = np.random.default_rng(7)
    dt = 1 / 640e9
    for n, rows in ((1 << 18, 6), (1 << 20, 2)):
        x = np.sqrt(5e-3) * (rng.standard_normal((rows, n)) + 1j * rng.standard_normal((rows, n))) / np.sqrt(2)
        xt = torch.from_numpy(x).cuda()
        # (1) lossless fibre conserves energy: Kerr steps are pure phase, the linear step is unitary
        y, info = ob.fiber_batch(xt, dt, length=3.0, alpha=0.0, beta_2=-21.27, beta_3=0.127, gamma=1.3, h=0.5,
                                 precision="fp64")
        e0 = (xt.abs() ** 2).sum(dim=-1); e1 = (y.abs() ** 2).sum(dim=-1)
        assert torch.allclose(e1, e0, rtol=1e-11, atol=0)
        assert (info.steps == 6).all()
        # (2) linear propagation followed by DBP is the identity
        y, _ = ob.fiber_batch(xt, dt, length=40.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=0.0, precision="fp64")
        zb, _ = ob.dbp_batch(y, dt, length=40.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=0.0, precision="fp64")
        assert float((zb - xt).norm() / xt.norm()) <= 1e-11
        # (3) attenuation only: exact power scaling exp(-alpha_lin L)   (reference test_FIBER)
        y, _ = ob.fiber_batch(xt, dt, length=10.0, alpha=0.2, precision="fp64")
        ratio = float(((y.abs() ** 2).sum() / (xt.abs() ** 2).sum()))
        assert abs(ratio / math.exp(-(0.2 / 4.343) * 10.0) - 1) < 1e-12
        # (4) first row against the oracle (one adaptive run, a few steps)
        ref = oracle_fiber(x[0], dt, length=0.5, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, real=np.float64,
                           phi_max=0.05)
        y, info = ob.fiber_batch(xt[:1], dt, length=0.5, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3,
                                 phi_max=0.05, precision="fp64")
        assert int(info.steps[0]) == ref["steps"]
        assert rel_l2(y[0].cpu().numpy(), ref["out"]) <= TOL64


def test_reference_unit_tests_pass_on_this_path(ob):
    """tests/devices_test.py:257-277 of the reference, run against the CUDA path."""
    ob.gv(sps=16, R=1e9, N=128)
    cw = ob.optical_signal(np.ones(ob.gv.N * ob.gv.sps) * np.sqrt(10e-3))
    fiber = ob.FIBER(cw, length=10, alpha=0.2)
    assert fiber.type is ob.optical_signal
    expected = np.mean(np.abs(cw.signal) ** 2) * np.exp(-(0.2 / 4.343) * 10)
    np.testing.assert_allclose(np.mean(np.abs(fiber.signal) ** 2), expected, rtol=1e-3)
    f0 = ob.FIBER(cw, length=10, alpha=0, beta_2=0, gamma=0)
    back = ob.DBP(f0, length=10, alpha=0, beta_2=0, gamma=0)
    np.testing.assert_allclose(back.signal, cw.signal, atol=1e-5)


def test_errors_match_reference(ob):
    with pytest.raises(TypeError, match="optical_signal"):
        ob.FIBER(np.ones(1024, complex), length=1.0)
    with pytest.raises(TypeError):
        ob.FIBER(ob.electrical_signal(np.ones(1024)), length=1.0)
    with pytest.raises(ValueError):                                    # unsupported length is loud, not a fallback
        ob.FIBER(ob.optical_signal(np.ones(1, complex)), length=1.0)
    with pytest.raises(ValueError):
        ob.fiber_batch(np.ones((1, 3 << 20), complex), 1e-12, length=1.0)


def test_c_abi_host_entry_point(ob):
    """ssfm_fiber_host straight through ctypes: plain pointers in, plain pointers out."""
    from opticomlib_b200 import _lib
    lib = _lib.load()
    g = golden("fiber_adaptive_4096")
    kw = fiber_kwargs(g)
    x = np.ascontiguousarray(g["x"], dtype=np.complex128)
    y = np.empty_like(x)
    plan = ctypes.c_void_p()
    assert lib.ssfm_plan_create(ctypes.byref(plan), x.size, 1, 1, _lib.SSFM_C128, 0) == 0
    prm = _lib.FiberParams(float(g["dt"]), kw["length"], kw["alpha"], kw["beta_2"], 0.0, kw["gamma"], 0.01, math.nan)
    assert lib.ssfm_fiber_host(plan, x.ctypes.data, y.ctypes.data, ctypes.byref(prm), None) == 0
    steps = np.zeros(1, np.int32)
    assert lib.ssfm_get_state(plan, steps.ctypes.data, None, None, None) == 0
    ref = oracle_fiber(x, float(g["dt"]), real=np.float64, **kw)
    assert steps[0] == ref["steps"] and rel_l2(y, ref["out"]) <= TOL64
    # error path: no exception across the ABI, a code and a message instead
    bad = ctypes.c_void_p()
    assert lib.ssfm_plan_create(ctypes.byref(bad), 3 << 20, 1, 1, _lib.SSFM_C128, 0) == _lib.SSFM_ERR_UNSUPPORTED
    assert b"power of two" in lib.ssfm_last_error()
    assert lib.ssfm_plan_destroy(plan) == 0
