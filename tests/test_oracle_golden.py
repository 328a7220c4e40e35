"""Pin the CPU oracle to outputs of the unmodified reference (tests/golden/*.npz).

CPU-only.  The fixtures were produced by tests/golden/make_golden.py, i.e. by the
reference's own FIBER / DBP / LPF / BPF / float64 loop.  The float32 oracle must
reproduce the shipped (complex64) result essentially bit for bit; the tolerance
below (1e-6 rel-L2) only allows for a different SIMD path of NumPy's exp/abs on
another host CPU -- on the machine that generated the fixtures it is exactly 0.
"""
import numpy as np
import pytest

from conftest import golden, fiber_kwargs
from oracle.ssfm_oracle import oracle_fiber, oracle_dbp, rel_l2
from oracle.filtfilt_oracle import oracle_lpf, oracle_bpf
from opticomlib_b200 import workloads as wl

F32_CASES = [
    "fiber_adaptive_4096", "fiber_beta3_4096", "dbp_adaptive_4096", "fiber_gamma0_4096",
    "fiber_nodisp_4096", "fiber_alpha_only_4096", "fiber_fixed_h03_2048", "fiber_fixed_h01_1024",
    "fiber_2pol_noise_4096", "dbp_2pol_fixed_4096",
    # lengths that are not powers of two (the reference accepts any N): 1270, 3000, 1000 (two polarisations), 999 (odd)
    "fiber_n1270_adaptive", "fiber_n3000_fixed", "dbp_2pol_n1000", "fiber_n999_odd",
]


@pytest.mark.parametrize("name", F32_CASES)
def test_f32_oracle_matches_reference_output(name):
    g = golden(name)
    kw = fiber_kwargs(g)
    fn = oracle_dbp if bool(g["is_dbp"]) else oracle_fiber
    with np.errstate(all="ignore"):
        o = fn(g["x"], float(g["dt"]), real=np.float32, **kw)
    assert o["out"].dtype == np.complex64 and o["out"].shape == g["out"].shape
    assert rel_l2(o["out"], g["out"]) <= 1e-6
    # identical step count and (float32) positions: z[0] = 0.0 then one entry per step
    assert o["steps"] == len(g["z"]) - 1
    if kw.get("h") is not None:
        assert np.array_equal(o["z"].astype(np.float64), g["z"][1:])  # fixed-h bookkeeping is bit exact
    else:
        np.testing.assert_allclose(o["z"].astype(np.float64), g["z"][1:], rtol=1e-5)


def test_fixed_h_sliver_step_counts():
    # float32 accumulation of z decides the count (501, not 500; 167 for h=0.3)
    assert len(golden("fiber_fixed_h01_1024")["z"]) - 1 == 501
    assert len(golden("fiber_fixed_h03_2048")["z"]) - 1 == 167


def test_cfg1_full_size_against_reference():
    g = golden("fiber_cfg1_65536")
    x, dt, kw = wl.config_input("cfg1")
    assert x.shape == (65536,) and dt == float(g["dt"])
    assert np.array_equal(x[:64], g["x_head"])  # builder reproduces the reference TX chain
    o = oracle_fiber(x, dt, real=np.float32, **kw)
    assert o["steps"] == 8 == len(g["z"]) - 1
    assert rel_l2(o["out"][::16], g["out_dec"]) <= 1e-6
    np.testing.assert_allclose(o["z"].astype(np.float64), g["z"][1:], rtol=1e-5)


def test_f64_oracle_matches_reference_float64_loop():
    # devices.py:2440-2486 (animated_fiber_propagation_with_phase) is the reference's float64 SSFM
    g = golden("fiber_f64_4096")
    kw = fiber_kwargs(g)
    o = oracle_fiber(g["x"], float(g["dt"]), real=np.float64, **kw)
    assert o["out"].dtype == np.complex128
    assert rel_l2(o["out"], g["out"]) <= 1e-13
    assert o["steps"] == len(g["z"]) - 1
    np.testing.assert_allclose(o["z"], g["z"][1:], rtol=1e-12)


def test_f32_and_f64_oracles_agree_to_single_precision():
    g = golden("fiber_adaptive_4096")
    kw = fiber_kwargs(g)
    a = oracle_fiber(g["x"], float(g["dt"]), real=np.float32, **kw)
    b = oracle_fiber(g["x"], float(g["dt"]), real=np.float64, **kw)
    assert a["steps"] == b["steps"]
    assert rel_l2(a["out"], b["out"]) < 1e-4


@pytest.mark.parametrize("name", ["lpf_ones_100", "lpf_n3_4096", "lpf_n4_4096", "lpf_n5_4096", "lpf_complex_1000"])
def test_lpf_oracle_matches_reference(name):
    g = golden(name)
    xn = g["xn"] if "xn" in g.files else None
    s, nz = oracle_lpf(g["x"], xn, float(g["bw"]), float(g["fs"]), int(g["n"]))
    assert s.dtype == np.float64 and s.shape == g["out"].shape
    assert rel_l2(s, g["out"]) <= 1e-12
    if xn is not None:
        assert rel_l2(nz, g["outn"]) <= 1e-12


@pytest.mark.parametrize("name", ["bpf_2pol_4096", "bpf_1pol_n5_8192"])
def test_bpf_oracle_matches_reference(name):
    g = golden(name)
    xn = g["xn"] if "xn" in g.files else None
    s, nz = oracle_bpf(g["x"], xn, float(g["bw"]), float(g["fs"]), int(g["n"]))
    assert s.dtype == np.complex128 and s.shape == g["out"].shape
    assert rel_l2(s, g["out"]) <= 1e-12
    if xn is not None:
        assert rel_l2(nz, g["outn"]) <= 1e-12
