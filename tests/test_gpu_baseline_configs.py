"""Parity at the sizes BASELINE.json names: config #2 in full (one 2^20-sample waveform, 100 km,
beta_3, phi_max control, 20 dBm: 186 adaptive steps), config #3 on a fixed subset of its 4096 Monte-Carlo rows with their own
step counts, and the first split steps of config #5's 2^26-sample waveform -- each against the CPU oracle on the same input."""
import numpy as np
import pytest

from oracle.ssfm_oracle import oracle_fiber, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import opticomlib_b200 as ob
    return ob


def test_config2_full(ob):
    from opticomlib_b200 import workloads as wl
    x, dt, kw = wl.config_input("cfg2")
    assert x.size == 1 << 20
    ref = oracle_fiber(x, dt, real=np.float64, **kw)
    out, info = ob.fiber_batch(x[None, :], dt, precision="fp64", want_log=True, **kw)
    assert int(info.steps[0]) == ref["steps"] and ref["steps"] > 150
    np.testing.assert_allclose(info.h_log[0, :ref["steps"]], ref["h"], rtol=1e-10)
    assert rel_l2(out[0], ref["out"]) <= 1e-10
    # the shipped precision on the same input: same step count as its own oracle, and the fp32-vs-fp64 gap of SURVEY 8(c)
    out32, info32 = ob.fiber_batch(x[None, :], dt, precision="fp32", **kw)
    assert abs(int(info32.steps[0]) - ref["steps"]) <= 1
    assert rel_l2(out32[0], ref["out"]) <= 1e-4


def test_config3_subset_of_rows(ob):
    from opticomlib_b200 import workloads as wl
    base, dt, kw = wl.config_input("cfg1")
    c = wl.CONFIGS["cfg3"]
    fs = c["R"] * c["sps"]
    idx = list(range(36)) + [1000, 4095]                                    # more rows than teams in flight
    rows = wl.ase_rows(base, idx, fs, c["gain_db"], c["nf_db"])
    for precision, real, tol in (("fp64", np.float64, 1e-10), ("fp32", np.float32, 1e-4)):
        out, info = ob.fiber_batch(rows, dt, precision=precision, **c["fiber"])
        for b in (0, 1, 17, 4095):                                           # the rows SURVEY.md 8(d) names
            i = idx.index(b)
            ref = oracle_fiber(rows[i], dt, real=real, **c["fiber"])
            assert int(info.steps[i]) == ref["steps"], (precision, b)
            assert rel_l2(out[i], ref["out"]) <= tol, (precision, b)
        assert info.done.all() and 60 < info.steps.mean() < 90


def test_config5_first_steps(ob):
    """2^26 samples, two fixed steps of 1 km (the CPU oracle needs ~15 s per step at this size)."""
    import torch
    from opticomlib_b200 import longwave as lw, workloads as wl
    c = wl.CONFIGS["cfg5"]
    n = 1 << 26
    bits = wl.prbs(c["order"], n // c["sps"])
    x = wl.mzm_field(wl.nrz_drive(bits, c["sps"]), c["p0_dbm"]).astype(np.complex128)
    kw = dict(c["fiber"]); kw["length"] = 2.0
    dt = 1.0 / (c["R"] * c["sps"])
    ref = oracle_fiber(x, dt, real=np.float64, **kw)
    out, info = lw.fiber_long(torch.from_numpy(x), dt, precision="fp64", **kw)
    assert int(info.steps[0]) == ref["steps"] == 2
    assert rel_l2(out.cpu().numpy(), ref["out"]) <= 1e-10
    lw.clear_plans()
