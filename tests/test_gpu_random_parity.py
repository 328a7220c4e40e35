"""Randomised parity: FIBER / DBP on the GPU against the CPU oracle over drawn lengths (powers of two through k_wf and the
multi-launch schedule, arbitrary lengths through chirp-z), polarisations, fibre parameters and step rules.
Hypothesis runs derandomised (fixed example sequence), so a failure reproduces."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

from oracle.ssfm_oracle import oracle_fiber, oracle_dbp, rel_l2

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "fp64": 1e-10}
REAL = {"fp32": np.float32, "fp64": np.float64}


@pytest.fixture(scope="module")
def ob():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import opticomlib_b200 as ob
    return ob


lengths = st.one_of(st.integers(8, 15).map(lambda e: 1 << e), st.integers(2, 6000))
case = st.fixed_dictionaries(dict(
    n=lengths, n_pol=st.sampled_from([1, 1, 2]), precision=st.sampled_from(["fp64", "fp32"]), dbp=st.booleans(),
    fixed=st.booleans(), seed=st.integers(0, 2 ** 16),
    length=st.floats(0.5, 40.0), alpha=st.sampled_from([0.0, 0.2, 0.35]),
    beta_2=st.sampled_from([0.0, -21.27, -20.0, 5.0]), beta_3=st.sampled_from([0.0, 0.127, -0.1]),
    gamma=st.sampled_from([0.0, 1.3, 2.0, 5.0]), phi_max=st.sampled_from([0.005, 0.01, 0.05, 0.2]),
    steps_fixed=st.integers(1, 12), power_mw=st.floats(0.1, 20.0)))


@settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(c=case)
def test_random_cases_match_the_oracle(ob, c):
    n, P = c["n"], c["n_pol"]
    rng = np.random.default_rng(c["seed"])
    t = np.arange(n) / n
    env = np.sqrt(c["power_mw"] * 1e-3) * (0.6 + 0.4 * np.cos(2 * np.pi * 3 * t + 0.4))
    x = np.stack([env * np.exp(2j * np.pi * (2 + p) * t) + 1e-3 * np.sqrt(c["power_mw"] * 1e-3) *
                  (rng.standard_normal(n) + 1j * rng.standard_normal(n)) for p in range(P)])
    x = x[0] if P == 1 else x
    kw = dict(length=c["length"], alpha=c["alpha"], beta_2=c["beta_2"], beta_3=c["beta_3"], gamma=c["gamma"])
    if c["fixed"]:
        kw["h"] = c["length"] / c["steps_fixed"] * 1.0001
    else:
        kw["phi_max"] = c["phi_max"]
    real = REAL[c["precision"]]
    orc, fn = (oracle_dbp, ob.dbp_batch) if c["dbp"] else (oracle_fiber, ob.fiber_batch)
    with np.errstate(all="ignore"):
        ref = orc(x, 1 / 160e9, real=real, max_steps=3000, **kw)
    if ref["steps"] >= 3000 or not np.isfinite(ref["out"]).all():
        return                                                              # runaway draw (huge phase budget): nothing to compare
    out, info = fn(x[None], 1 / 160e9, precision=c["precision"], **kw)
    assert int(info.steps[0]) == ref["steps"], c
    # fp32 noise floor grows with the number of steps: 1e-4 holds to ~2000 steps
    assert rel_l2(out[0], ref["out"]) <= TOL[c["precision"]], (c, rel_l2(out[0], ref["out"]))
