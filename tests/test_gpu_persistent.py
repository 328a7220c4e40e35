"""GPU parity tests of the persistent whole-propagation kernel k_wf (opticomlib_b200/csrc/ssfm_wf.cuh).

The golden fixtures of the reference are 1024..4096 samples long and therefore run on the multi-launch
schedule for 1024 and 2048 samples; k_wf adopts waveforms of 2^12 .. 2^20 samples.  Every case here is checked against the CPU oracle
(oracle/ssfm_oracle.py, pinned to the reference) and against the multi-launch schedule on the same input.
Tolerances are BASELINE.json's: rel-L2 <= 1e-4 (fp32), <= 1e-10 (fp64), identical step counts.
"""
import numpy as np
import pytest

from oracle.ssfm_oracle import oracle_fiber, oracle_dbp, rel_l2

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "fp64": 1e-10}
REAL = {"fp32": np.float32, "fp64": np.float64}
DT = 1 / 640e9


@pytest.fixture(scope="module")
def ob():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import opticomlib_b200 as ob
    return ob


def _wave(n, seed, power=2e-3, n_pol=1):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / n
    rows = []
    for p in range(n_pol):
        env = np.sqrt(power) * (0.55 + 0.45 * np.sign(np.sin(2 * np.pi * (37 + 5 * p) * t + 0.3)))   # OOK-like envelope
        env = np.convolve(env, np.ones(9) / 9, mode="same")
        rows.append(env * np.exp(2j * np.pi * 3 * t) + 2e-3 * np.sqrt(power) * (rng.standard_normal(n) + 1j * rng.standard_normal(n)))
    return rows[0] if n_pol == 1 else np.stack(rows)


def _kind(ob, n, n_pol, rows, precision):
    import torch
    from opticomlib_b200 import engine
    td = torch.complex64 if precision == "fp32" else torch.complex128
    return engine.get_plan(n, n_pol, rows, td, torch.device("cuda", 0)).last_timing()


CASES = [
    # name, log2n, kwargs
    ("adaptive_2_12", 12, dict(length=12.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.01)),
    ("fixed_2_12", 12, dict(length=3.05, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, h=0.3)),
    ("adaptive_2_13", 13, dict(length=10.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, phi_max=0.02)),
    ("adaptive_b3", 14, dict(length=12.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.01)),
    ("adaptive_odd", 15, dict(length=10.0, alpha=0.2, beta_2=-20.0, beta_3=0.0, gamma=2.0, phi_max=0.02)),
    ("adaptive_16", 16, dict(length=25.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, phi_max=0.01)),
    ("fixed_h", 16, dict(length=3.05, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, h=0.3)),
    ("large_phase", 14, dict(length=20.0, alpha=0.1, beta_2=-20.0, gamma=30.0, phi_max=0.3)),
    ("gamma0", 14, dict(length=40.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=0.0)),
    ("nodisp", 15, dict(length=30.0, alpha=0.2, gamma=2.0)),
    ("alpha_only", 14, dict(length=10.0, alpha=0.2)),
    ("adaptive_17", 17, dict(length=6.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.02)),
]


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("name,log2n,kw", CASES, ids=[c[0] for c in CASES])
def test_persistent_matches_oracle_and_multilaunch(ob, name, log2n, kw, precision):
    n = 1 << log2n
    x = _wave(n, log2n)
    with np.errstate(all="ignore"):
        ref = oracle_fiber(x, DT, real=REAL[precision], **kw)
    out, info = ob.fiber_batch(x[None, :], DT, precision=precision, persistent=True, want_log=True, **kw)
    assert _kind(ob, n, 1, 1, precision)[0] == 2, "the persistent kernel did not run"
    assert int(info.steps[0]) == ref["steps"]
    assert rel_l2(out[0], ref["out"]) <= TOL[precision]
    np.testing.assert_allclose(info.z[0], ref["z"][-1], rtol=1e-6 if precision == "fp32" else 1e-12)
    np.testing.assert_allclose(info.h_log[0, :ref["steps"]], ref["h"], rtol=1e-3 if precision == "fp32" else 1e-10)
    out_m, info_m = ob.fiber_batch(x[None, :], DT, precision=precision, persistent=False, **kw)
    assert _kind(ob, n, 1, 1, precision)[0] == 1
    assert int(info_m.steps[0]) == int(info.steps[0])
    assert rel_l2(out[0], out_m[0]) <= (2e-6 if precision == "fp32" else 1e-13)
    if kw.get("h") is not None:
        assert info.z[0] == info_m.z[0]                                    # fixed-h bookkeeping is bit exact


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_persistent_dbp_two_polarisations(ob, precision):
    n = 1 << 14
    x = _wave(n, 5, n_pol=2)
    kw = dict(length=15.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.01)
    with np.errstate(all="ignore"):
        ref = oracle_dbp(x, DT, real=REAL[precision], **kw)
    out, info = ob.dbp_batch(x[None], DT, precision=precision, persistent=True, **kw)
    assert _kind(ob, n, 2, 1, precision)[0] == 2
    assert int(info.steps[0]) == ref["steps"]
    assert rel_l2(out[0], ref["out"]) <= TOL[precision]


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_more_waveforms_than_teams_with_diverging_step_counts(ob, precision):
    """2^16 samples: 18 (fp64) teams of 16 CTAs are in flight; 45 waveforms of different power are handed out
    dynamically and finish after different numbers of steps."""
    n, rows = 1 << 16, 45
    base = _wave(n, 11)
    scales = 0.2 + 2.3 * np.random.default_rng(3).random(rows)
    x = np.stack([np.sqrt(s) * np.roll(base, 997 * i) for i, s in enumerate(scales)])
    kw = dict(length=8.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, phi_max=0.01)
    import torch
    from opticomlib_b200 import engine
    td = torch.complex64 if precision == "fp32" else torch.complex128
    xd = torch.from_numpy(x).cuda().to(td)                      # device-resident batch: ONE plan / one launch for all 45 rows
    out_t, info = ob.fiber_batch(xd, DT, precision=precision, persistent=True, **kw)
    out = out_t.cpu().numpy()
    kind, teams, ms = _kind(ob, n, 1, rows, precision)
    assert kind == 2 and 1 <= teams < rows and ms > 0
    out_m, info_m = ob.fiber_batch(xd, DT, precision=precision, persistent=False, **kw)
    assert np.array_equal(info.steps, info_m.steps) and info.done.all()
    assert len(set(info.steps.tolist())) > 5
    assert rel_l2(out, out_m.cpu().numpy()) <= (2e-6 if precision == "fp32" else 1e-13)
    for i in (0, 7, 44):
        with np.errstate(all="ignore"):
            ref = oracle_fiber(x[i], DT, real=REAL[precision], **kw)
        assert int(info.steps[i]) == ref["steps"]
        assert rel_l2(out[i], ref["out"]) <= TOL[precision]
    # the same rows from pageable host memory (threaded lanes, several chunks) give the same numbers
    out_h, info_h = ob.fiber_batch(x, DT, precision=precision, **kw)
    assert np.array_equal(out_h, out) and np.array_equal(info_h.steps, info.steps)
    # capping the number of teams changes the schedule, not the numbers (options are set on the plan and the plan is driven
    # directly: fiber_batch resets every scheduling option of the cached plan it uses)
    plan = engine.get_plan(n, 1, rows, td, torch.device("cuda", 0))
    # (cluster = 1: each team is a thread-block cluster with hardware barriers; 0: cooperative launch, barriers through L2)
    for teams_cap, placement, cluster in ((3, -1, 0), (5, 0, 0), (5, 1, 0), (0, 0, 0), (0, 1, 0), (0, -1, 1), (3, -1, 1)):
        plan.reset_schedule()                                           # (the multi-launch call above left persistent = 0)
        plan.set_option("teams", teams_cap); plan.set_option("placement", placement); plan.set_option("cluster", cluster)
        try:
            w = xd.clone()
            info_3 = plan.propagate(w, DT, **kw)
            if teams_cap:
                assert plan.last_timing()[1] == teams_cap               # waveforms in flight
        finally:
            plan.reset_schedule()
        assert np.array_equal(w.cpu().numpy(), out) and np.array_equal(info_3.steps, info.steps)
    # ... and fiber_batch does not inherit an option another caller left on the cached plan
    plan.set_option("teams", 2)
    ob.fiber_batch(xd, DT, precision=precision, persistent=True, **kw)
    assert plan.last_timing()[1] > 2 and plan.get_option("teams") == 0


def test_persistent_step_budget_and_resume(ob):
    """return_steps=True drives the kernel one step per call (max_steps=1, resume): the trajectory must equal
    the oracle's snapshots, and the final field the uninterrupted run."""
    n = 1 << 14
    x = _wave(n, 21)
    kw = dict(length=5.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, phi_max=0.02)
    ob.gv.dt = DT; ob.gv.fs = 1 / DT
    for precision in ("fp64", "fp32"):
        z, traj = ob.FIBER(ob.optical_signal(x), return_steps=True, precision=precision, **kw)
        assert _kind(ob, n, 1, 1, precision)[0] == 2
        with np.errstate(all="ignore"):
            ref = oracle_fiber(x, DT, real=REAL[precision], return_steps=True, **kw)
        assert len(z) == ref["steps"] + 1 and traj.shape == (len(z), n) and z[0] == 0.0
        np.testing.assert_allclose(z[1:], ref["z"], rtol=1e-3 if precision == "fp32" else 1e-10)
        for k in (1, len(z) // 2, len(z) - 1):
            assert rel_l2(traj[k], ref["traj"][k]) <= TOL[precision]
        full = ob.FIBER(ob.optical_signal(x), precision=precision, **kw)
        assert rel_l2(traj[-1], full.signal) <= (2e-6 if precision == "fp32" else 1e-13)


def test_single_waveform_2_20(ob):
    """BASELINE config #2 geometry (one 2^20-sample waveform = one team of 256 CTAs), a few adaptive steps."""
    n = 1 << 20
    x = _wave(n, 9, power=20e-3)
    kw = dict(length=1.2, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.01)
    for precision in ("fp64", "fp32"):
        with np.errstate(all="ignore"):
            ref = oracle_fiber(x, DT, real=REAL[precision], **kw)
        out, info = ob.fiber_batch(x[None, :], DT, precision=precision, persistent=True, **kw)
        assert _kind(ob, n, 1, 1, precision)[0] == 2
        assert int(info.steps[0]) == ref["steps"] and ref["steps"] >= 3
        assert rel_l2(out[0], ref["out"]) <= TOL[precision]


@pytest.mark.parametrize("log2n", [11, 12, 14])
def test_degenerate_step_rules(ob, log2n):
    """length = 0 (no step, field untouched), h > length (one clamped step), a sliver of a last step: same bookkeeping as the
    reference loop (devices.py:1159-1162, 1172-1173, 1195-1196) on the multi-launch schedule (2^11) and on k_wf."""
    n = 1 << log2n
    x = _wave(n, 30 + log2n)
    base = dict(alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0)
    for precision in ("fp64", "fp32"):
        out, info = ob.fiber_batch(x[None, :], DT, precision=precision, length=0.0, **base)
        assert int(info.steps[0]) == 0 and bool(info.done[0])
        assert rel_l2(out[0], x.astype(np.complex64 if precision == "fp32" else np.complex128)) == 0.0
        for kw in (dict(length=2.0, h=5.0), dict(length=1.0, h=0.3), dict(length=3.0, phi_max=100.0)):
            with np.errstate(all="ignore"):
                ref = oracle_fiber(x, DT, real=REAL[precision], **base, **kw)
            out, info = ob.fiber_batch(x[None, :], DT, precision=precision, **base, **kw)
            assert int(info.steps[0]) == ref["steps"], (precision, kw)
            assert rel_l2(out[0], ref["out"]) <= TOL[precision], (precision, kw)
            np.testing.assert_allclose(info.z[0], ref["z"][-1], rtol=1e-6 if precision == "fp32" else 1e-12)


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("variant", ["multi_cluster", "multi_tile"])
@pytest.mark.parametrize("log2n,n_pol", [(17, 1), (16, 2), (18, 1)], ids=["2^17", "2^16x2pol", "2^18"])
def test_multi_cluster_teams_match_oracle_and_flag_teams(ob, monkeypatch, log2n, n_pol, variant, precision):
    """Waveforms of 32 / 64 tiles with plan option cluster = 1 -- as ONE 16-CTA cluster with several tiles per CTA and two
    passes per column phase around the team maximum (multi_tile, the default), or as clusters of 8 with one flag hop between
    the cluster leaders (multi_cluster, SSFM_NO_MT=1) -- against the oracle and against the flag-based cooperative teams
    (cluster = 0): several adaptive steps, rows with different step counts, more rows than teams, a step budget with resume."""
    import torch
    from opticomlib_b200 import engine
    if variant == "multi_cluster":
        monkeypatch.setenv("SSFM_NO_MT", "1")
    else:
        monkeypatch.delenv("SSFM_NO_MT", raising=False)
    n = 1 << log2n
    rows = 11
    kw = dict(length=2.5, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.004)
    base = _wave(n, 40 + log2n, power=8e-3, n_pol=n_pol)
    x = np.stack([base * (1.0 + 0.07 * b) for b in range(rows)])
    td = torch.complex64 if precision == "fp32" else torch.complex128
    dev = torch.device("cuda", 0)
    with np.errstate(all="ignore"):
        refs = [oracle_fiber(x[b], DT, real=REAL[precision], **kw) for b in (0, rows - 1)]
    assert refs[0]["steps"] >= 4 and refs[1]["steps"] > refs[0]["steps"]
    outs = {}
    for cluster in (1, 0):
        plan = engine.get_plan(n, n_pol, rows, td, dev, lane=5)
        plan.set_option("cluster", cluster)
        f = torch.from_numpy(x).to(dev).to(td).contiguous()
        info = plan.propagate(f, DT, max_steps=3, **kw)
        while not info.done.all():
            info = plan.propagate(f, DT, max_steps=3, resume=True, **kw)
        assert plan.last_timing()[0] == 2
        for ref, b in zip(refs, (0, rows - 1)):
            assert int(info.steps[b]) == ref["steps"]
            assert rel_l2(f[b].cpu().numpy().reshape(ref["out"].shape), ref["out"]) <= TOL[precision]
        outs[cluster] = (f.cpu().numpy(), info.steps.copy())
        plan.set_option("cluster", -1)
    np.testing.assert_array_equal(outs[1][1], outs[0][1])
    np.testing.assert_array_equal(outs[1][0], outs[0][0])          # same arithmetic, only the synchronisation differs


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("log2n,n_pol", [(17, 1), (16, 2), (18, 1)], ids=["2^17", "2^16x2pol", "2^18"])
def test_multi_tile_cluster_teams_fixed_step(ob, log2n, n_pol, precision):
    """Fixed-step propagation of waveforms of 32 / 64 tiles as ONE 16-CTA cluster per waveform with several tiles per CTA
    (k_wf<.., TM = 3>, plan option cluster = 1; Kerr phase in the L2-resident team stash) against the oracle and bit for bit
    against the flag-based teams (cluster = 0): more rows than teams, a last step shorter than h, a step budget with resume,
    a zero-length call, DBP."""
    import torch
    from opticomlib_b200 import engine
    n = 1 << log2n
    rows = 19
    kw = dict(length=2.3, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, h=0.5)       # 5 steps, the last one 0.3 km
    base = _wave(n, 60 + log2n, power=8e-3, n_pol=n_pol)
    x = np.stack([base * (1.0 + 0.05 * b) for b in range(rows)])
    td = torch.complex64 if precision == "fp32" else torch.complex128
    dev = torch.device("cuda", 0)
    with np.errstate(all="ignore"):
        refs = [oracle_fiber(x[b], DT, real=REAL[precision], **kw) for b in (0, rows - 1)]
        ref_dbp = oracle_dbp(x[3], DT, real=REAL[precision], **kw)
    assert refs[0]["steps"] == 5
    outs = {}
    for cluster in (1, 0):
        plan = engine.get_plan(n, n_pol, rows, td, dev, lane=6)
        plan.set_option("cluster", cluster)
        f = torch.from_numpy(x).to(dev).to(td).contiguous()
        info = plan.propagate(f, DT, **kw)
        assert plan.last_timing()[0] == 2
        assert plan.last_timing()[1] >= (14 if cluster == 1 else 1)           # teams in flight: 14 clusters (+ fill teams)
        for ref, b in zip(refs, (0, rows - 1)):
            assert int(info.steps[b]) == ref["steps"]
            assert rel_l2(f[b].cpu().numpy().reshape(ref["out"].shape), ref["out"]) <= TOL[precision]
        g = torch.from_numpy(x).to(dev).to(td).contiguous()                    # the same in budgets of two steps
        info2 = plan.propagate(g, DT, max_steps=2, **kw)
        while not info2.done.all():
            info2 = plan.propagate(g, DT, max_steps=2, resume=True, **kw)
        # (a budget boundary splits the merged Kerr rotation e^{j(a+b)} into e^{ja} e^{jb}: equal to rounding, not bit for bit)
        assert rel_l2(g.cpu().numpy(), f.cpu().numpy()) <= (1e-5 if precision == "fp32" else 1e-13)
        np.testing.assert_array_equal(info2.steps, info.steps)
        z = torch.from_numpy(x).to(dev).to(td).contiguous()                    # zero length: nothing moves
        info0 = plan.propagate(z, DT, **dict(kw, length=0.0))
        assert int(info0.steps.max()) == 0
        np.testing.assert_array_equal(z.cpu().numpy(), torch.from_numpy(x).to(td).numpy())
        d = torch.from_numpy(x).to(dev).to(td).contiguous()                    # DBP = negated parameters
        plan.propagate(d, DT, length=kw["length"], alpha=-kw["alpha"], beta_2=-kw["beta_2"], beta_3=-kw["beta_3"], gamma=-kw["gamma"], h=kw["h"])
        assert rel_l2(d[3].cpu().numpy().reshape(ref_dbp["out"].shape), ref_dbp["out"]) <= TOL[precision]
        outs[cluster] = (f.cpu().numpy(), info.steps.copy())
        plan.set_option("cluster", -1)
    np.testing.assert_array_equal(outs[1][1], outs[0][1])
    np.testing.assert_array_equal(outs[1][0], outs[0][0])          # same arithmetic, only the team structure differs


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_small_cluster_fill_teams_with_adaptive_steps(ob, precision):
    """A batch large enough that the CTA slots the 16-CTA clusters leave are filled by clusters of 2 CTAs carrying 8 tiles each
    (k_wf<.., TM = 3>: two passes per column phase around the team maximum, Kerr phase in the L2-resident team stash, teams that
    stop drawing near the end of the batch): adaptive steps with diverging step counts, bit for bit against the flag-based
    teams and against the oracle."""
    import torch
    from opticomlib_b200 import engine
    n, rows = 1 << 16, (340 if precision == "fp64" else 460)           # >= 3 x (14 | 21 clusters) x 7: the batch size from which
    kw = dict(length=3.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.004)   # the small clusters are used
    base = _wave(n, 91, power=8e-3)
    scale = 1.0 + 0.7 * np.arange(rows) / rows
    td = torch.complex64 if precision == "fp32" else torch.complex128
    dev = torch.device("cuda", 0)
    x = torch.from_numpy(base).to(dev).to(td)[None, :] * torch.from_numpy(scale).to(dev).to(td)[:, None]
    x = x.contiguous()
    picks = (0, rows // 2, rows - 1)
    with np.errstate(all="ignore"):
        refs = [oracle_fiber(x[b].cpu().numpy(), DT, real=REAL[precision], **kw) for b in picks]
    assert refs[-1]["steps"] > refs[0]["steps"] >= 4
    outs = {}
    for cluster in (-1, 0):
        plan = engine.get_plan(n, 1, rows, td, dev, lane=7)
        plan.set_option("cluster", cluster)
        f = x.clone()
        info = plan.propagate(f, DT, **kw)
        kind, in_flight, _ = plan.last_timing()
        assert kind == 2
        if cluster == -1:
            assert in_flight > (27 if precision == "fp32" else 18)      # more teams than 16-CTA clusters + flag-based teams could be
        for ref, b in zip(refs, picks):
            assert int(info.steps[b]) == ref["steps"]
            assert rel_l2(f[b].cpu().numpy(), ref["out"]) <= TOL[precision]
        assert info.done.all()
        outs[cluster] = (f.cpu().numpy(), info.steps.copy())
        if cluster == -1:                                             # the same batch in budgets of 3 steps with resume: rows finish at
            g = x.clone()                                             # different calls, later calls skip the rows that are done
            info2 = plan.propagate(g, DT, max_steps=3, **kw)
            calls = 1
            while not info2.done.all():
                info2 = plan.propagate(g, DT, max_steps=3, resume=True, **kw)
                calls += 1
                assert calls < 40
            assert calls >= 3
            np.testing.assert_array_equal(info2.steps, info.steps)
            assert rel_l2(g.cpu().numpy(), outs[-1][0]) <= (1e-5 if precision == "fp32" else 1e-13)
        plan.set_option("cluster", -1)
    np.testing.assert_array_equal(outs[-1][1], outs[0][1])
    np.testing.assert_array_equal(outs[-1][0], outs[0][0])            # same arithmetic whatever the team structure
