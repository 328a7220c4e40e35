"""GPU parity tests of the host path that bench.py's `e2e` number is measured on.

`fiber_batch` with PINNED input and a pinned `out=` takes `_propagate_host_pipelined` (opticomlib_b200/devices.py): one
host thread enqueues, round-robin over three streams, H2D copy -> persistent kernel (plan option "async") -> D2H copy ->
`ssfm_copy_state_async`.  With pageable NumPy input the same call takes the threaded pipeline instead, which is what
every other GPU test exercises -- so the timed path gets its own parity tests here: many rows with DIVERGING step
counts, both precisions, chunks spread over all lanes (several persistent launches in flight on different streams, the
cluster launch and the launch that fills the free CTA slots of each), a ragged last chunk, against the oracle
(oracle/ssfm_oracle.py, a restatement of reference devices.py:1137-1196) row by row.
Tolerances are BASELINE.json's: rel-L2 <= 1e-4 (fp32), <= 1e-10 (fp64), identical step counts.
"""
import numpy as np
import pytest

from oracle.ssfm_oracle import oracle_fiber, rel_l2

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "fp64": 1e-10}
REAL = {"fp32": np.float32, "fp64": np.float64}
DT = 1 / 640e9
KW = dict(length=14.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.01)


@pytest.fixture(scope="module")
def ob():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import opticomlib_b200 as ob
    return ob


def _rows(n, rows, seed=5):
    """OOK-like rows whose peak powers span 1:6, so the adaptive step counts differ from row to row."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / n
    out = np.empty((rows, n), np.complex128)
    for b in range(rows):
        power = 1e-3 * (1.0 + 5.0 * b / max(1, rows - 1))
        env = np.sqrt(power) * (0.55 + 0.45 * np.sign(np.sin(2 * np.pi * (23 + b % 7) * t + 0.1 * b)))
        env = np.convolve(env, np.ones(9) / 9, mode="same")
        out[b] = env * np.exp(2j * np.pi * 2 * t) + 2e-3 * np.sqrt(power) * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return out


def _oracle_rows(x, precision, kw):
    refs = []
    with np.errstate(all="ignore"):
        for row in x:
            refs.append(oracle_fiber(row, DT, real=REAL[precision], **kw))
    return refs


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("single_launch", [True, False], ids=["one_launch", "launch_per_chunk"])
@pytest.mark.parametrize("log2n,rows,chunk_rows", [(13, 44, 6), (16, 41, 5)], ids=["2^13x44", "2^16x41"])
def test_pinned_async_pipeline_matches_oracle(ob, monkeypatch, precision, single_launch, log2n, rows, chunk_rows):
    """The two pipelines behind fiber_batch for pinned host buffers -- ONE persistent launch that adopts waveforms as their
    chunks arrive (ssfm_propagate_streamed; the path bench.py's e2e times; same dtype on host and device) and one launch per
    chunk over three streams (the fallback, and the path for fp32 computation on complex128 host data) -- against the oracle:
    rows with diverging step counts, >= 8 chunks, a ragged last chunk."""
    import torch
    from opticomlib_b200 import devices
    n = 1 << log2n
    x = _rows(n, rows)
    refs = _oracle_rows(x, precision, KW)
    steps_ref = np.array([r["steps"] for r in refs])
    assert len(set(steps_ref.tolist())) >= 5, "the rows were meant to need different numbers of steps"

    # small chunks: >= 8 chunks (round-robin over the three lanes when every chunk is a launch), and a ragged last one
    monkeypatch.setattr(devices, "HOST_CHUNK_BYTES", chunk_rows * n * 16)
    monkeypatch.setattr(devices, "HOST_SINGLE_CHUNK_BYTES", chunk_rows * n * (8 if precision == "fp32" else 16))
    monkeypatch.setattr(devices, "HOST_SINGLE_LAUNCH", single_launch)
    calls, singles = [], []
    real, real_single = devices._propagate_host_pipelined, devices._propagate_host_single_launch

    def spy(*a, **k):
        calls.append(len(a[7]))                                   # chunks
        return real(*a, **k)

    def spy_single(*a, **k):
        res = real_single(*a, **k)
        singles.append(res is not None)
        return res

    monkeypatch.setattr(devices, "_propagate_host_pipelined", spy)
    monkeypatch.setattr(devices, "_propagate_host_single_launch", spy_single)
    tdt = torch.complex64 if precision == "fp32" else torch.complex128
    host = torch.from_numpy(x).to(tdt if single_launch else torch.complex128).pin_memory()   # (one launch: host dtype = compute dtype)
    out = torch.empty(host.shape, dtype=tdt, pin_memory=True)
    res, info = ob.fiber_batch(host, DT, precision=precision, out=out, **KW)
    if single_launch:
        assert singles == [True] and not calls, "the pinned inputs did not take the single streamed launch: %r %r" % (singles, calls)
    else:
        assert calls and calls[0] >= 8, "the pinned inputs did not take the asynchronous pipeline (or too few chunks): %r" % calls
    assert res.data_ptr() == out.data_ptr()
    np.testing.assert_array_equal(info.steps, steps_ref)
    got = res.numpy()
    worst = max(rel_l2(got[b], refs[b]["out"]) for b in range(rows))
    assert worst <= TOL[precision], worst
    np.testing.assert_allclose(info.z, [r["z"][-1] for r in refs], rtol=1e-6 if precision == "fp32" else 1e-12)
    assert info.done.all()


def test_pinned_pipeline_equals_device_resident_and_pageable(ob, monkeypatch):
    """The three ways into fiber_batch (pinned host, pageable host, device tensor) give the same rows and step counts."""
    import torch
    from opticomlib_b200 import devices
    n, rows = 1 << 14, 26
    x = _rows(n, rows, seed=9)
    monkeypatch.setattr(devices, "HOST_CHUNK_BYTES", 4 * n * 16)
    host = torch.from_numpy(x).pin_memory()
    out = torch.empty(host.shape, dtype=torch.complex128, pin_memory=True)
    a, ia = ob.fiber_batch(host, DT, precision="fp64", out=out, **KW)
    b, ib = ob.fiber_batch(x, DT, precision="fp64", **KW)                                  # pageable: threaded lanes
    c, ic = ob.fiber_batch(torch.from_numpy(x).cuda(), DT, precision="fp64", **KW)         # device resident
    np.testing.assert_array_equal(ia.steps, ib.steps)
    np.testing.assert_array_equal(ia.steps, ic.steps)
    assert rel_l2(a.numpy(), b) <= 1e-13
    assert rel_l2(a.numpy(), c.cpu().numpy()) <= 1e-13


def test_resume_and_zero_length_with_many_rows(ob):
    """More rows than teams, a step budget of 4 with `resume` (rows finish at different calls, so later calls meet rows that
    are already done) and a zero-length run with a fixed step: every skipped row must still hand the team on to the next
    row (the flag-based teams could hang on the mailbox word on these paths)."""
    import torch
    from opticomlib_b200 import engine
    n, rows = 1 << 13, 700                                       # teams in flight on a B200: ~150 for 2-CTA teams
    x = _rows(n, 64, seed=3)                                     # powers 1:6 -> step counts differ by more than a budget
    xs = np.tile(x, (rows // 64 + 1, 1))[:rows]
    dev = torch.device("cuda", 0)
    for cluster in (0, -1):
        field = torch.from_numpy(xs).to(dev)
        plan = engine.get_plan(n, 1, rows, torch.complex128, dev, lane=7)
        plan.set_option("cluster", cluster)
        info = plan.propagate(field, DT, max_steps=4, **KW)
        calls = 1
        while not info.done.all():
            info = plan.propagate(field, DT, max_steps=4, resume=True, **KW)
            calls += 1
            assert calls < 50
        assert calls >= 3
        ref, iref = ob.fiber_batch(torch.from_numpy(xs).to(dev), DT, precision="fp64", **KW)
        np.testing.assert_array_equal(info.steps, iref.steps)
        assert rel_l2(field.cpu().numpy(), ref.cpu().numpy()) <= 1e-13
        # zero length, fixed step: nothing to do for any row
        field0 = torch.from_numpy(xs).to(dev)
        info0 = plan.propagate(field0, DT, length=0.0, alpha=0.2, beta_2=-20.0, gamma=2.0, h=0.5)
        assert (info0.steps == 0).all() and info0.done.all()
        np.testing.assert_array_equal(field0.cpu().numpy(), xs)
        plan.set_option("cluster", -1)
