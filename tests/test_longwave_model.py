"""Long waveforms on CPU: the sequencer of opticomlib_b200.longwave with the NumPy stage model
(oracle/long_stages.py) -- one rank, and world_size 2 / 4 gloo process groups (layout exchange with
all_to_all_single, max all-reduce) -- against the plain oracle of the reference loop."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opticomlib_b200 import longwave as lw
from oracle.long_stages import NumpyStages
from oracle.ssfm_oracle import oracle_fiber, rel_l2

DT = 1 / 640e9
CASES = {
    "fixed": dict(length=1.3, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, h=0.4),
    "adaptive": dict(length=6.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, phi_max=0.02),
    "gamma0": dict(length=30.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=0.0),
}


def _wave(n, seed=1, power=2e-3):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / n
    env = np.sqrt(power) * (0.55 + 0.45 * np.sign(np.sin(2 * np.pi * 11 * t + 0.3)))
    env = np.convolve(env, np.ones(9) / 9, mode="same")
    return env * np.exp(2j * np.pi * 3 * t) + 2e-3 * np.sqrt(power) * (rng.standard_normal(n) + 1j * rng.standard_normal(n))


def test_split_sizes_and_step_count():
    assert lw.split_sizes(1 << 26, 8) == (256, 1 << 18)
    assert lw.split_sizes(1 << 23, 1) == (32, 1 << 18)
    assert lw.split_sizes(1 << 30, 8) == (2048, 1 << 19)
    for n, g in ((1 << 12, 1), (1 << 14, 2), (1 << 20, 8), (1 << 28, 4)):
        n0, nl = lw.split_sizes(n, g)
        assert n0 * nl == n and n0 % g == 0 and nl % g == 0 and 16 <= n0 <= 2048 and 256 <= nl <= (1 << 22)
    with pytest.raises(ValueError):
        lw.split_sizes(3000)
    assert lw.fixed_step_count(50, 0.1, "fp32") == 501          # the float32 sliver step of the reference
    assert lw.fixed_step_count(50, 0.1, "fp64") == 500
    assert lw.fixed_step_count(1000, 1.0, "fp64") == 1000
    x = np.arange(64)
    cols = [lw.local_columns(x, 4, 2, r) for r in range(2)]
    assert np.array_equal(np.concatenate(cols, axis=1).reshape(-1), x)


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("real,tol", [(np.float64, 1e-11), (np.float32, 2e-5)])
def test_one_rank_model_matches_the_oracle(name, real, tol):
    n = 1 << 12
    x = _wave(n)
    kw = CASES[name]
    with np.errstate(all="ignore"):
        ref = oracle_fiber(x, DT, real=real, **kw)
    cd = torch.complex64 if real is np.float32 else torch.complex128
    plan = lw.LongPlan(n, cd, stages=NumpyStages(n, 16, 1, 0, real), n_outer=16)
    mine = torch.from_numpy(np.ascontiguousarray(lw.local_columns(x, 16, 1, 0))).to(cd).contiguous()
    info = plan.propagate(mine, DT, want_log=True, **kw)
    assert int(info.steps[0]) == ref["steps"] and bool(info.done[0])
    assert rel_l2(mine.numpy().reshape(-1), ref["out"]) <= tol
    np.testing.assert_allclose(info.h_log[0], ref["h"], rtol=1e-3 if real is np.float32 else 1e-10)


def test_trajectory_callback_forces_open_close_stages():
    """on_step (return_steps) on the NumPy stage model: fixed and adaptive runs, snapshots against the oracle's trajectory."""
    n = 1 << 12
    x = _wave(n, seed=2)
    for name in ("fixed", "adaptive"):
        kw = CASES[name]
        with np.errstate(all="ignore"):
            ref = oracle_fiber(x, DT, real=np.float64, return_steps=True, **kw)
        plan = lw.LongPlan(n, torch.complex128, stages=NumpyStages(n, 16, 1, 0, np.float64), n_outer=16)
        mine = torch.from_numpy(np.ascontiguousarray(lw.local_columns(x, 16, 1, 0))).contiguous()
        zs, snaps = [], []
        info = plan.propagate(mine, DT, on_step=lambda f, st: (zs.append(float(st.z[0])), snaps.append(f.numpy().reshape(-1).copy())), **kw)
        assert len(snaps) == ref["steps"] == int(info.steps[0])
        np.testing.assert_allclose(zs, ref["z"], rtol=1e-12)
        for k in (0, len(snaps) // 2, len(snaps) - 1):
            assert rel_l2(snaps[k], ref["traj"][k + 1]) <= 1e-11


def test_two_polarisations_share_one_step_sequence():
    """Two plans (one per polarisation) in lock step, maxima combined before the controller: the 2-pol oracle."""
    n, n0 = 1 << 12, 16
    x = np.stack([_wave(n, seed=3), 0.4j * _wave(n, seed=4)[::-1]])
    for name in ("adaptive", "fixed"):
        kw = CASES[name]
        with np.errstate(all="ignore"):
            ref = oracle_fiber(x, DT, real=np.float64, **kw)
        plans = [lw.LongPlan(n, torch.complex128, stages=NumpyStages(n, n0, 1, 0, np.float64), n_outer=n0) for _ in range(2)]
        mine = [torch.from_numpy(np.ascontiguousarray(lw.local_columns(x[p], n0, 1, 0))).contiguous() for p in range(2)]
        info = lw.propagate_together(plans, mine, DT, **kw)
        assert int(info.steps[0]) == ref["steps"]
        out = np.stack([m.numpy().reshape(-1) for m in mine])
        assert rel_l2(out, ref["out"]) <= 1e-11


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, n0 = 1 << 13, 16
        x = _wave(n, seed=5)
        ok = True
        for name, kw in CASES.items():
            with np.errstate(all="ignore"):
                ref = oracle_fiber(x, DT, real=np.float64, **kw)
            plan = lw.LongPlan(n, torch.complex128, group=dist.group.WORLD, n_outer=n0,
                               stages=NumpyStages(n, n0, world, rank, np.float64))
            mine = torch.from_numpy(np.ascontiguousarray(lw.local_columns(x, n0, world, rank))).contiguous()
            info = plan.propagate(mine, DT, **kw)
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            out = torch.cat(parts, dim=1).reshape(-1).numpy()
            ok &= int(info.steps[0]) == ref["steps"] and rel_l2(out, ref["out"]) <= 1e-11
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_ranks_exchange_layouts_under_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == list(range(world)) and all(ok for _, ok in res)
