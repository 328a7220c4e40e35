"""GPU parity tests of the long-waveform path (N = N0 x N_l; opticomlib_b200/longwave.py, ssfm_long_* in the C-ABI):
one GPU against the oracle and against the ordinary two-pass plan, BASELINE config #5's shape (2^26 samples) through
size-independent properties, and -- when the box has two GPUs -- the NCCL exchange against the one-GPU result."""
import os
import socket

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


def _wave(n, seed=1, power=2e-3):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / n
    env = np.sqrt(power) * (0.55 + 0.45 * np.sign(np.sin(2 * np.pi * 37 * t + 0.3)))
    env = np.convolve(env, np.ones(9) / 9, mode="same")
    return env * np.exp(2j * np.pi * 3 * t) + 2e-3 * np.sqrt(power) * (rng.standard_normal(n) + 1j * rng.standard_normal(n))


CASES = [
    ("fixed_2_14", 14, 16, dict(length=1.3, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, h=0.4)),
    ("adaptive_2_14", 14, 16, dict(length=6.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, phi_max=0.02)),
    ("adaptive_2_15_n0_64", 15, 64, dict(length=4.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, phi_max=0.02)),
    ("gamma0_2_14", 14, 32, dict(length=30.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=0.0)),
    ("nodisp_2_14", 14, 16, dict(length=10.0, alpha=0.2, gamma=2.0)),
    ("fixed_2_20", 20, 16, dict(length=0.9, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, h=0.3)),
    ("adaptive_2_20_n0_1024", 20, 1024, dict(length=0.8, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.005)),
]


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("name,log2n,n_outer,kw", CASES, ids=[c[0] for c in CASES])
def test_one_gpu_matches_oracle(ob, name, log2n, n_outer, kw, precision):
    from opticomlib_b200 import longwave as lw
    n = 1 << log2n
    x = _wave(n, log2n, power=2e-3 if log2n < 20 else 20e-3)
    with np.errstate(all="ignore"):
        ref = oracle_fiber(x, DT, real=REAL[precision], **kw)
    out, info = lw.fiber_long(x, DT, precision=precision, n_outer=n_outer, want_log=True, **kw)
    assert int(info.steps[0]) == ref["steps"] and bool(info.done[0])
    assert rel_l2(out, ref["out"]) <= TOL[precision]
    np.testing.assert_allclose(info.z[0], ref["z"][-1], rtol=1e-6 if precision == "fp32" else 1e-12)
    np.testing.assert_allclose(info.h_log[0, :ref["steps"]], ref["h"], rtol=1e-3 if precision == "fp32" else 1e-10)
    # the ordinary two-pass plan on the same input
    out2, info2 = ob.fiber_batch(x[None, :], DT, precision=precision, **kw)
    assert int(info2.steps[0]) == int(info.steps[0])
    assert rel_l2(out, out2[0]) <= (1e-5 if precision == "fp32" else 1e-12)
    if kw.get("h") is not None:
        assert info.z[0] == info2.z[0]


def test_return_steps_long(ob):
    """The trajectory of return_steps=True through the staged transform (open / close stages per step)."""
    from opticomlib_b200 import longwave as lw
    n = 1 << 14
    x = _wave(n, 4)
    for kw in (dict(length=5.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.02),
               dict(length=2.2, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, h=0.5)):
        with np.errstate(all="ignore"):
            ref = oracle_fiber(x, DT, real=np.float64, return_steps=True, **kw)
        z, traj = lw.fiber_long(x, DT, precision="fp64", n_outer=16, return_steps=True, **kw)
        assert z[0] == 0.0 and len(z) == ref["steps"] + 1 and traj.shape == (len(z), n)
        np.testing.assert_allclose(z[1:], ref["z"], rtol=1e-10)
        for k in (0, 1, len(z) // 2, len(z) - 1):
            assert rel_l2(traj[k], ref["traj"][k]) <= TOL["fp64"]


def test_two_polarisations_long(ob):
    """[2, N] through the staged transform: one plan per polarisation in lock step, one step-size sequence (devices.py:1194)."""
    from opticomlib_b200 import longwave as lw
    n = 1 << 14
    x = np.stack([_wave(n, 5), 0.4j * _wave(n, 6)[::-1]])
    for precision in ("fp64", "fp32"):
        for kw in (dict(length=5.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.02),
                   dict(length=2.2, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, h=0.5)):
            with np.errstate(all="ignore"):
                ref = oracle_fiber(x, DT, real=REAL[precision], **kw)
            out, info = lw.fiber_long(x, DT, precision=precision, n_outer=16, **kw)
            assert out.shape == x.shape and int(info.steps[0]) == ref["steps"]
            assert rel_l2(out, ref["out"]) <= TOL[precision]
            # the ordinary plan with two polarisation rows
            out2, info2 = ob.fiber_batch(x[None], DT, precision=precision, **kw)
            assert int(info2.steps[0]) == int(info.steps[0])
            assert rel_l2(out, out2[0]) <= (1e-5 if precision == "fp32" else 1e-12)
    z, traj = lw.fiber_long(x, DT, precision="fp64", n_outer=16, return_steps=True, length=2.2, alpha=0.2, beta_2=-21.27,
                            gamma=1.3, h=0.5)
    assert traj.shape == (len(z), 2, n) and len(z) == 6


def test_dbp_long(ob):
    from opticomlib_b200 import longwave as lw
    n = 1 << 14
    x = _wave(n, 3)
    kw = dict(length=5.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.02)
    with np.errstate(all="ignore"):
        ref = oracle_dbp(x, DT, real=np.float64, **kw)
    out, info = lw.dbp_long(x, DT, precision="fp64", n_outer=16, **kw)
    assert int(info.steps[0]) == ref["steps"] and rel_l2(out, ref["out"]) <= TOL["fp64"]


def test_2_23_samples_against_the_oracle(ob):
    """Beyond the two-pass limit (N = 2^23 > 2^22): FIBER itself dispatches to the long path."""
    n = 1 << 23
    x = _wave(n, 7, power=5e-3)
    kw = dict(length=0.6, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, h=0.3)
    ob.gv.dt = DT; ob.gv.fs = 1 / DT
    for precision in ("fp64", "fp32"):
        ref = oracle_fiber(x, DT, real=REAL[precision], **kw)
        out = ob.FIBER(ob.optical_signal(x), precision=precision, **kw)
        assert int(out.ssfm_info.steps[0]) == ref["steps"] == 2
        assert rel_l2(out.signal, ref["out"]) <= TOL[precision]


def test_config5_shape_properties(ob):
    """BASELINE config #5's waveform length (2^26 samples, one GPU here): size-independent properties in fp64."""
    import torch
    from opticomlib_b200 import longwave as lw
    n = 1 << 26
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    x = (torch.randn(n, 2, dtype=torch.float64, device="cuda", generator=g) * (5e-3 / 2) ** 0.5)
    x = torch.view_as_complex(x).contiguous()
    e0 = float((x.abs() ** 2).sum())
    # (1) lossless fibre conserves energy (Kerr steps are pure phase, the linear step is unitary)
    y, info = lw.fiber_long(x, DT, length=3.0, alpha=0.0, beta_2=-21.27, beta_3=0.127, gamma=1.3, h=1.0, precision="fp64")
    assert int(info.steps[0]) == 3
    assert abs(float((y.abs() ** 2).sum()) / e0 - 1) < 1e-11
    # (2) attenuation only: exact power scaling (reference test_FIBER, tests/devices_test.py:257-269)
    y, _ = lw.fiber_long(x, DT, length=10.0, alpha=0.2, precision="fp64")
    assert abs(float((y.abs() ** 2).sum()) / e0 / np.exp(-(0.2 / 4.343) * 10.0) - 1) < 1e-12
    # (3) linear propagation followed by DBP is the identity
    y, _ = lw.fiber_long(x, DT, length=40.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=0.0, precision="fp64")
    z, _ = lw.dbp_long(y, DT, length=40.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=0.0, precision="fp64")
    assert float((z - x).norm() / x.norm()) <= 1e-11
    # (4) a low-pass check of the bin map: a pure tone at bin k0 only picks up the phase of D~(w_k0) h
    k0 = 123457
    tone = torch.exp(2j * np.pi * k0 * torch.arange(n, device="cuda", dtype=torch.float64) / n).to(torch.complex128) * 1e-2
    y, _ = lw.fiber_long(tone, DT, length=2.0, alpha=0.0, beta_2=-21.27, beta_3=0.127, gamma=0.0, h=2.0, precision="fp64")
    w = 2 * np.pi * k0 / (n * DT) * 1e-12
    phase = (0.5 * -21.27 * w ** 2 + (1 / 6) * 0.127 * w ** 3) * 2.0
    assert float((y - tone * np.exp(1j * phase)).norm() / tone.norm()) <= 1e-9
    lw.clear_plans()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from opticomlib_b200 import longwave as lw
        ok = True
        n = 1 << 20
        x = _wave(n, 9, power=20e-3)
        for precision in ("fp64", "fp32"):
            for kw in (dict(length=0.9, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, h=0.3),
                       dict(length=0.8, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.005)):
                with np.errstate(all="ignore"):
                    ref = oracle_fiber(x, DT, real=REAL[precision], **kw)
                for fused in (True, False):                       # kernels store into peer memory / NCCL all-to-all
                    out, info = lw.fiber_long(x, DT, precision=precision, group=dist.group.WORLD, fused_exchange=fused, **kw)
                    plan = lw.get_long_plan(n, precision, None, dist.group.WORLD, None, fused)
                    ok &= plan.fused == fused
                    ok &= int(info.steps[0]) == ref["steps"] and rel_l2(out, ref["out"]) <= TOL[precision]
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_several_gpus_fused_and_nccl_exchange(ob, world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == list(range(world)) and all(ok for _, ok in res)
