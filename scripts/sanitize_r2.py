"""Small runs of the round-2 device paths for compute-sanitizer: k_wf with resume / skipped rows (cluster and flag-based teams),
the chunked filter pipeline with its side stream, PD -> LPF -> SAMPLER, EDFA / Philox, Welch."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import opticomlib_b200 as ob
from opticomlib_b200 import engine
rng = np.random.default_rng(0)
def wave(n, rows=2):
    t = np.arange(n) / n
    return np.sqrt(2e-3) * (1 + 0.5 * np.cos(2 * np.pi * 5 * t)) * np.exp(2j * np.pi * 3 * t) + 1e-3 * (rng.standard_normal((rows, n)) + 1j * rng.standard_normal((rows, n)))
dt = 1 / 160e9
dev = torch.device('cuda', 0)
kw = dict(length=3.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, phi_max=0.02)
for n, rows in ((1 << 13, 5), (1 << 14, 3)):
    for cluster in (1, 0):
        x = torch.from_numpy(wave(n, rows) * (1 + np.arange(rows))[:, None] ** 0.5).to(dev)
        plan = engine.get_plan(n, 1, rows, torch.complex128, dev, lane=3)
        plan.set_option('cluster', cluster)
        info = plan.propagate(x, dt, max_steps=3, **kw)
        while not info.done.all():
            info = plan.propagate(x, dt, max_steps=3, resume=True, **kw)
        info0 = plan.propagate(x, dt, length=0.0, alpha=0.2, beta_2=-20.0, gamma=2.0, h=0.5)
        print(n, 'cluster', cluster, info.steps.tolist(), info0.steps.tolist(), flush=True)
ob.gv.dt = dt; ob.gv.fs = 1 / dt; ob.gv.sps = 16
sos = ob.devices._bessel_sos(4, 7.5e9, 160e9)
x = torch.from_numpy(wave(1 << 12, 40)).to(dev)
y = engine.filtfilt_sos(x, sos)
s, nz = engine.pd_lpf(x, sos, 0.01 * x, engine.gaussian_noise((40, 1 << 12), 1e-6, 1), 0.9, 50.0, 1e-8, 8, 16)
s2, _ = engine.pd_lpf(torch.from_numpy(wave(1000, 3)).to(dev), sos)
e = ob.edfa_batch(x[0], 7, 10.0, 5.0, seed=2)
p = engine.welch_psd(x)
torch.cuda.synchronize()
print('filters / pd / edfa / welch ok', tuple(s.shape), tuple(e.shape), tuple(p.shape))
