"""Tiny k_wf runs for compute-sanitizer (memcheck / racecheck): 2^12 (one-CTA teams), 2^14 (4-CTA teams: cluster and flags)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import opticomlib_b200 as ob
from opticomlib_b200 import engine
rng = np.random.default_rng(0)
for n, rows in ((1 << 12, 3), (1 << 14, 3)):
    t = np.arange(n) / n
    x = np.sqrt(2e-3) * (1 + 0.5 * np.cos(2 * np.pi * 5 * t)) * np.exp(2j * np.pi * 3 * t) + 1e-3 * (rng.standard_normal((rows, n)) + 1j * rng.standard_normal((rows, n)))
    for prec in ('fp64', 'fp32'):
        for cluster in (1, 0):
            td = torch.complex128 if prec == 'fp64' else torch.complex64
            xt = torch.from_numpy(x).cuda().to(td)
            plan = engine.get_plan(n, 1, rows, td, torch.device('cuda', 0))
            plan.set_option('cluster', cluster)
            info = plan.propagate(xt, 1 / 160e9, length=3.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, phi_max=0.02)
            torch.cuda.synchronize()
            print(n, prec, 'cluster', cluster, 'kind', plan.last_timing()[0], 'steps', info.steps.tolist(), flush=True)
