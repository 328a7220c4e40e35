import sys, numpy as np, torch
sys.path.insert(0, '.')
import opticomlib_b200 as ob
from oracle.ssfm_oracle import oracle_fiber, rel_l2
for log2n in (21, 22):
    n = 1 << log2n
    rng = np.random.default_rng(log2n)
    t = np.arange(n) / n
    x = (np.sqrt(1e-3) * (1 + 0.5 * np.cos(2 * np.pi * 5 * t)) * np.exp(2j * np.pi * 3 * t) + 1e-3 * (rng.standard_normal(n) + 1j * rng.standard_normal(n)))
    dt = 1 / 160e9
    for steps, kw in ((1, dict(length=1.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, h=1.0)), (2, dict(length=2.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, h=1.0)), (2, dict(length=2.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=0.0, h=1.0))):
        ref = oracle_fiber(x, dt, real=np.float64, **kw)
        for fused in (False, True):
            out, info = ob.fiber_batch(x[None, :], dt, precision='fp64', fused=fused, **kw)
            d = np.abs(out[0] - ref['out'])
            bad = np.nonzero(d > 1e-9)[0]
            print(log2n, 'gamma', kw['gamma'], 'steps', steps, 'fused', fused, 'rel', rel_l2(out[0], ref['out']), 'nbad', bad.size, bad[:8], (bad[:8] // 2048, bad[:8] % 2048) if bad.size else '')
