"""Multi-cluster teams (N >= 2^17): cluster=-1 (clusters of 8 + one flag hop) against cluster=0 (flag-based cooperative teams)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from opticomlib_b200 import engine, workloads as wl
from oracle.ssfm_oracle import oracle_fiber, rel_l2
dev = torch.device('cuda', 0)
x, dt, kw = wl.config_input('cfg2')
ref = None
for prec in ('fp64', 'fp32'):
    td = torch.complex128 if prec == 'fp64' else torch.complex64
    x0 = torch.from_numpy(x).to(dev).to(td).reshape(1, -1)
    plan = engine.get_plan(x0.shape[1], 1, 1, td, dev)
    outs = {}
    for cluster in (0, -1):
        plan.set_option('cluster', cluster)
        best = 1e9
        for i in range(3):
            w = x0.clone(); info = plan.propagate(w, dt, **kw)
            best = min(best, plan.last_timing()[2])
        outs[cluster] = w
        print('cfg2 %s cluster %d: %.3f ms, %d steps, %.3e sample*steps/s' % (prec, cluster, best, int(info.steps[0]), info.sample_steps(x0.shape[1]) / best * 1e3), flush=True)
    print('   flag vs multi-cluster rel-L2 %.2e' % float((outs[0] - outs[-1]).norm() / outs[0].norm()))
# DBP-like: 64 frames x 2^18, fixed h (config #4) and adaptive
base = torch.from_numpy(wl.ook_field(15, 4096, 64, 0.0)).to(dev)
rx = base.repeat(64, 1) * (1 + 0.01 * torch.rand((64, 1), device=dev, dtype=torch.float64))
c4 = wl.CFG4_RX['dbp']
plan = engine.get_plan(1 << 18, 1, 64, torch.complex128, dev)
for name, kw4 in (('fixed h', dict(length=c4['length'], alpha=-c4['alpha'], beta_2=-c4['beta_2'], beta_3=-c4['beta_3'], gamma=-c4['gamma'], h=c4['h'])),
                  ('adaptive', dict(length=20.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.02))):
    outs = {}
    for cluster in (0, -1):
        plan.set_option('cluster', cluster)
        best = 1e9
        for i in range(3):
            w = rx.clone(); info = plan.propagate(w, 1 / 640e9, **kw4)
            best = min(best, plan.last_timing()[2])
        outs[cluster] = w
        print('2^18 x 64 %s cluster %d: %.3f ms, teams %d, %.3e sample*steps/s' % (name, cluster, best, plan.last_timing()[1], info.sample_steps(1 << 18) / best * 1e3), flush=True)
    print('   flag vs multi-cluster rel-L2 %.2e' % float((outs[0] - outs[-1]).norm() / outs[0].norm()))
# parity at 2^17 against the oracle, both modes
n = 1 << 17
t = np.arange(n) / n
xw = np.sqrt(2e-3) * (0.55 + 0.45 * np.sign(np.sin(2 * np.pi * 37 * t + 0.3))) * np.exp(2j * np.pi * 3 * t)
kw17 = dict(length=6.0, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, phi_max=0.02)
ref = oracle_fiber(xw, 1 / 640e9, real=np.float64, **kw17)
plan = engine.get_plan(n, 1, 3, torch.complex128, dev)
for cluster in (0, -1):
    plan.set_option('cluster', cluster)
    w = torch.from_numpy(np.stack([xw, xw * 1.1, xw * 0.9])).to(dev)
    info = plan.propagate(w, 1 / 640e9, **kw17)
    print('2^17 cluster %d: steps %s (oracle %d), rel-L2 row 0 vs oracle %.2e' % (cluster, info.steps.tolist(), ref['steps'], rel_l2(w[0].cpu().numpy(), ref['out'])))
