"""k_wf experiments: cluster / cooperative variants on `rows` config-#3-like waveforms."""
import sys, torch
sys.path.insert(0, '.')
from opticomlib_b200 import engine, workloads as wl
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 288
precs = sys.argv[2].split(',') if len(sys.argv) > 2 else ['fp64', 'fp32']
x, dt, kw = wl.config_input('cfg1')
dev = torch.device('cuda', 0)
for prec in precs:
    td = torch.complex128 if prec == 'fp64' else torch.complex64
    x0 = (torch.from_numpy(x).to(dev) * 10 ** 0.5).to(td).repeat(rows, 1).contiguous()
    x0 = x0 * (1 + 0.01 * torch.rand((rows, 1), device=dev, dtype=torch.float64)).to(td)
    plan = engine.get_plan(x0.shape[1], 1, rows, td, dev)
    for cluster, teams in ((0, 0), (1, 0), (-1, 0), (1, 8), (0, 8)):
        plan.set_option('cluster', cluster); plan.set_option('teams', teams)
        best = 1e9
        for i in range(3):
            w = x0.clone()
            info = plan.propagate(w, dt, **kw)
            kind, tm, ms = plan.last_timing()
            best = min(best, ms)
        print('%s cluster %d teams %d (in flight %d): %.2f ms  %.3e sample*steps/s' % (prec, cluster, teams, tm, best, info.sample_steps(x0.shape[1]) / best * 1e3), flush=True)
