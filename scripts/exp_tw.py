"""Four-step twiddles of k_wf: full table from L2 (tw_full = 1) against the recurrence (tw_full = 2); config-#3 rows, parity of row 0."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
from opticomlib_b200 import engine, workloads as wl
from oracle.ssfm_oracle import oracle_fiber, rel_l2
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 576
precs = sys.argv[2].split(',') if len(sys.argv) > 2 else ['fp64', 'fp32']
x, dt, kw = wl.config_input('cfg1')
dev = torch.device('cuda', 0)
for prec in precs:
    td = torch.complex128 if prec == 'fp64' else torch.complex64
    x0 = (torch.from_numpy(x).to(dev) * 10 ** 0.5).to(td).repeat(rows, 1).contiguous()
    x0 = x0 * (1 + 0.01 * torch.rand((rows, 1), device=dev, dtype=torch.float64)).to(td)
    with np.errstate(all='ignore'):
        ref = oracle_fiber(x0[0].cpu().numpy(), dt, real=np.float64 if prec == 'fp64' else np.float32, **kw)
    plan = engine.get_plan(x0.shape[1], 1, rows, td, dev)
    for twf in (1, 2, 0):
        plan.set_option('tw_full', twf)
        best = 1e9
        for i in range(3):
            w = x0.clone()
            info = plan.propagate(w, dt, **kw)
            kind, tm, ms = plan.last_timing()
            best = min(best, ms)
        print('%s tw_full %d (in flight %d): %.2f ms  %.3e sample*steps/s | row 0: rel-L2 %.2e, steps %d vs %d' % (
            prec, twf, tm, best, info.sample_steps(x0.shape[1]) / best * 1e3, rel_l2(w[0].cpu().numpy(), ref['out']), int(info.steps[0]), ref['steps']), flush=True)
    plan.set_option('tw_full', -1)
