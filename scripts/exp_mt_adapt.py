"""Adaptive batches of 2^17 / 2^18-sample waveforms: multi-tile 16-CTA clusters (two passes per column phase) against the
flag-based / multi-cluster teams."""
import sys, os, torch, numpy as np
sys.path.insert(0, '.')
from opticomlib_b200 import engine, workloads as wl
dev = torch.device('cuda', 0)
x1, dt1, kw1 = wl.config_input('cfg1')
for log2n, rows in ((18, 96), (17, 192)):
    n = 1 << log2n
    reps = n // x1.shape[0]
    base = torch.from_numpy(np.tile(x1, reps)).to(dev) * 10 ** 0.5
    for prec, td in (('fp64', torch.complex128), ('fp32', torch.complex64)):
        x0 = (base.to(td).repeat(rows, 1) * (1 + 0.01 * torch.rand((rows, 1), device=dev, dtype=torch.float64)).to(td)).contiguous()
        plan = engine.get_plan(n, 1, rows, td, dev)
        for mt in ('', '1'):
            if mt: os.environ['SSFM_MT_ADAPTIVE'] = '1'
            else: os.environ.pop('SSFM_MT_ADAPTIVE', None)
            best = 1e9
            for i in range(2):
                w = x0.clone(); info = plan.propagate(w, dt1, **kw1)
                best = min(best, plan.last_timing()[2])
            print('%s 2^%d x %d adaptive, multi-tile %s (in flight %d): %.2f ms %.3e  steps %d' % (prec, log2n, rows, 'on ' if mt else 'off', plan.last_timing()[1], best,
                  info.sample_steps(n) / best * 1e3, int(info.steps[0])), flush=True)
