"""Config-#3 rows through k_wf: default (16-CTA clusters + small multi-tile clusters in the free slots) against the variants
selected by SSFM_FILL_CS (0 = flag-based fill teams); parity of the first and last row."""
import sys, os, torch, numpy as np
sys.path.insert(0, '.')
from opticomlib_b200 import engine, workloads as wl
from oracle.ssfm_oracle import oracle_fiber, rel_l2
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 576
precs = sys.argv[2].split(',') if len(sys.argv) > 2 else ['fp64']
x, dt, kw = wl.config_input('cfg1')
dev = torch.device('cuda', 0)
for prec in precs:
    td = torch.complex128 if prec == 'fp64' else torch.complex64
    x0 = (torch.from_numpy(x).to(dev) * 10 ** 0.5).to(td).repeat(rows, 1).contiguous()
    x0 = x0 * (1 + 0.01 * torch.rand((rows, 1), device=dev, dtype=torch.float64)).to(td)
    refs = {}
    with np.errstate(all='ignore'):
        for b in (0, rows - 1):
            refs[b] = oracle_fiber(x0[b].cpu().numpy(), dt, real=np.float64 if prec == 'fp64' else np.float32, **kw)
    plan = engine.get_plan(x0.shape[1], 1, rows, td, dev)
    best = 1e9
    for i in range(3):
        w = x0.clone()
        info = plan.propagate(w, dt, **kw)
        kind, tm, ms = plan.last_timing()
        best = min(best, ms)
    err = max(rel_l2(w[b].cpu().numpy(), refs[b]['out']) for b in refs)
    ok = all(int(info.steps[b]) == refs[b]['steps'] for b in refs)
    print('%s FILL_CS=%s rows %d (in flight %d): %.2f ms  %.3e sample*steps/s | rel-L2 %.2e, steps equal %s' % (
        prec, os.environ.get('SSFM_FILL_CS', 'default'), rows, tm, best, info.sample_steps(x0.shape[1]) / best * 1e3, err, ok), flush=True)
