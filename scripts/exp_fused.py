import sys, torch, numpy as np
sys.path.insert(0, '.')
from opticomlib_b200 import engine, workloads as wl
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
precs = sys.argv[2].split(',') if len(sys.argv) > 2 else ['fp64', 'fp32']
x, dt, kw = wl.config_input('cfg1')
dev = torch.device('cuda', 0)
for prec in precs:
    td = torch.complex128 if prec == 'fp64' else torch.complex64
    x0 = (torch.from_numpy(x).to(dev) * 10 ** 0.5).to(td).repeat(rows, 1).contiguous()
    plan = engine.get_plan(x0.shape[1], 1, rows, td, dev)
    for name, fused, phi, pipe, l2a in (('fused/LL', 1, 0.01, 0, 0), ('fused/LL l2 296', 1, 0.01, 0, 296), ('fused/LL l2 444', 1, 0.01, 0, 444), ('fused/LL l2 600', 1, 0.01, 0, 600), ('fused/fixed', 1, -1.0, 0, 0), ('fused/fixed l2 296', 1, -1.0, 0, 296), ('fused/fixed l2 444', 1, -1.0, 0, 444)):
        plan.set_option('fused', fused); plan.set_option('pipe', pipe); plan.set_option('l2_ahead', l2a)
        w = x0.clone()
        ms = plan.time_step_kernels(w, dt, reps=4, **{**kw, 'h': 0.01, 'phi_max': phi})
        tot = ms[1] + ms[2] + (ms[0] if fused == 0 else 0)
        print('%s %-14s col_fwd %.3f row %.3f col %.3f  step %.3f ms  -> %.3e sample*steps/s' % (prec, name, ms[0], ms[1], ms[2], tot, rows * x0.shape[1] / tot * 1e3), flush=True)
