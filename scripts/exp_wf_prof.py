import sys, torch
sys.path.insert(0, '.')
from opticomlib_b200 import engine, workloads as wl
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 72
x, dt, kw = wl.config_input('cfg1')
dev = torch.device('cuda', 0)
for prec in ('fp64', 'fp32'):
    td = torch.complex128 if prec == 'fp64' else torch.complex64
    x0 = (torch.from_numpy(x).to(dev) * 10 ** 0.5).to(td).repeat(rows, 1).contiguous()
    plan = engine.get_plan(x0.shape[1], 1, rows, td, dev)
    for cluster in (0, 1):
        plan.set_option('cluster', cluster)
        w = x0.clone(); info = plan.propagate(w, dt, **kw)
        torch.cuda.synchronize()
        print(prec, 'cluster', cluster, plan.last_timing(), flush=True)
