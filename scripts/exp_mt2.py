"""Multi-tile cluster teams: throughput against the number of teams in flight (L2 footprint)."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
from opticomlib_b200 import engine, workloads as wl
dev = torch.device('cuda', 0)
frames = 128
base4 = torch.from_numpy(wl.ook_field(15, 4096, 64, 0.0)).to(dev)
kw4 = dict(length=80.0, alpha=-0.2, beta_2=21.27, beta_3=-0.127, gamma=-1.3, h=10.0)
for log2n in (18, 17):
    n = 1 << log2n
    rows = frames * (1 << (18 - log2n))
    x = (base4[:n].repeat(rows, 1) * (1 + 0.01 * torch.rand((rows, 1), device=dev, dtype=torch.float64))).contiguous()
    plan = engine.get_plan(n, 1, rows, x.dtype, dev)
    for cluster, caps in ((1, (2, 4, 6, 8, 10, 12, 14, 0)), (0, (1, 2, 3, 0))):
        plan.set_option('cluster', cluster)
        for cap in caps:
            plan.set_option('teams', cap)
            best = 1e9
            for i in range(2):
                w = x.clone(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); info = plan.propagate(w, 1 / 640e9, **kw4); e1.record(); e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            tm = plan.last_timing()[1]
            v = info.sample_steps(n) / best * 1e3
            print('2^%d cluster %d cap %2d teams %2d: %.2f ms %.3e  per team %.3e' % (log2n, cluster, cap, tm, best, v, v / tm), flush=True)
    plan.set_option('cluster', -1); plan.set_option('teams', 0)
