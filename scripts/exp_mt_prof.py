"""clock64 phase profile (profile build) of the multi-tile cluster teams with one and two CTAs per SM."""
import sys, os, torch
sys.path.insert(0, '.')
os.environ['SSFM_MT_NOFILL'] = '1'
from opticomlib_b200 import engine, workloads as wl
dev = torch.device('cuda', 0)
base4 = torch.from_numpy(wl.ook_field(15, 4096, 64, 0.0)).to(dev)
kw4 = dict(length=80.0, alpha=-0.2, beta_2=21.27, beta_3=-0.127, gamma=-1.3, h=10.0)
n = 1 << 18
rows = 56
x = (base4[:n].repeat(rows, 1) * (1 + 0.01 * torch.rand((rows, 1), device=dev, dtype=torch.float64))).contiguous()
plan = engine.get_plan(n, 1, rows, x.dtype, dev)
plan.set_option('cluster', 1)
for cap in (7, 14):
    plan.set_option('teams', cap)
    w = x.clone(); info = plan.propagate(w, 1 / 640e9, **kw4); torch.cuda.synchronize()
    print('cap', cap, plan.last_timing(), flush=True)
