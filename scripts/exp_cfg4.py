"""Flag-based teams: DBP of 2^18-sample frames (teams of 64 CTAs, fixed h), config #2 (one team of 256 CTAs), 2^16 with cluster=0."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
from opticomlib_b200 import engine, workloads as wl
dev = torch.device('cuda', 0)
def run(x0, dt, kw, cluster, reps=3):
    plan = engine.get_plan(x0.shape[-1], 1, x0.shape[0], x0.dtype, dev)
    plan.set_option('cluster', cluster)
    best = 1e9
    for i in range(reps):
        w = x0.clone(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); info = plan.propagate(w, dt, **kw); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, info.sample_steps(x0.shape[-1]) / best * 1e3
base4 = torch.from_numpy(wl.ook_field(15, 4096, 64, 0.0)).to(dev)
x4 = base4.repeat(64, 1) * (1 + 0.01 * torch.rand((64, 1), device=dev, dtype=torch.float64))
kw4 = dict(length=80.0, alpha=-0.2, beta_2=21.27, beta_3=-0.127, gamma=-1.3, h=10.0)
x2, dt2, kw2 = wl.config_input('cfg2')
x1, dt1, kw1 = wl.config_input('cfg1')
for prec, td in (('fp64', torch.complex128), ('fp32', torch.complex64)):
    print(prec, 'dbp 64 x 2^18: %.2f ms %.3e' % run(x4.to(td).contiguous(), 1 / 640e9, kw4, -1))
    print(prec, 'cfg2 2^20    : %.2f ms %.3e' % run(torch.from_numpy(x2).to(dev).to(td).reshape(1, -1).contiguous(), dt2, kw2, -1))
    x16 = ((10 ** 0.5) * torch.from_numpy(x1).to(dev)).to(td).repeat(288, 1).contiguous()
    print(prec, '288 x 2^16 flags: %.2f ms %.3e' % run(x16, dt1, kw1, 0))
    print(prec, '288 x 2^16 auto : %.2f ms %.3e' % run(x16, dt1, kw1, -1))
