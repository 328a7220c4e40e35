"""Multi-tile cluster teams (k_wf TM = 3) against flag-based teams: DBP of `frames` 2^18-sample frames at fixed h (config #4),
2^17-sample frames, and the one-launch transfer path (BPF) on the same frames."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
import opticomlib_b200 as ob
from opticomlib_b200 import engine, workloads as wl
dev = torch.device('cuda', 0)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 128
def run(x0, dt, kw, cluster, reps=3):
    plan = engine.get_plan(x0.shape[-1], 1, x0.shape[0], x0.dtype, dev)
    plan.set_option('cluster', cluster)
    best = 1e9
    for i in range(reps):
        w = x0.clone(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); info = plan.propagate(w, dt, **kw); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    plan.set_option('cluster', -1)
    return best, info.sample_steps(x0.shape[-1]) / best * 1e3, plan.last_timing()[1]
base4 = torch.from_numpy(wl.ook_field(15, 4096, 64, 0.0)).to(dev)
kw4 = dict(length=80.0, alpha=-0.2, beta_2=21.27, beta_3=-0.127, gamma=-1.3, h=10.0)
ob.gv(sps=64, R=10e9)
sos_b = ob.devices._bessel_sos(4, 20e9, ob.gv.fs)
for log2n in (18, 17):
    n = 1 << log2n
    x4 = base4[:n].repeat(frames * (1 << (18 - log2n)), 1) * (1 + 0.01 * torch.rand((frames * (1 << (18 - log2n)), 1), device=dev, dtype=torch.float64))
    for prec, td in (('fp64', torch.complex128), ('fp32', torch.complex64)):
        x = x4.to(td).contiguous()
        for cluster in (0, -1):
            print(prec, 'dbp %d x 2^%d cluster %2d: %.2f ms %.3e (teams %d)' % ((x.shape[0], log2n, cluster) + run(x, 1 / 640e9, kw4, cluster)), flush=True)
    x = x4.contiguous(); y = torch.empty_like(x)
    import os
    for env in ('1', ''):
        if env: os.environ['SSFM_NO_MT'] = env
        else: os.environ.pop('SSFM_NO_MT', None)
        engine.filtfilt_sos(x, sos_b, out=y); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); engine.filtfilt_sos(x, sos_b, out=y); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print('BPF %d x 2^%d %s: %.3f ms' % (x.shape[0], log2n, 'flag teams' if env else 'multi-tile clusters', best), flush=True)
