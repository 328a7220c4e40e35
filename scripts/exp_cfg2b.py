import sys, torch
sys.path.insert(0, '.')
from opticomlib_b200 import engine, workloads as wl
x, dt, kw = wl.config_input('cfg2')
dev = torch.device('cuda', 0)
for prec, td in (('fp64', torch.complex128), ('fp32', torch.complex64)):
    x0 = torch.from_numpy(x).to(dev).to(td).reshape(1, -1).contiguous()
    plan = engine.get_plan(x0.shape[1], 1, 1, td, dev)
    for pers in (1, 0):
        plan.set_option('persistent', pers)
        best = 1e9
        for i in range(4):
            w = x0.clone(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); info = plan.propagate(w, dt, **kw); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print('cfg2', prec, 'persistent', pers, 'steps', int(info.steps[0]), '%.2f ms  %.1f us/step  %.3e' % (best, best * 1e3 / info.steps[0], info.sample_steps(x0.shape[1]) / best * 1e3), flush=True)
