import sys, torch, time
sys.path.insert(0, '.')
from opticomlib_b200 import engine, workloads as wl
x, dt, kw = wl.config_input('cfg2')
dev = torch.device('cuda', 0)
for prec, td in (('fp64', torch.complex128), ('fp32', torch.complex64)):
    x0 = torch.from_numpy(x).to(dev).to(td).reshape(1, -1).contiguous()
    plan = engine.get_plan(x0.shape[1], 1, 1, td, dev)
    for fused, phi in ((0, 0.01), (1, 0.01), (1, -1.0)):
        plan.set_option('fused', fused)
        w = x0.clone()
        ms = plan.time_step_kernels(w, dt, reps=20, **{**kw, 'h': 0.01, 'phi_max': phi})
        print(prec, 'fused', fused, 'phi', phi, ['%.1f us' % (m * 1e3) for m in ms], flush=True)
    plan.set_option('fused', 1)
    for burst in (8, 32, 128):
        plan.set_option('burst_steps', burst)
        best = 1e9
        for i in range(4):
            w = x0.clone(); torch.cuda.synchronize(); t = time.perf_counter()
            info = plan.propagate(w, dt, **kw); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
        print(prec, 'burst', burst, 'steps', int(info.steps[0]), '%.2f ms  %.1f us/step' % (best * 1e3, best * 1e6 / info.steps[0]), flush=True)
