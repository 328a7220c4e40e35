"""One propagation of `rows` config-#3 waveforms between cudaProfilerStart / Stop, for `ncu --replay-mode range`: the two k_wf
launches (16-CTA clusters + small multi-tile clusters) then run CONCURRENTLY as in production, and the DRAM counters cover
the pair."""
import sys, torch
sys.path.insert(0, '.')
from opticomlib_b200 import engine, workloads as wl
prec = sys.argv[1] if len(sys.argv) > 1 else 'fp64'
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
x, dt, kw = wl.config_input('cfg1')
dev = torch.device('cuda', 0)
td = torch.complex128 if prec == 'fp64' else torch.complex64
x0 = (torch.from_numpy(x).to(dev) * 10 ** 0.5).to(td).repeat(rows, 1).contiguous()
x0 = x0 * (1 + 0.01 * torch.rand((rows, 1), device=dev, dtype=torch.float64)).to(td)
plan = engine.get_plan(x0.shape[1], 1, rows, td, dev)
w = x0.clone(); info = plan.propagate(w, dt, **kw)
w.copy_(x0); torch.cuda.synchronize()
torch.cuda.profiler.start()
info = plan.propagate(w, dt, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(prec, rows, 'sample*steps', info.sample_steps(x0.shape[1]), 'last_timing', plan.last_timing(), flush=True)
