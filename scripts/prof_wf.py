"""One propagation of `rows` config-#3-like waveforms through the persistent kernel (for ncu captures)."""
import sys, torch
sys.path.insert(0, '.')
from opticomlib_b200 import engine, workloads as wl
prec = sys.argv[1] if len(sys.argv) > 1 else 'fp64'
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 36
cfg = sys.argv[3] if len(sys.argv) > 3 else 'cfg1'
x, dt, kw = wl.config_input(cfg)
dev = torch.device('cuda', 0)
td = torch.complex128 if prec == 'fp64' else torch.complex64
scale = 10 ** 0.5 if cfg == 'cfg1' else 1.0
x0 = (torch.from_numpy(x).to(dev) * scale).to(td).repeat(rows, 1).contiguous()
plan = engine.get_plan(x0.shape[1], 1, rows, td, dev)
for i in range(2):
    w = x0.clone()
    info = plan.propagate(w, dt, **kw)
    print(prec, rows, cfg, 'steps', int(info.steps[0]), plan.last_timing())
