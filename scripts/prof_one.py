"""Run a few steps of one configuration through the timing hook (for ncu captures)."""
import sys, torch
sys.path.insert(0, '.')
from opticomlib_b200 import engine, workloads as wl
prec = sys.argv[1] if len(sys.argv) > 1 else 'fp64'
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 512
fused = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = sys.argv[4] if len(sys.argv) > 4 else 'cfg1'
x, dt, kw = wl.config_input(cfg)
dev = torch.device('cuda', 0)
td = torch.complex128 if prec == 'fp64' else torch.complex64
x0 = (torch.from_numpy(x).to(dev) * 10 ** 0.5).to(td).repeat(rows, 1).contiguous()
plan = engine.get_plan(x0.shape[1], 1, rows, td, dev)
plan.set_option('fused', fused)
ms = plan.time_step_kernels(x0, dt, reps=2, **{**kw, 'h': 0.01})
print(prec, rows, fused, ms)
