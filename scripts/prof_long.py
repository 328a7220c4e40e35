"""A few split steps of a 2^24-sample waveform through the staged transform on one GPU (for ncu captures of the outer / inner stages)."""
import sys, torch
sys.path.insert(0, '.')
from opticomlib_b200 import longwave as lw, workloads as wl
n = 1 << 24
dev = torch.device('cuda', 0)
plan = lw.get_long_plan(n, torch.complex128, dev)
gen = torch.Generator(device=dev); gen.manual_seed(5)
x = torch.view_as_complex(torch.randn((plan.n_outer, plan.cols, 2), dtype=torch.float64, device=dev, generator=gen) * 0.02).contiguous()
c5 = dict(wl.CONFIGS["cfg5"]["fiber"]); c5["length"] = 4.0
for i in range(2):
    info = plan.propagate(x, 1.0 / 640e9, **c5)
print('ok', int(info.steps[0]))
