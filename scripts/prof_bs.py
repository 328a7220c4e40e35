"""A batch whose length is no power of two (64 x 60000 samples, chirp-z transforms of 2^17 points) for ncu captures of the k_bs_* kernels."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
from opticomlib_b200 import engine, workloads as wl
import opticomlib_b200 as ob
x, dt, kw = wl.config_input('cfg1')
dev = torch.device('cuda', 0)
n = 60000
x0 = (torch.from_numpy(x[:n].copy()).to(dev) * 10 ** 0.5).repeat(64, 1).contiguous()
for i in range(2):
    out, info = ob.fiber_batch(x0.clone(), dt, precision='fp64', **dict(kw, length=5.0))
torch.cuda.synchronize()
print('ok', int(info.steps[0]))
