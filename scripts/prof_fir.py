"""PD -> LPF -> SAMPLER (stride 64) on `frames` x 2^18 through the decimating FIR kernel, for ncu."""
import sys, torch
sys.path.insert(0, '.')
import opticomlib_b200 as ob
from opticomlib_b200 import engine, workloads as wl
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ob.gv(sps=64, R=10e9)
dev = torch.device('cuda', 0)
base = torch.from_numpy(wl.ook_field(15, 4096, 64, 0.0)).to(dev)
x = base[:1 << 18].repeat(frames, 1).contiguous()
sos_l = ob.devices._bessel_sos(4, 7.5e9, ob.gv.fs)
for i in range(2):
    s, nz = engine.pd_lpf(x, sos_l, None, None, 1.0, 50.0, 0.0, 32, 64)
torch.cuda.synchronize()
print('ok', tuple(s.shape))
