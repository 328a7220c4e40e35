"""Filter timings (CUDA events, best of 3 after a warm-up): BPF and PD -> LPF at 1 / 16 / 128 / 1024 frames of 2^18, and one 2^16 row."""
import sys, torch
sys.path.insert(0, '.')
import opticomlib_b200 as ob
from opticomlib_b200 import engine, workloads as wl
ob.gv(sps=64, R=10e9)
dev = torch.device('cuda', 0)
base = torch.from_numpy(wl.ook_field(15, 4096, 64, 0.0)).to(dev)
sos_b = ob.devices._bessel_sos(4, 20e9, ob.gv.fs)
sos_l = ob.devices._bessel_sos(4, 7.5e9, ob.gv.fs)
def timed(fn):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for frames, n in ((1, 1 << 16), (1, 1 << 18), (16, 1 << 18), (128, 1 << 18), (1024, 1 << 18), (4096, 1 << 16), (512, 1 << 16)):
    x = base[:n].repeat(frames, 1).contiguous()
    y = torch.empty_like(x)
    t_b = timed(lambda: engine.filtfilt_sos(x, sos_b, out=y))
    t_l = timed(lambda: engine.pd_lpf(x, sos_l, None, None, 1.0, 50.0, 0.0, 32, 64))
    print('%5d x 2^%d: BPF %.3f ms, PD+LPF+SAMPLER %.3f ms' % (frames, n.bit_length() - 1, t_b, t_l), flush=True)
