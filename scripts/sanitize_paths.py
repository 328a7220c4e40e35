"""Small runs of the other device paths for compute-sanitizer: multi-launch schedule, chirp-z lengths, long-waveform stages, filters."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import opticomlib_b200 as ob
from opticomlib_b200 import longwave as lw
rng = np.random.default_rng(0)
def wave(n, rows=2):
    t = np.arange(n) / n
    return np.sqrt(2e-3) * (1 + 0.5 * np.cos(2 * np.pi * 5 * t)) * np.exp(2j * np.pi * 3 * t) + 1e-3 * (rng.standard_normal((rows, n)) + 1j * rng.standard_normal((rows, n)))
dt = 1 / 160e9
kw = dict(length=3.0, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0)
for prec in ('fp64', 'fp32'):
    for extra in (dict(phi_max=0.02), dict(h=1.0)):
        for fused in (True, False):
            out, info = ob.fiber_batch(wave(2048), dt, precision=prec, persistent=False, fused=fused, **kw, **extra)
        out, info = ob.fiber_batch(wave(1000), dt, precision=prec, **kw, **extra)           # chirp-z
        out, info = lw.fiber_long(wave(1 << 14, 1)[0], dt, precision=prec, n_outer=16, **kw, **extra)
    print(prec, 'ok', flush=True)
x = wave(4096)
y = ob.filtfilt_batch(x, ob.devices._bessel_sos(4, 7.5e9, 640e9))
y = ob.filtfilt_batch(wave(1000), ob.devices._bessel_sos(4, 7.5e9, 640e9))
ob.gv.dt = dt; ob.gv.fs = 1 / dt
o = ob.DM(ob.optical_signal(x), D=2000.0); o = ob.DM(ob.optical_signal(wave(1000)), D=2000.0)
torch.cuda.synchronize()
print('filters/DM ok')
