"""Small runs of the device paths added late in round 2, for compute-sanitizer: multi-tile cluster teams (k_wf TM = 3) as the
main launch (fixed step and adaptive, resume, zero length) and as the clusters of 2 that fill the slots 16-CTA clusters leave
(forced on a small batch with SSFM_CL_CAP / SSFM_FILL_MIN), the overlap-save filter kernel k_ols (complex rows of an odd
length, in place, photodetector front end with sampler)."""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
os.environ['SSFM_CL_CAP'] = '2'; os.environ['SSFM_FILL_MIN'] = '1'
import opticomlib_b200 as ob
from opticomlib_b200 import engine
rng = np.random.default_rng(0)
def wave(n, rows=2, n_pol=1):
    t = np.arange(n) / n
    shape = (rows, n) if n_pol == 1 else (rows, n_pol, n)
    return np.sqrt(2e-3) * (1 + 0.5 * np.cos(2 * np.pi * 5 * t)) * np.exp(2j * np.pi * 3 * t) + 1e-3 * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape))
dt = 1 / 160e9
dev = torch.device('cuda', 0)
adapt = dict(length=1.5, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, phi_max=0.02)
fixed = dict(length=1.1, alpha=0.2, beta_2=-20.0, beta_3=0.1, gamma=2.0, h=0.5)
# main launch = multi-tile 16-CTA clusters: 2^16 samples x 2 polarisations = 32 tiles, 2 per CTA
x = torch.from_numpy(wave(1 << 16, 3, 2) * (1 + np.arange(3))[:, None, None] ** 0.5).to(dev)
plan = engine.get_plan(1 << 16, 2, 3, torch.complex128, dev, lane=3)
plan.set_option('cluster', 1)
for kw in (fixed, adapt):
    f = x.clone()
    info = plan.propagate(f, dt, max_steps=2, **kw)
    while not info.done.all():
        info = plan.propagate(f, dt, max_steps=2, resume=True, **kw)
    info0 = plan.propagate(f, dt, **dict(kw, length=0.0))
    print('multi-tile main', 'h' in kw, info.steps.tolist(), info0.steps.tolist(), plan.last_timing()[:2], flush=True)
# fill launch = clusters of 2 CTAs next to two 4-CTA clusters: 2^14 samples (4 tiles), 9 rows, at most 5 teams
x = torch.from_numpy(wave(1 << 14, 9) * (1 + np.arange(9))[:, None] ** 0.5).to(dev)
plan = engine.get_plan(1 << 14, 1, 9, torch.complex128, dev, lane=3)
plan.set_option('cluster', 1); plan.set_option('teams', 5)
for kw in (adapt, fixed):
    f = x.clone()
    info = plan.propagate(f, dt, **kw)
    print('multi-tile fill', 'h' in kw, info.steps.tolist(), plan.last_timing()[:2], flush=True)
ob.gv.dt = dt; ob.gv.fs = 1 / dt; ob.gv.sps = 16
sos = ob.devices._bessel_sos(4, 7.5e9, 160e9)
x = torch.from_numpy(wave(5003, 3)).to(dev)
y = engine.filtfilt_sos(x, sos)
engine.filtfilt_sos(x, sos, out=x)
e2 = torch.from_numpy(wave(9001, 2, 2)).to(dev)
s, nz = engine.pd_lpf(e2, sos, 0.01 * e2, engine.gaussian_noise((2, 9001), 1e-6, 1), 0.9, 50.0, 1e-8, 8, 16)
s2, _ = engine.pd_lpf(e2, sos)
s3, n3 = engine.pd_lpf(e2, sos, 0.01 * e2, engine.gaussian_noise((2, 9001), 1e-6, 1), 0.9, 50.0, 1e-8, 5, 40)     # stride 40: k_pd_fir<false>
s4, _ = engine.pd_lpf(e2[:, 0, :].contiguous(), sos, None, None, 1.0, 50.0, 0.0, 3, 64)                              # k_pd_fir<true>
# streamed batch: pinned host rows -> ONE launch that adopts them as they arrive -> pinned host rows
from opticomlib_b200 import devices
devices.HOST_CHUNK_BYTES = 2 * (1 << 13) * 16; devices.HOST_SINGLE_CHUNK_BYTES = 2 * (1 << 13) * 16
host = torch.from_numpy(wave(1 << 13, 9) * (1 + np.arange(9))[:, None] ** 0.5).pin_memory()
outp = torch.empty(host.shape, dtype=torch.complex128, pin_memory=True)
res, info = ob.fiber_batch(host, dt, precision='fp64', out=outp, **adapt)
print('streamed batch', info.steps.tolist(), bool(info.done.all()), flush=True)
torch.cuda.synchronize()
print('decimating FIR ok', tuple(s3.shape), tuple(s4.shape))
print('overlap-save filters ok', tuple(y.shape), tuple(s.shape), tuple(s2.shape), float((y - x).abs().max()))
