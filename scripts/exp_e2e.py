"""e2e (pinned host -> device -> pinned host) of config #3 through fiber_batch for different chunk sizes / lane counts."""
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from opticomlib_b200 import devices, engine, workloads as wl
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x, dt, kw = wl.config_input('cfg1')
n = x.size
base = torch.from_numpy(x)
xh = torch.empty((rows, n), dtype=torch.complex128, pin_memory=True)
xh[:] = (10 ** 0.5) * base
xh *= (1 + 0.01 * torch.rand((rows, 1), dtype=torch.float64)).to(torch.complex128)
out_h = torch.empty_like(xh)
for chunk_mib, lanes in ((256, 3), (512, 3), (1024, 3), (512, 2), (512, 4), (1024, 2), (2048, 2)):
    devices.HOST_CHUNK_BYTES = chunk_mib << 20
    devices.HOST_LANES = lanes
    engine.clear_plans(); torch.cuda.empty_cache()
    devices.fiber_batch(xh, dt, precision='fp64', out=out_h, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _, info = devices.fiber_batch(xh, dt, precision='fp64', out=out_h, **kw)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    print('chunk %4d MiB lanes %d: %.1f ms  %.3e sample*steps/s' % (chunk_mib, lanes, t * 1e3, info.sample_steps(n) / t), flush=True)
