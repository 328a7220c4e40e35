"""e2e (pinned host -> device -> pinned host) of config #3 through fiber_batch: chunk size / lanes / cluster teams, 3 repeats."""
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
out_h = torch.empty(xh.shape, dtype=xh.dtype, pin_memory=True)
orig_get_plan = engine.get_plan
for chunk_mib, lanes, cluster, pipe in ((256, 3, -1, "threads"), (256, 3, -1, "async"), (256, 3, -1, "async_sync"), (256, 4, -1, "async_sync"),
                                       (256, 2, -1, "async_sync"), (512, 3, -1, "async_sync"), (256, 4, -1, "threads")):
    devices.HOST_PIPELINE = pipe
    devices.HOST_CHUNK_BYTES = chunk_mib << 20
    devices.HOST_LANES = lanes
    def get_plan(*a, **k):
        p = orig_get_plan(*a, **k); p.set_option('cluster', cluster); return p
    engine.get_plan = get_plan
    engine.clear_plans(); torch.cuda.empty_cache()
    devices.fiber_batch(xh, dt, precision='fp64', out=out_h, **kw)
    ts = []
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _, info = devices.fiber_batch(xh, dt, precision='fp64', out=out_h, **kw)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(pipe, 'chunk %4d MiB lanes %d cluster %2d: %s ms  best %.3e sample*steps/s' % (chunk_mib, lanes, cluster, ' '.join('%.0f' % (t * 1e3) for t in ts), info.sample_steps(n) / min(ts)), flush=True)
