"""Where does the single-thread host pipeline block?  Times every host call of one chunk sequence."""
import sys, time, torch, numpy as np, ctypes
sys.path.insert(0, '.')
from opticomlib_b200 import devices, engine, workloads as wl, _lib
rows_total, rows = 1024, 256
x, dt, kw = wl.config_input('cfg1')
n = x.size
xh = torch.empty((rows_total, n), dtype=torch.complex128, pin_memory=True); xh[:] = (10 ** 0.5) * torch.from_numpy(x)
out_h = torch.empty(xh.shape, dtype=xh.dtype, pin_memory=True)
dev = torch.device('cuda', 0)
rec = torch.empty(rows_total * 40, dtype=torch.uint8, pin_memory=True)
lanes = [(torch.cuda.Stream(device=dev), torch.empty((rows, n), dtype=torch.complex128, device=dev)) for _ in range(3)]
plans = [engine.get_plan(n, 1, rows, torch.complex128, dev, lane=l) for l in range(3)]
for p in plans:
    w = lanes[0][1]; w.copy_(xh[:rows]); p.propagate(w, dt, **kw)      # warm
torch.cuda.synchronize()
for trial in range(2):
    t_all = time.perf_counter()
    marks = []
    for ci in range(rows_total // rows):
        stream, xbuf = lanes[ci % 3]; plan = plans[ci % 3]
        r0, r1 = ci * rows, (ci + 1) * rows
        with torch.cuda.stream(stream):
            t0 = time.perf_counter(); xbuf.copy_(xh[r0:r1], non_blocking=True)
            t1 = time.perf_counter(); plan.propagate(xbuf, dt, state_out=rec[r0 * 40:r1 * 40], **kw)
            t2 = time.perf_counter(); out_h[r0:r1].copy_(xbuf, non_blocking=True)
            t3 = time.perf_counter()
        marks.append((t1 - t0, t2 - t1, t3 - t2))
    t_enq = time.perf_counter() - t_all
    torch.cuda.synchronize()
    t_tot = time.perf_counter() - t_all
    print('trial', trial, 'enqueue %.1f ms, total %.1f ms' % (t_enq * 1e3, t_tot * 1e3))
    for m in marks:
        print('   h2d %.2f ms  propagate %.2f ms  d2h %.2f ms' % tuple(v * 1e3 for v in m))
