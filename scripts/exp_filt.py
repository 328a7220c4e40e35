import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from opticomlib_b200 import engine
from scipy import signal as sg
dev = torch.device('cuda', 0)
sos = sg.bessel(4, 20e9, 'low', fs=640e9, output='sos', norm='mag')
for rows, n in ((1024, 1 << 18), (4096, 1 << 16), (1, 1 << 20)):
    x = torch.randn(rows, n, dtype=torch.complex128, device=dev)
    y = torch.empty_like(x)
    for i in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        engine.filtfilt_sos(x, sos, out=y)
        torch.cuda.synchronize(); dtm = time.perf_counter() - t
    print('filtfilt rows %d n %d: %.3f ms  %.2f GB/s (1R+1W of complex128)' % (rows, n, dtm * 1e3, rows * n * 32 / dtm / 1e9), flush=True)
