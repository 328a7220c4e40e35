"""opticomlib_b200 -- B200-native split-step Fourier engine behind opticomlib's FIBER / DBP / LPF / BPF.

    from opticomlib_b200 import gv, optical_signal, FIBER, DBP, LPF, BPF
    # or, next to the reference:  import opticomlib_b200; opticomlib_b200.install()

See DESIGN.md for the kernels and INTEGRATION.md for the C-ABI (include/ssfm_b200.h).
"""
from .typing import NULL, gv, electrical_signal, optical_signal
from .utils import tic, toc
from .devices import (FIBER, DBP, LPF, BPF, DM, PD, fiber_batch, dbp_batch, filtfilt_batch, transfer_batch, edfa_batch,
                      edfa_fiber_batch, pd_lpf_batch, psd_batch, install, uninstall)
from .longwave import fiber_long, dbp_long
from .batch import optical_batch, electrical_batch

__version__ = "0.1.0"
__all__ = ["NULL", "gv", "electrical_signal", "optical_signal", "tic", "toc", "FIBER", "DBP", "LPF", "BPF", "DM", "PD",
           "edfa_batch", "edfa_fiber_batch", "pd_lpf_batch", "psd_batch", "optical_batch", "electrical_batch", "fiber_batch", "transfer_batch", "dbp_batch", "filtfilt_batch", "fiber_long", "dbp_long", "install", "uninstall"]
