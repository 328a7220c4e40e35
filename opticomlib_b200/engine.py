"""Host-side engine: plan cache + batched propagation over PyTorch CUDA tensors.

PyTorch is plumbing here (device memory, streams, ``torch.distributed``); every numerical
operation of the hot path runs in the hand-written kernels behind the C-ABI (``_lib``).
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass

import numpy as np

from . import _lib


def _torch():
    import torch

    return torch


def require_cuda(device=None):
    """Resolve the CUDA device to run on or fail loudly (no CPU fallback)."""
    torch = _torch()
    _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("opticomlib_b200: no CUDA device visible; FIBER/DBP/LPF/BPF have no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("opticomlib_b200 runs on CUDA devices only, got %r" % (device,))
    return torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())


@dataclass
class StepInfo:
    """Per-waveform outcome of a propagation (the bookkeeping of devices.py:1155-1196)."""
    steps: np.ndarray      # int32[B]   number of steps taken
    z: np.ndarray          # float64[B] position reached [km] (values of the compute real type)
    h_next: np.ndarray     # float64[B]
    done: np.ndarray       # bool[B]
    h_log: np.ndarray | None = None   # float64[B, cap] step sizes taken (row b valid up to steps[b])

    def sample_steps(self, samples_per_waveform: int) -> int:
        return int(self.steps.astype(np.int64).sum()) * int(samples_per_waveform)


STATE_RECORD = 40      # bytes per waveform copied by ssfm_copy_state_async
_STATE_DTYPE = np.dtype([("z", "<f8"), ("h", "<f8"), ("pmax", "<u8"), ("steps", "<i4"), ("done", "<i4"), ("arrived", "<u4"),
                         ("pad", "<i4")])


def decode_state(records: np.ndarray) -> StepInfo:
    """Controller records (uint8 array of B * STATE_RECORD bytes, see ssfm_copy_state_async) -> StepInfo."""
    r = np.frombuffer(np.ascontiguousarray(records).tobytes(), dtype=_STATE_DTYPE)
    return StepInfo(r["steps"].astype(np.int32), r["z"].copy(), r["h"].copy(), r["done"].astype(bool), None)


class Plan:
    """Owner of one ``ssfm_plan_t`` (twiddles, Kerr-phase stash, controller state)."""

    def __init__(self, n, n_pol, batch, complex_dtype, device):
        torch = _torch()
        self.lib = _lib.load()
        self.n, self.n_pol, self.batch = int(n), int(n_pol), int(batch)
        self.device = require_cuda(device)
        self.cdtype = torch.complex64 if complex_dtype in (torch.complex64, np.complex64, "fp32") else torch.complex128
        self.code = _lib.SSFM_C64 if self.cdtype == torch.complex64 else _lib.SSFM_C128
        h = ctypes.c_void_p()
        _lib.check(self.lib.ssfm_plan_create(ctypes.byref(h), self.n, self.n_pol, self.batch, self.code,
                                             self.device.index))
        self.handle = h

    def set_option(self, name: str, value: int):
        _lib.check(self.lib.ssfm_plan_set_option(self.handle, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        v = ctypes.c_int64(0)
        _lib.check(self.lib.ssfm_plan_get_option(self.handle, name.encode(), ctypes.byref(v)))
        return int(v.value)

    def reset_schedule(self):
        """Every scheduling knob back to its default (a cached plan is shared by whoever asks for the same shape: an option
        left behind by one caller must not steer the next)."""
        for name, value in (("chunk_waveforms", 0), ("fused", 1), ("persistent", 1), ("teams", 0), ("cluster", -1),
                            ("placement", -1), ("burst_steps", 8), ("debug", 0), ("tw_full", -1), ("lin_sep", -1), ("l2_ahead", 0), ("async", 0)):
            self.set_option(name, value)

    def peek(self, row=0):
        """(steps, z, done) of waveform ``row`` while an asynchronous propagation is running (``ssfm_peek_state``)."""
        steps, z, done = ctypes.c_int32(0), ctypes.c_double(0.0), ctypes.c_int32(0)
        _lib.check(self.lib.ssfm_peek_state(self.handle, int(row), ctypes.byref(steps), ctypes.byref(z), ctypes.byref(done)))
        return int(steps.value), float(z.value), bool(done.value)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ssfm_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def propagate(self, field, dt, length, alpha=0.0, beta_2=0.0, beta_3=0.0, gamma=0.0, phi_max=0.01,
                  h=None, max_steps=0, resume=False, want_log=False, state_out=None) -> StepInfo:
        """In place on ``field`` (CUDA tensor [B, P, N] or [B, N], plan dtype, contiguous).

        ``state_out`` (a pinned uint8 host tensor of B * STATE_RECORD bytes): do not wait -- the call returns once the
        work is enqueued on the current stream (persistent schedule), the controller records are copied to ``state_out``
        on the same stream, and None is returned; decode with ``decode_state`` after synchronising the stream."""
        torch = _torch()
        if field.dtype != self.cdtype or not field.is_cuda or not field.is_contiguous():
            raise ValueError("field must be a contiguous CUDA tensor of dtype %s" % self.cdtype)
        if field.numel() != self.batch * self.n_pol * self.n:
            raise ValueError("field has %d elements, plan expects %d" % (field.numel(), self.batch * self.n_pol * self.n))
        prm = _lib.FiberParams(float(dt), float(length), float(alpha), float(beta_2), float(beta_3), float(gamma),
                               float(phi_max), math.nan if h is None else float(h))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            if state_out is not None:
                self.set_option("async", 1)
            try:
                _lib.check(self.lib.ssfm_propagate(self.handle, field.data_ptr(), ctypes.byref(prm), int(max_steps),
                                                   1 if resume else 0, ctypes.c_void_p(stream)))
                if state_out is not None:
                    _lib.check(self.lib.ssfm_copy_state_async(self.handle, state_out.data_ptr(), ctypes.c_void_p(stream)))
            finally:
                if state_out is not None:
                    self.set_option("async", 0)
        return None if state_out is not None else self.state(want_log)

    def propagate_streamed(self, field, ready_ptr, done_ptr, chunk_rows, dt, length, alpha=0.0, beta_2=0.0, beta_3=0.0, gamma=0.0,
                           phi_max=0.01, h=None, state_out=None) -> bool:
        """Enqueue ONE persistent launch over all rows of ``field`` that adopts row w once ``ready_ptr[0] > w`` and counts
        finished tiles per chunk in ``done_ptr`` (C-ABI ``ssfm_propagate_streamed``; both device pointers, zeroed).  Never
        blocks.  False when this geometry has no persistent kernel (nothing was enqueued)."""
        torch = _torch()
        if field.dtype != self.cdtype or not field.is_cuda or not field.is_contiguous():
            raise ValueError("field must be a contiguous CUDA tensor of dtype %s" % self.cdtype)
        if field.numel() != self.batch * self.n_pol * self.n:
            raise ValueError("field has %d elements, plan expects %d" % (field.numel(), self.batch * self.n_pol * self.n))
        prm = _lib.FiberParams(float(dt), float(length), float(alpha), float(beta_2), float(beta_3), float(gamma),
                               float(phi_max), math.nan if h is None else float(h))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.ssfm_propagate_streamed(self.handle, field.data_ptr(), ctypes.byref(prm), ctypes.c_void_p(ready_ptr),
                                                  ctypes.c_void_p(done_ptr), int(chunk_rows), ctypes.c_void_p(stream))
            if rc == _lib.SSFM_ERR_UNSUPPORTED:
                return False
            _lib.check(rc)
            if state_out is not None:
                _lib.check(self.lib.ssfm_copy_state_async(self.handle, state_out.data_ptr(), ctypes.c_void_p(stream)))
        return True

    def time_step_kernels(self, field, dt, reps=5, **fiber):
        """Average device time [ms] of (column forward, row, column inverse) over ``reps`` steps.
        Advances ``field`` -- pass a scratch copy.  Measurement hook for bench.py."""
        torch = _torch()
        prm = _lib.FiberParams(float(dt), 1.0, float(fiber.get("alpha", 0.0)), float(fiber.get("beta_2", 0.0)),
                               float(fiber.get("beta_3", 0.0)), float(fiber.get("gamma", 0.0)),
                               float(fiber.get("phi_max", 0.01)), float(fiber.get("h") or 1e-3))
        ms = (ctypes.c_float * 3)()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ssfm_time_step_kernels(self.handle, field.data_ptr(), ctypes.byref(prm), int(reps),
                                                       ms, ctypes.c_void_p(stream)))
        return [float(v) for v in ms]

    def apply_transfer(self, field, h):
        """In place: row <- ifft(fft(row) * H) for every row of ``field`` ([rows, N] CUDA tensor, plan dtype);
        ``h`` = H[N] in numpy bin order, same dtype and device (DM, devices.py:1025-1029; FBG apply step 2314-2316)."""
        torch = _torch()
        if field.dtype != self.cdtype or h.dtype != self.cdtype or not field.is_cuda or not h.is_cuda:
            raise ValueError("field and h must be CUDA tensors of dtype %s" % self.cdtype)
        if not field.is_contiguous() or not h.is_contiguous() or h.numel() != self.n:
            raise ValueError("field must be contiguous and h must hold %d bins" % self.n)
        if field.numel() != self.batch * self.n_pol * self.n:
            raise ValueError("field has %d elements, plan expects %d" % (field.numel(), self.batch * self.n_pol * self.n))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ssfm_apply_transfer(self.handle, field.data_ptr(), h.data_ptr(), ctypes.c_void_p(stream)))
        return field

    def last_timing(self):
        """(kind, teams, kernel_ms) of the last propagate: kind 2 = persistent kernel (one launch, timed with
        CUDA events on the launching stream), kind 1 = multi-launch schedule (kernel_ms = 0)."""
        kind, teams, ms = ctypes.c_int32(0), ctypes.c_int32(0), ctypes.c_float(0.0)
        _lib.check(self.lib.ssfm_get_last_timing(self.handle, ctypes.byref(kind), ctypes.byref(teams), ctypes.byref(ms)))
        return int(kind.value), int(teams.value), float(ms.value)

    def state(self, want_log=False) -> StepInfo:
        B = self.batch
        steps = np.empty(B, np.int32); z = np.empty(B, np.float64); hn = np.empty(B, np.float64)
        done = np.empty(B, np.int32)
        _lib.check(self.lib.ssfm_get_state(self.handle, steps.ctypes.data, z.ctypes.data, hn.ctypes.data,
                                           done.ctypes.data))
        log = None
        if want_log:
            cap = int(max(1, steps.max()))
            kept = self.get_option("hlog_cap")
            if cap > kept:                                   # the device keeps the first `kept` step sizes of every waveform
                import warnings
                warnings.warn("opticomlib_b200: %d steps were taken but the step-size log holds %d entries per waveform; "
                              "h_log[:, %d:] is NaN" % (cap, kept, kept), RuntimeWarning)
            log = np.full((B, cap), np.nan, np.float64)
            _lib.check(self.lib.ssfm_get_step_log(self.handle, log.ctypes.data, cap))
            for b in range(B):                               # beyond a waveform's own step count the row is zero, as before
                log[b, min(int(steps[b]), kept):kept] = 0.0
        return StepInfo(steps, z, hn, done.astype(bool), log)


_PLANS: dict = {}


_PLANS_LOCK = __import__("threading").Lock()


def get_plan(n, n_pol, batch, complex_dtype, device=None, lane=0) -> Plan:
    """Cached plan.  A plan is not thread-safe: concurrent host pipelines pass distinct ``lane`` ids."""
    torch = _torch()
    dev = require_cuda(device)
    cd = torch.complex64 if complex_dtype in (torch.complex64, np.complex64, "fp32") else torch.complex128
    key = (int(n), int(n_pol), int(batch), cd, dev.index, int(lane))
    with _PLANS_LOCK:
        pl = _PLANS.pop(key, None)
        if pl is None:
            while len(_PLANS) >= 12:  # plans own O(B*N) device memory: keep the cache small (least recently used goes first)
                # Dropped from the cache, NOT destroyed: another lane thread may be inside ssfm_propagate with this plan; the
                # handle is released by Plan.__del__ when the last reference is gone.
                _PLANS.pop(next(iter(_PLANS)))
            pl = Plan(n, n_pol, batch, cd, dev)
        _PLANS[key] = pl                                     # (re)insert at the most-recently-used end
    return pl


def launch_count() -> int:
    """Kernels launched by the extension in this process (bench.py's ``gpu_launches``)."""
    return int(_lib.load().ssfm_launch_count())


def clear_plans():
    while _PLANS:
        _PLANS.popitem()[1].close()


def filtfilt_sos(x, sos: np.ndarray, out=None):
    """Zero-phase cascaded-biquad filter of a CUDA complex128 tensor [..., N] along the last axis."""
    torch = _torch()
    lib = _lib.load()
    if x.dtype != torch.complex128 or not x.is_cuda or not x.is_contiguous():
        raise ValueError("x must be a contiguous CUDA complex128 tensor")
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    if sos.ndim != 2 or sos.shape[1] != 6:
        raise ValueError("sos must have shape (n_sections, 6)")
    y = torch.empty_like(x) if out is None else out
    n = x.shape[-1]
    rows = x.numel() // n
    stream = torch.cuda.current_stream(x.device).cuda_stream
    with torch.cuda.device(x.device):
        _lib.check(lib.ssfm_filtfilt_sos(x.data_ptr(), y.data_ptr(), rows, n, sos.ctypes.data, sos.shape[0],
                                         x.device.index, ctypes.c_void_p(stream)))
    return y


def pd_lpf(field, sos: np.ndarray, noise=None, extra_noise=None, responsivity=1.0, r_load=50.0, i_dark=0.0,
           sample_offset=0, sample_stride=1):
    """Photodetector square law + zero-phase low-pass + sampler in one pass over the field (C-ABI ``ssfm_pd_lpf``).

    ``field`` (and ``noise``): CUDA complex128 ``[rows, N]`` or ``[rows, P, N]``; ``extra_noise``: CUDA float64 ``[rows, N]``
    noise current.  Returns ``(signal, noise)`` float64 CUDA tensors ``[rows, m]`` (``noise`` is None without noise inputs)."""
    torch = _torch()
    lib = _lib.load()
    if field.dtype != torch.complex128 or not field.is_cuda or not field.is_contiguous() or field.ndim not in (2, 3):
        raise ValueError("field must be a contiguous CUDA complex128 tensor [rows, N] or [rows, P, N]")
    rows, n_pol, n = (field.shape[0], 1, field.shape[1]) if field.ndim == 2 else tuple(field.shape)
    for name, t, dt, shape in (("noise", noise, torch.complex128, tuple(field.shape)), ("extra_noise", extra_noise, torch.float64, (rows, n))):
        if t is not None and (t.dtype != dt or not t.is_cuda or not t.is_contiguous() or tuple(t.shape) != shape):
            raise ValueError("%s must be a contiguous CUDA %s tensor of shape %s" % (name, dt, shape))
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    if sos.ndim != 2 or sos.shape[1] != 6:
        raise ValueError("sos must have shape (n_sections, 6)")
    m = -(-(n - int(sample_offset)) // int(sample_stride)) if 0 <= sample_offset < n and sample_stride >= 1 else 0
    want_noise = noise is not None or extra_noise is not None
    out_s = torch.empty((rows, max(m, 0)), dtype=torch.float64, device=field.device)
    out_n = torch.empty_like(out_s) if want_noise else None
    stream = torch.cuda.current_stream(field.device).cuda_stream
    with torch.cuda.device(field.device):
        _lib.check(lib.ssfm_pd_lpf(field.data_ptr(), noise.data_ptr() if noise is not None else None,
                                   extra_noise.data_ptr() if extra_noise is not None else None, out_s.data_ptr(),
                                   out_n.data_ptr() if out_n is not None else None, rows, n_pol, n, float(responsivity),
                                   float(r_load), float(i_dark), sos.ctypes.data, sos.shape[0], int(sample_offset),
                                   int(sample_stride), field.device.index, ctypes.c_void_p(stream)))
    return out_s, out_n


def gaussian_noise(shape, sigma, seed, substream=0, device=None, mean=0.0):
    """float64 CUDA tensor of ``mean + sigma N(0,1)`` values from the extension's Philox generator (``ssfm_gaussian_noise``):
    a pure function of (seed, substream, element index)."""
    torch = _torch()
    lib = _lib.load()
    dev = require_cuda(device)
    out = torch.empty(shape, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    with torch.cuda.device(dev):
        _lib.check(lib.ssfm_gaussian_noise(out.data_ptr(), out.numel(), float(mean), float(sigma), int(seed) & (2 ** 64 - 1),
                                           int(substream) & 0xffffffff, dev.index, ctypes.c_void_p(stream)))
    return out


def edfa(field, rows, gain_db, p_ase_w, seed, out_pol=None, out=None, first_row=0):
    """``sqrt(G) field + ASE`` for ``rows`` independent noise realisations (``ssfm_edfa``; reference devices.py:921-936).
    ``field``: CUDA complex128 ``[N]`` / ``[P, N]`` (one waveform, broadcast) or ``[rows, N]`` / ``[rows, P, N]``.
    Returns CUDA complex128 ``[rows, out_pol, N]`` (``[rows, N]`` when the input has no polarisation axis and out_pol is 1).
    ``first_row``: index of the first produced row in the whole batch (chunked generation gives the same noise as one call)."""
    torch = _torch()
    lib = _lib.load()
    if field.dtype != torch.complex128 or not field.is_cuda or not field.is_contiguous():
        raise ValueError("field must be a contiguous CUDA complex128 tensor")
    rows = int(rows)
    n = field.shape[-1]
    if field.ndim == 1:
        in_rows, in_pol, flat = 1, 1, True
    elif field.ndim == 2 and field.shape[0] == rows and rows > 2:
        in_rows, in_pol, flat = rows, 1, True
    elif field.ndim == 2:
        in_rows, in_pol, flat = 1, field.shape[0], False
    elif field.ndim == 3:
        in_rows, in_pol, flat = field.shape[0], field.shape[1], False
    else:
        raise ValueError("field must have shape [N], [P, N], [rows, N] or [rows, P, N]")
    out_pol = int(out_pol) if out_pol is not None else (1 if flat else 2)
    shape = (rows, n) if (flat and out_pol == 1) else (rows, out_pol, n)
    if out is None:
        out = torch.empty(shape, dtype=torch.complex128, device=field.device)
    elif tuple(out.shape) != shape or out.dtype != torch.complex128 or not out.is_cuda or not out.is_contiguous():
        raise ValueError("out must be a contiguous CUDA complex128 tensor of shape %s" % (shape,))
    stream = torch.cuda.current_stream(field.device).cuda_stream
    with torch.cuda.device(field.device):
        _lib.check(lib.ssfm_edfa(field.data_ptr(), out.data_ptr(), rows, in_rows, in_pol, out_pol, n, float(gain_db),
                                 float(p_ase_w), int(seed) & (2 ** 64 - 1), int(first_row), field.device.index, ctypes.c_void_p(stream)))
    return out


def welch_psd(x, nperseg=None):
    """Welch PSD of every row of a CUDA complex128 tensor ``[..., N]`` (``ssfm_welch_psd``): the estimate the reference plots
    and returns (typing.py:1899-1902, utils.py:2074-2079), bins in fftshift order.  Returns a float64 CUDA tensor
    ``[..., nperseg]``."""
    torch = _torch()
    lib = _lib.load()
    if x.dtype != torch.complex128 or not x.is_cuda or not x.is_contiguous():
        raise ValueError("x must be a contiguous CUDA complex128 tensor")
    n = x.shape[-1]
    nperseg = min(2048, n) if nperseg is None else int(nperseg)
    rows = x.numel() // n
    out = torch.empty(tuple(x.shape[:-1]) + (nperseg,), dtype=torch.float64, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    with torch.cuda.device(x.device):
        for r0 in range(0, rows, 65535):                      # grid.y limit
            r1 = min(rows, r0 + 65535)
            _lib.check(lib.ssfm_welch_psd(x.view(-1, n)[r0:r1].data_ptr(), out.view(-1, nperseg)[r0:r1].data_ptr(), r1 - r0, n,
                                          nperseg, x.device.index, ctypes.c_void_p(stream)))
    return out

