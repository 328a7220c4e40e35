"""Drop-in ``FIBER`` / ``DBP`` / ``LPF`` / ``BPF`` running on the B200 kernels.

Same names, argument meaning, defaults, return types and error behaviour as the reference
(``opticomlib/devices.py``: FIBER 1038-1206, DBP 1209-1283, LPF 1286-1375, BPF 788-826).  The
functions accept this package's signal objects and, by duck typing, the reference's own
``optical_signal`` / ``electrical_signal`` objects, and return an object of the same class as the
input.  ``install()`` rebinds the reference's module attributes so existing scripts use this path.

Additive keyword-only arguments (not in the reference): ``precision`` ('fp32' = the algorithm as
shipped, complex64 result; 'fp64' = the dtype-lifted algorithm, complex128 result) and ``device``.
Batch entry points ``fiber_batch`` / ``dbp_batch`` / ``filtfilt_batch`` work on ``[B,(P,)N]`` arrays
or CUDA tensors; the reference has no batch API.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import engine
from .typing import NULL, electrical_signal, gv, optical_signal
from .utils import tic, toc

DEFAULT_PRECISION = "fp32"   # the reference computes FIBER in float32/complex64 (devices.py:1137-1147)


# ---- duck typing of signal objects -----------------------------------------------------------
def _kind(obj):
    """'optical' | 'electrical' | None for this package's or the reference's signal classes."""
    if isinstance(obj, optical_signal):
        return "optical"
    if isinstance(obj, electrical_signal):
        return "electrical"
    for klass in type(obj).__mro__:
        if klass.__module__.startswith("opticomlib") and klass.__name__ in ("optical_signal", "electrical_signal"):
            return "optical" if klass.__name__ == "optical_signal" else "electrical"
    return None


def _gv_of(obj):
    """The ``gv`` the object's class reads (the reference's when given a reference object)."""
    mod = sys.modules.get(type(obj).__module__)
    return getattr(mod, "gv", gv) if mod is not None else gv


def _is_null(x):
    return x is NULL or type(x).__name__ in ("NULLType", "_NullType")


def _complex_dtype(precision):
    torch = engine._torch()
    p = (precision or DEFAULT_PRECISION).lower()
    if p in ("fp32", "float32", "single", "complex64"):
        return torch.complex64, np.complex64
    if p in ("fp64", "float64", "double", "complex128"):
        return torch.complex128, np.complex128
    raise ValueError("precision must be 'fp32' or 'fp64'")


def _to_device(a: np.ndarray, tdtype, dev):
    """Host array -> contiguous CUDA tensor of dtype ``tdtype`` (one H2D copy, cast on the device)."""
    torch = engine._torch()
    a = np.ascontiguousarray(a)
    if not np.iscomplexobj(a):
        a = a.astype(np.complex128)
    t = torch.from_numpy(a).to(dev, non_blocking=False)
    return t.to(tdtype).contiguous()


# ---- FIBER / DBP --------------------------------------------------------------------------------
def fiber_batch(field, dt, length, alpha=0.0, beta_2=0.0, beta_3=0.0, gamma=0.0, phi_max=0.01, h=None, *,
                precision=None, device=None, want_log=False, chunk_waveforms=None, inplace=False, fused=True,
                persistent=None, out=None):
    """Propagate a batch ``field[B, N]`` or ``field[B, P, N]`` (NumPy array or CUDA tensor).

    Rows are independent waveforms, each with its own step-size sequence (the global max of
    devices.py:1194 is taken over the P rows of one waveform).  Returns ``(out, StepInfo)``; ``out``
    is a CUDA tensor when a CUDA tensor was given, a host tensor for a host tensor (pass a pinned ``out=`` to
    avoid the allocation) and a NumPy array for a NumPy array.  Host inputs are streamed through the device in
    chunks so that the PCIe copies overlap the propagation.

    Scheduling knobs (results agree to rounding): ``persistent`` (default on: the whole propagation is one
    persistent kernel that keeps the waveforms in flight L2-resident; off = two launches per step),
    ``fused`` / ``chunk_waveforms`` (multi-launch schedule only).
    """
    torch = engine._torch()
    tdtype, ndtype = _complex_dtype(precision)
    as_tensor = torch.is_tensor(field)
    on_host = not (as_tensor and field.is_cuda)
    if on_host:
        dev = engine.require_cuda(device)
        host = field if as_tensor else torch.from_numpy(np.ascontiguousarray(field))
        if not host.is_complex():
            host = host.to(torch.complex128)
        res, info = _propagate_host_streamed(host.contiguous(), out, tdtype, dev, want_log, chunk_waveforms, fused,
                                             persistent, (dt, length, alpha, beta_2, beta_3, gamma, phi_max, h))
        return (res if as_tensor else res.numpy()), info
    dev = engine.require_cuda(field.device)
    x = field.to(dtype=tdtype).contiguous()
    if not inplace and x.data_ptr() == field.data_ptr():
        x = x.clone()
    if x.ndim not in (2, 3):
        raise ValueError("field must have shape [B, N] or [B, P, N]")
    B, P, N = (x.shape[0], 1, x.shape[1]) if x.ndim == 2 else tuple(x.shape)
    _check_length(N)
    plan = engine.get_plan(N, P, B, tdtype, dev)
    _set_schedule(plan, fused, persistent, chunk_waveforms)
    info = plan.propagate(x, dt, length, alpha, beta_2, beta_3, gamma, phi_max, h, want_log=want_log)
    return x, info


def _set_schedule(plan, fused=True, persistent=None, chunk_waveforms=None):
    plan.reset_schedule()                                  # cached plans are shared: no option survives from an earlier caller
    plan.set_option("chunk_waveforms", 0 if chunk_waveforms is None else int(chunk_waveforms))
    plan.set_option("fused", 1 if fused else 0)
    plan.set_option("persistent", int(bool(fused)) if persistent is None else int(bool(persistent)))


MAX_TWO_PASS = 1 << 22          # longest waveform of one two-pass transform; longer ones are split N0 x N_l (longwave.py)
MAX_CHIRP = 1 << 21             # longest waveform whose length is not a power of two (chirp-z transforms of 2^22 points)


def _check_length(n):
    """The reference accepts any N (numpy.fft).  Here: any N in [2, 2^21]; powers of two up to 2^30.  Anything else is refused
    up front with the supported lengths spelled out, not from inside the plan layer."""
    n = int(n)
    pow2 = n >= 2 and (n & (n - 1)) == 0
    if n < 2 or (not pow2 and n > MAX_CHIRP) or n > (1 << 30):
        raise ValueError("opticomlib_b200 supports waveforms of 2 ... 2^21 samples of any length and power-of-two lengths up to 2^30; "
                         "got %d samples%s" % (n, "" if n < 2 or n > (1 << 30) else " (pad or crop to a power of two, or to at most 2^21 samples)"))


HOST_PIPELINE = "auto"          # "async": ONE host thread enqueues, per chunk and round-robin over HOST_LANES streams, H2D copy ->
                                # persistent kernel (plan option "async": the call does not wait) -> D2H copy -> copy of the
                                # controller records; nothing blocks until the final synchronisation.  Needs PINNED input and
                                # output (a pageable copy blocks the enqueueing thread).  "threads": one host thread + stream per
                                # lane, blocking calls (pageable buffers: the blocking copies of one lane overlap the others).
                                # "auto" = async for pinned buffers, threads otherwise.  Measured on B200 (scripts/exp_e2e.py,
                                # 4096 x 2^16 fp64, pinned): async 392 ms per propagation, +-1 ms (device-resident: 387 ms);
                                # threads 403 ... 654 ms.
HOST_LANES = 3                  # concurrent host->device->host pipelines (threads + streams) of the host path
HOST_CHUNK_BYTES = 256 << 20    # largest chunk of rows on the device
HOST_MIN_CHUNKS = 12            # ... and at least this many chunks when the batch allows it (see host_chunk_rows)


# Pinned host batches: one persistent launch that adopts waveforms as their chunks arrive (set to False to pipeline one launch
# per chunk as before)
HOST_SINGLE_LAUNCH = os.environ.get("SSFM_HOST_SINGLE_LAUNCH", "1") != "0"
HOST_SINGLE_CHUNK_BYTES = 32 << 20        # no launch per chunk there, so its chunks are small (at most 256 of them)


def host_chunk_rows(B, P, N, tdtype):
    """Rows per chunk of the host pipelines.  Only the first chunk's H2D copy and the last chunk's D2H copy are exposed
    (everything else overlaps a propagation), so chunks should be SMALL -- at least HOST_MIN_CHUNKS of them -- but a chunk
    is one persistent launch whose teams adopt its rows dynamically, so it should still hold a few rows per team in flight
    (one team = P*N/4096 CTAs; a B200 holds 296 fp64 / 444 fp32 CTAs); chunks of different lanes run concurrently, which fills
    the tail of one launch with the head of the next.  Round 1 used 256 MiB chunks: 3 chunks per GPU at 512 rows per GPU,
    171 MiB exposed each way, 5.9x instead of 7.8x at 8 GPUs."""
    torch = engine._torch()
    row_bytes = P * N * 16
    cap = max(1, HOST_CHUNK_BYTES // row_bytes)
    slots = 296 if tdtype == torch.complex128 else 444
    in_flight = max(1, (slots * 4096) // max(4096, P * N))
    rows = max(min(B, 2 * in_flight), -(-B // HOST_MIN_CHUNKS))
    rows = max(1, min(B, cap, rows))
    if B > rows:
        rows = -(-B // max(HOST_LANES, -(-B // rows)))            # even chunks, at least HOST_LANES of them
    return rows


def _propagate_host_streamed(host, out, tdtype, dev, want_log, chunk_waveforms, fused, persistent, args):
    """Host buffers in, host buffers out: rows are cut into chunks and `HOST_LANES` worker threads, each with
    its own CUDA stream and plan, run  H2D copy -> cast -> split-step loop -> D2H copy  so that the PCIe
    transfers of one chunk overlap the propagation of another (the C call releases the GIL)."""
    import threading
    torch = engine._torch()
    if host.ndim not in (2, 3):
        raise ValueError("field must have shape [B, N] or [B, P, N]")
    B = host.shape[0]
    P, N = (1, host.shape[1]) if host.ndim == 2 else (host.shape[1], host.shape[2])
    _check_length(N)
    if out is None:
        out = torch.empty(host.shape, dtype=tdtype, pin_memory=host.is_pinned())
    elif out.shape != host.shape or out.dtype != tdtype or out.is_cuda:
        raise ValueError("out must be a host tensor of shape %s and dtype %s" % (tuple(host.shape), tdtype))
    rows = host_chunk_rows(B, P, N, tdtype)
    chunks = [(r0, min(B, r0 + rows)) for r0 in range(0, B, rows)]
    lanes = min(HOST_LANES, len(chunks))
    pipe = HOST_PIPELINE
    if pipe == "auto":
        pipe = "async" if (host.is_pinned() and out.is_pinned()) else "threads"
    if pipe != "threads" and not want_log and fused and persistent in (None, True):
        if pipe == "async" and host.dtype == tdtype and len(chunks) > 1 and HOST_SINGLE_LAUNCH and not chunk_waveforms:
            res = _propagate_host_single_launch(host, out, tdtype, dev, args, chunks)
            if res is not None:
                return res
        return _propagate_host_pipelined(host, out, tdtype, dev, chunk_waveforms, args, rows, chunks, lanes, pipe)
    steps = np.zeros(B, np.int32); z = np.zeros(B); hn = np.zeros(B); done = np.zeros(B, bool)
    logs, errors = {}, []

    def worker(lane):
        try:
            with torch.cuda.device(dev):
                stream = torch.cuda.Stream(device=dev)
                with torch.cuda.stream(stream):
                    for ci in range(lane, len(chunks), lanes):
                        r0, r1 = chunks[ci]
                        x = host[r0:r1].to(dev, non_blocking=True).to(tdtype).contiguous()
                        plan = engine.get_plan(N, P, r1 - r0, tdtype, dev, lane=lane)
                        _set_schedule(plan, fused, persistent, chunk_waveforms)
                        info = plan.propagate(x, *args, want_log=want_log)
                        out[r0:r1].copy_(x, non_blocking=True)
                        steps[r0:r1], z[r0:r1], hn[r0:r1], done[r0:r1] = info.steps, info.z, info.h_next, info.done
                        if want_log:
                            logs[ci] = info.h_log
                        stream.synchronize()
        except Exception as e:                                     # surfaced by the caller's thread
            errors.append(e)

    if lanes == 1:
        worker(0)
    else:
        threads = [threading.Thread(target=worker, args=(l,)) for l in range(lanes)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
    h_log = None
    if want_log:
        width = max(l.shape[1] for l in logs.values())
        h_log = np.zeros((B, width))
        for ci, (r0, r1) in enumerate(chunks):
            h_log[r0:r1, :logs[ci].shape[1]] = logs[ci]
    return out, engine.StepInfo(steps, z, hn, done, h_log)


def single_launch_chunks(B, row_bytes):
    """Row ranges of the copies around the single streamed launch: ~HOST_SINGLE_CHUNK_BYTES each, at most 256 of them (the
    kernel is not launched per chunk, so only the first copy in and the last copy out are exposed)."""
    rows = max(1, HOST_SINGLE_CHUNK_BYTES // max(1, row_bytes), -(-B // 256))
    return rows, [(r0, min(B, r0 + rows)) for r0 in range(0, B, rows)]


def _propagate_host_single_launch(host, out, tdtype, dev, args, chunks):
    """Pinned host buffers in and out, ONE persistent launch for the whole batch (C-ABI ``ssfm_propagate_streamed``): the kernel
    is enqueued first and adopts a waveform as soon as the host-to-device copy of its chunk has been flagged (a stream memory
    operation behind every copy); a third stream copies a chunk back as soon as the kernel has counted all of its tiles as
    final.  The copies of the whole batch overlap its propagation, and -- unlike one launch per chunk -- the teams never drain
    at a chunk boundary and the small multi-tile clusters see the whole batch.  Returns None (nothing enqueued) when the
    geometry has no persistent kernel or the driver lacks stream memory operations: the caller then pipelines chunk by chunk."""
    import ctypes
    torch = engine._torch()
    lib = engine._lib.load()
    B = host.shape[0]
    P, N = (1, host.shape[1]) if host.ndim == 2 else (host.shape[1], host.shape[2])
    if (P * N) % 4096 or B * P * N * host.element_size() > (64 << 30):
        return None
    tiles = (P * N) // 4096
    # no launch per chunk here, so the chunks can be small: ~32 MiB each (at most 256 of them) -- only the first copy in and the
    # last copy out are exposed
    rows, chunks = single_launch_chunks(B, P * N * host.element_size())
    rec = torch.empty(B * engine.STATE_RECORD, dtype=torch.uint8, pin_memory=True)
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream(dev)
        x = torch.empty(tuple(host.shape), dtype=tdtype, device=dev)
        flags = torch.zeros(1 + len(chunks), dtype=torch.int32, device=dev)      # [ready | done per chunk]
        plan = engine.get_plan(N, P, B, tdtype, dev, lane=0)
        _set_schedule(plan, True, True, 0)
        h2d, d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        ready_ptr, done_ptr = flags.data_ptr(), flags.data_ptr() + 4
        reset = torch.cuda.Event(); reset.record(main)                        # buffer and zeroed flags exist; recorded BEFORE the kernel,
        h2d.wait_event(reset); d2h.wait_event(reset)                          # which waits for the copies below
        # the driver must offer stream memory operations BEFORE a kernel that depends on them is enqueued
        if (lib.ssfm_stream_write_u32(ctypes.c_void_p(h2d.cuda_stream), ctypes.c_void_p(ready_ptr), 0) != 0 or
                lib.ssfm_stream_wait_geq_u32(ctypes.c_void_p(d2h.cuda_stream), ctypes.c_void_p(ready_ptr), 0) != 0):
            return None
        if not plan.propagate_streamed(x, ready_ptr, done_ptr, rows, *args, state_out=rec):
            return None
        try:
            with torch.cuda.stream(h2d):
                for r0, r1 in chunks:
                    x[r0:r1].copy_(host[r0:r1], non_blocking=True)
                    engine._lib.check(lib.ssfm_stream_write_u32(ctypes.c_void_p(h2d.cuda_stream), ctypes.c_void_p(ready_ptr), r1))
            with torch.cuda.stream(d2h):
                for ci, (r0, r1) in enumerate(chunks):
                    engine._lib.check(lib.ssfm_stream_wait_geq_u32(ctypes.c_void_p(d2h.cuda_stream), ctypes.c_void_p(done_ptr + 4 * ci),
                                                                   (r1 - r0) * tiles))
                    out[r0:r1].copy_(x[r0:r1], non_blocking=True)
        except BaseException:
            with torch.cuda.stream(h2d):                                      # never leave the kernel waiting for rows that will not come
                flags[:1].fill_(B)
            h2d.synchronize(); main.synchronize()
            raise
        h2d.synchronize(); d2h.synchronize(); main.synchronize()
    return out, engine.decode_state(rec.numpy())


def _propagate_host_pipelined(host, out, tdtype, dev, chunk_waveforms, args, rows, chunks, lanes, pipe="async"):
    """One host thread, `lanes` CUDA streams: per chunk  H2D copy -> (cast) -> persistent kernel -> D2H copy -> copy of the
    controller records, all enqueued without waiting (plan option "async"), so the PCIe transfers of one chunk overlap
    the propagation of another and the device never waits for the host.  (If a plan has to use the multi-launch schedule its
    call blocks; the result is the same.)"""
    torch = engine._torch()
    B = host.shape[0]
    P, N = (1, host.shape[1]) if host.ndim == 2 else (host.shape[1], host.shape[2])
    rec = torch.empty(B * engine.STATE_RECORD, dtype=torch.uint8, pin_memory=True)
    with torch.cuda.device(dev):
        ln = []
        for _ in range(lanes):
            ln.append((torch.cuda.Stream(device=dev),
                       torch.empty((rows,) + tuple(host.shape[1:]), dtype=host.dtype, device=dev) if host.dtype != tdtype else None,
                       torch.empty((rows,) + tuple(host.shape[1:]), dtype=tdtype, device=dev)))
        torch.cuda.current_stream(dev).synchronize()            # the lane buffers exist before any lane stream touches them
        for ci, (r0, r1) in enumerate(chunks):
            stream, stage, xbuf = ln[ci % lanes]
            m = r1 - r0
            if pipe == "async_sync":
                stream.synchronize()
            with torch.cuda.stream(stream):
                x = xbuf[:m]
                if stage is not None:
                    stage[:m].copy_(host[r0:r1], non_blocking=True)
                    x.copy_(stage[:m])                          # cast to the compute dtype on the device
                else:
                    x.copy_(host[r0:r1], non_blocking=True)
                plan = engine.get_plan(N, P, m, tdtype, dev, lane=ci % lanes)
                _set_schedule(plan, True, True, chunk_waveforms)
                plan.propagate(x, *args, state_out=rec[r0 * engine.STATE_RECORD:r1 * engine.STATE_RECORD])
                out[r0:r1].copy_(x, non_blocking=True)
        for stream, _, _ in ln:
            stream.synchronize()
    return out, engine.decode_state(rec.numpy())


def dbp_batch(field, dt, length, alpha=0.0, beta_2=0.0, beta_3=0.0, gamma=0.0, phi_max=0.01, h=None, **kw):
    """Batched digital back-propagation (devices.py:1280-1283: FIBER with negated parameters)."""
    return fiber_batch(field, dt, length, -alpha, -beta_2, -beta_3, -gamma, phi_max, h, **kw)


def FIBER(input, length, alpha=0.0, beta_2=0.0, beta_3=0.0, gamma=0.0, phi_max=0.01, h=None,
          show_progress=False, return_steps=False, *, precision=None, device=None):
    """Optical fibre by the split-step Fourier method -- reference signature, devices.py:1038-1048.

    Units: km, dB/km, ps^2/km, ps^3/km, 1/(W km), rad.  Returns an ``optical_signal`` (noise NULL),
    or ``(z, A_z)`` with ``return_steps=True`` exactly like devices.py:1201-1202.
    ``show_progress``: the reference refreshes a tqdm bar after every step (devices.py:1164-1170, 1188-1191); here the
    propagation is one kernel, so the bar polls the device-side controller record (position reached, steps taken) while the
    kernel runs.
    """
    tic()
    try:
        if _kind(input) != "optical":
            raise TypeError("`input` must be of type 'optical_signal'.")
        return _fiber(input, length, alpha, beta_2, beta_3, gamma, phi_max, h, show_progress, return_steps, precision, device)
    except Exception:
        toc()                                               # never leave a tic behind (the reference leaks its own on errors)
        raise


def _progress_bar(show):
    if not show:
        return None
    try:
        from tqdm.auto import tqdm
        return tqdm(total=100, desc="Propagando", bar_format="{l_bar}{bar}|[{elapsed}{postfix}]", postfix={"FFTs": 0})
    except Exception:
        return None


def _bar_set(bar, z, length, steps):
    if bar is None:
        return
    bar.set_postfix(FFTs=2 * int(steps))
    bar.n = round(min(100.0, 100.0 * float(z) / float(length)) if length else 100.0, 1)
    bar.refresh()


def _fiber(input, length, alpha, beta_2, beta_3, gamma, phi_max, h, show_progress, return_steps, precision, device):
    torch = engine._torch()
    tdtype, ndtype = _complex_dtype(precision)
    dev = engine.require_cuda(device)
    dt = _gv_of(input).dt
    a = np.asarray(input.to_numpy())                      # signal + noise (typing.py:1596)
    n_pol = 1 if a.ndim == 1 else a.shape[0]
    n = a.shape[-1]
    _check_length(n)
    if n > MAX_TWO_PASS:                                   # beyond one two-pass transform: N = N0 x N_l stages (longwave.py)
        from . import longwave
        if return_steps:
            z, traj = longwave.fiber_long(a, dt, length, alpha, beta_2, beta_3, gamma, phi_max, h, return_steps=True,
                                          precision=precision or DEFAULT_PRECISION, device=dev)
            toc()
            return z, traj
        out, info = longwave.fiber_long(a, dt, length, alpha, beta_2, beta_3, gamma, phi_max, h,
                                        precision=precision or DEFAULT_PRECISION, device=dev)
        output = type(input)(out.astype(ndtype, copy=False))
        output.execution_time = toc()
        output.ssfm_info = info
        return output
    x = _to_device(a.reshape(1, n_pol, n), tdtype, dev)    # cast to the compute dtype on the device
    plan = engine.get_plan(n, n_pol, 1, tdtype, dev)
    _set_schedule(plan, True, None)
    args = (dt, length, alpha, beta_2, beta_3, gamma, phi_max, h)
    bar = _progress_bar(show_progress)

    if return_steps:
        # one step per call, resuming the device-side controller; snapshots stay on the device
        z_list, snaps = [0.0], [x.clone()]
        info = plan.propagate(x, *args, max_steps=1, resume=False)
        while True:
            if int(info.steps[0]) == len(z_list) - 1:      # nothing happened (length <= 0)
                break
            z_list.append(np.float32(info.z[0]) if ndtype is np.complex64 else np.float64(info.z[0]))
            snaps.append(x.clone())
            _bar_set(bar, info.z[0], length, info.steps[0])
            if info.done[0]:
                break
            info = plan.propagate(x, *args, max_steps=1, resume=True)
        if bar is not None:
            bar.close()
        toc()                                               # (the reference leaks its tic here)
        traj = torch.stack(snaps).cpu().numpy().reshape((len(snaps),) + a.shape)
        return np.array(z_list), traj

    if bar is not None and plan.get_option("persistent"):
        # asynchronous launch + polling of the controller record the kernel rewrites after every step (ssfm_peek_state)
        import time
        rec = torch.empty(engine.STATE_RECORD, dtype=torch.uint8, pin_memory=True)
        stream = torch.cuda.current_stream(dev)
        plan.propagate(x, *args, state_out=rec)
        while not stream.query():
            steps_now, z_now, _ = plan.peek(0)
            _bar_set(bar, z_now, length, steps_now)
            time.sleep(0.02)
        stream.synchronize()
        info = engine.decode_state(rec.numpy())
    else:
        info = plan.propagate(x, *args)
    if bar is not None:
        _bar_set(bar, length, length, info.steps[0]); bar.close()
    out = x.cpu().numpy().reshape(a.shape)
    output = type(input)(out)
    output.execution_time = toc()
    output.ssfm_info = info                                 # additive: steps / z / h of this call
    return output


def DBP(input, length, alpha=0.0, beta_2=0.0, beta_3=0.0, gamma=0.0, phi_max=0.01, h=None,
        show_progress=False, return_steps=False, **kw):
    """Digital back-propagation -- reference devices.py:1209-1283 (negated fibre parameters)."""
    return FIBER(input, length=length, alpha=-alpha, beta_2=-beta_2, beta_3=-beta_3, gamma=-gamma,
                 phi_max=phi_max, h=h, show_progress=show_progress, return_steps=return_steps, **kw)


# ---- DM: dispersive medium = one FFT-domain transfer function ------------------------------------------
def transfer_batch(x, h, *, device=None):
    """``ifft(fft(x) * H)`` along the last axis of ``x[..., N]`` (NumPy array or CUDA tensor, complex128 arithmetic) with
    ``H[N]`` in numpy bin order -- the operation of DM and of the apply step of FBG.  Any N in [2, 2^21] (powers of two up to
    2^22); lengths that are not powers of two run chirp-z transforms."""
    torch = engine._torch()
    as_tensor = torch.is_tensor(x)
    dev = engine.require_cuda(x.device if as_tensor and x.is_cuda else device)
    xt = (x if as_tensor else torch.from_numpy(np.ascontiguousarray(x))).to(dev).to(torch.complex128)
    ht = (h if torch.is_tensor(h) else torch.from_numpy(np.ascontiguousarray(h))).to(dev).to(torch.complex128).contiguous()
    n = xt.shape[-1]
    y = xt.reshape(-1, n).contiguous()
    if y.data_ptr() == xt.data_ptr() and as_tensor and x.is_cuda and x.dtype == torch.complex128:
        y = y.clone()
    plan = engine.get_plan(n, 1, y.shape[0], torch.complex128, dev, lane=99)
    plan.apply_transfer(y, ht)
    y = y.reshape(xt.shape)
    return y if as_tensor else y.cpu().numpy()


def DM(input, D, retH=False, *, device=None):
    """Dispersive medium -- reference devices.py:941-1035: ``H = exp(j w^2 D/2)``, ``output = ifft(fft(input) * H)``
    for signal and noise (complex128, as NumPy computes it).  ``D`` in ps^2."""
    tic()
    if _kind(input) != "optical":
        toc()
        raise TypeError("`input` must be of type 'optical_signal'.")
    dev = engine.require_cuda(device)
    D = D * 1e-12 ** 2                                        # devices.py:1023 (the reference rebinds its argument the same way)
    n = input.size
    try:
        _check_length(n)
        if n > MAX_TWO_PASS:
            raise ValueError("opticomlib_b200.DM supports waveforms of up to 2^22 samples; got %d" % n)
    except Exception:
        toc()
        raise
    w = np.fft.fftfreq(n, _gv_of(input).dt) * 2 * np.pi      # input.w(), typing.py:1641
    H = np.exp(1j * w ** 2 * D / 2)
    sig = np.asarray(input.signal, dtype=np.complex128)
    has_noise = not _is_null(input.noise)
    rows = [sig.reshape(-1, n)]
    if has_noise:
        rows.append(np.asarray(input.noise, dtype=np.complex128).reshape(-1, n))
    try:
        y = transfer_batch(np.concatenate(rows), H, device=dev)
    except Exception:
        toc()
        raise
    k = rows[0].shape[0]
    output = type(input)(y[:k].reshape(sig.shape), y[k:].reshape(sig.shape)) if has_noise else type(input)(y[:k].reshape(sig.shape))
    if retH:
        return output, np.fft.fftshift(H)                     # (the reference leaks its tic on this branch)
    output.execution_time = toc()
    return output


# ---- LPF / BPF ------------------------------------------------------------------------------------
def _bessel_sos(n, wn, fs):
    """Filter design on the host exactly as the reference asks SciPy for it (devices.py:814, 1363)."""
    from scipy import signal as sg

    return sg.bessel(N=n, Wn=wn, btype="low", fs=fs, output="sos", norm="mag")


def filtfilt_batch(x, sos, *, device=None):
    """Zero-phase filter rows of ``x[..., N]`` (NumPy or CUDA tensor) with cascaded biquads ``sos``."""
    torch = engine._torch()
    if torch.is_tensor(x):
        dev = engine.require_cuda(x.device if x.is_cuda else device)
        y = engine.filtfilt_sos(x.to(device=dev, dtype=torch.complex128).contiguous(), sos)
        return y
    dev = engine.require_cuda(device)
    a = np.asarray(x)
    y = engine.filtfilt_sos(_to_device(a, torch.complex128, dev), sos).cpu().numpy()
    return y if np.iscomplexobj(a) else y.real


def _filter_pair(sig, noi, sos, dev):
    """Filter signal and noise separately (as devices.py:820-823 / 1365-1368 do) in ONE launch:
    the rows are stacked; for real inputs signal and noise travel as re/im of one complex row."""
    torch = engine._torch()
    sig = np.asarray(sig)
    has_noise = not _is_null(noi)
    if has_noise:
        noi = np.asarray(noi)
    if not np.iscomplexobj(sig) and (not has_noise or not np.iscomplexobj(noi)):
        packed = sig.astype(np.float64) + 1j * (noi.astype(np.float64) if has_noise else 0.0)
        y = engine.filtfilt_sos(_to_device(packed, torch.complex128, dev), sos).cpu().numpy()
        return y.real.copy(), (y.imag.copy() if has_noise else NULL)
    rows = np.stack([sig, noi]) if has_noise else sig[np.newaxis]
    y = engine.filtfilt_sos(_to_device(rows, torch.complex128, dev), sos).cpu().numpy()
    return y[0], (y[1] if has_noise else NULL)


def LPF(input, BW, n=4, fs=None, retH=False, *, device=None):
    """Zero-phase Bessel low-pass of an electrical signal -- reference devices.py:1286-1375."""
    tic()
    if _kind(input) is None:
        input = electrical_signal(input)
    if input.ndim != 1:
        toc()
        raise ValueError("`input` must be a 1D-array.")
    if not fs:
        fs = _gv_of(input).fs
    dev = engine.require_cuda(device)
    sos = _bessel_sos(n, BW, fs)
    output = input[:]
    try:
        s, nz = _filter_pair(input.signal, input.noise, sos, dev)
    except Exception:
        toc()
        raise
    output.signal = np.asarray(s).real
    if not _is_null(input.noise):
        output.noise = np.asarray(nz).real
    if retH:
        from scipy import signal as sg
        _, H = sg.sosfreqz(sos, worN=input.size, fs=fs, whole=True)
        toc()
        return output, np.fft.fftshift(H)
    output.execution_time = toc()
    return output


def BPF(input, BW, n=4, *, device=None):
    """Zero-phase Bessel band-pass of an optical envelope -- reference devices.py:788-826."""
    tic()
    if _kind(input) != "optical":
        toc()
        raise TypeError("`input` must be of type (optical_signal).")
    dev = engine.require_cuda(device)
    sos = _bessel_sos(n, BW / 2, _gv_of(input).fs)
    output = input[:]
    try:
        s, nz = _filter_pair(input.signal, input.noise, sos, dev)
    except Exception:
        toc()
        raise
    output.signal = s
    if not _is_null(output.noise):
        output.noise = nz
    output.execution_time = toc()
    return output


# ---- N2 / N3: EDFA noise realisations and the photodetector chain on the device -------------
PLANCK = 6.62607015e-34
K_BOLTZMANN = 1.380649e-23
Q_ELECTRON = 1.602176634e-19


def edfa_batch(field, rows, G, NF, *, n_pol_out=2, seed=0, fs=None, f0=None, device=None):
    """``rows`` independent EDFA outputs ``sqrt(G) E + ASE`` generated ON the device (reference EDFA, devices.py:921-936,
    gain and noise; follow with ``filtfilt_batch`` for its optional BPF).  ``field``: one waveform ``[N]`` / ``[P, N]`` (NumPy
    or CUDA tensor) amplified into ``rows`` noise realisations -- the Monte-Carlo batch of BASELINE config #3 without moving a
    single noise sample over PCIe -- or a batch ``[rows, (P,) N]``.  ``G``, ``NF`` in dB; ASE power
    ``idb(NF) h f0 (idb(G) - 1) fs`` (devices.py:930) split over four real components; the generator is the extension's
    Philox (a function of ``seed`` and the element index), not NumPy's global stream.  ``n_pol_out`` = 2 like the reference
    (a polarisation the input does not have carries ASE only), 1 keeps a one-polarisation batch ``[rows, N]``.
    Returns a CUDA complex128 tensor."""
    torch = engine._torch()
    dev = engine.require_cuda(field.device if torch.is_tensor(field) and field.is_cuda else device)
    x = (field if torch.is_tensor(field) else torch.from_numpy(np.ascontiguousarray(field))).to(dev).to(torch.complex128).contiguous()
    fs = gv.fs if fs is None else fs
    f0 = gv.f0 if f0 is None else f0
    p_ase = 10 ** (NF / 10) * PLANCK * f0 * (10 ** (G / 10) - 1) * fs
    return engine.edfa(x, rows, G, p_ase, seed, out_pol=n_pol_out)


def edfa_fiber_batch(field, rows, G, NF, dt, length, alpha=0.0, beta_2=0.0, beta_3=0.0, gamma=0.0, phi_max=0.01, h=None, *,
                     seed=0, precision=None, out=None, fs=None, f0=None, device=None, first_row=0):
    """Monte-Carlo batch end to end WITHOUT moving the noise realisations over PCIe: ``rows`` EDFA outputs of ONE waveform
    ``field[N]`` (gain ``G`` dB, noise figure ``NF`` dB, x polarisation kept: BASELINE config #3) are generated on the device
    chunk by chunk (``ssfm_edfa``, Philox), propagated through FIBER and copied to the pinned host tensor ``out[rows, N]``;
    generation, propagation and the D2H copy of different chunks overlap on ``HOST_LANES`` streams, one enqueueing host
    thread (the pipeline of ``fiber_batch`` with the H2D copy replaced by the generator).  Row b is the same whatever the
    chunking: its noise depends on ``seed`` and ``first_row + b`` only.  Returns ``(out, StepInfo)``."""
    torch = engine._torch()
    tdtype, ndtype = _complex_dtype(precision)
    dev = engine.require_cuda(device)
    x = (field if torch.is_tensor(field) else torch.from_numpy(np.ascontiguousarray(field))).to(dev).to(torch.complex128).contiguous()
    if x.ndim != 1:
        raise ValueError("field must be one waveform [N]")
    N, B = x.shape[0], int(rows)
    _check_length(N)
    fs = gv.fs if fs is None else fs
    f0 = gv.f0 if f0 is None else f0
    p_ase = 10 ** (NF / 10) * PLANCK * f0 * (10 ** (G / 10) - 1) * fs
    if out is None:
        out = torch.empty((B, N), dtype=tdtype, pin_memory=True)
    elif tuple(out.shape) != (B, N) or out.dtype != tdtype or out.is_cuda:
        raise ValueError("out must be a host tensor of shape (%d, %d) and dtype %s" % (B, N, tdtype))
    crows = host_chunk_rows(B, 1, N, tdtype)
    chunks = [(r0, min(B, r0 + crows)) for r0 in range(0, B, crows)]
    lanes = min(HOST_LANES, len(chunks))
    args = (dt, length, alpha, beta_2, beta_3, gamma, phi_max, h)
    rec = torch.empty(B * engine.STATE_RECORD, dtype=torch.uint8, pin_memory=True)
    if (tdtype == torch.complex128 and HOST_SINGLE_LAUNCH and len(chunks) > 1 and N % 4096 == 0 and B * N * 16 <= (64 << 30)
            and B * N >= (1 << 26)):
        # (smaller batches -- one GPU's share of 4096 x 2^16 at 8 GPUs -- keep one launch per chunk: a chunk of the single launch
        # is only complete when the slowest team that drew one of its rows is done, which delays the first copies out by a
        # few ms; measured at 8 GPUs, where the copies are what the run waits for: 58 against 51 ms)
        # the whole batch is generated on the device, ONE persistent launch propagates it, and a second stream copies every
        # ~32 MiB chunk to the host as soon as the kernel has counted its tiles as final (ssfm_propagate_streamed with all
        # rows "arrived" from the start)
        import ctypes
        lib = engine._lib.load()
        srows, sch = single_launch_chunks(B, N * 16)
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            xdev = torch.empty((B, N), dtype=torch.complex128, device=dev)
            engine.edfa(x, B, G, p_ase, seed, out_pol=1, out=xdev, first_row=first_row)
            flags = torch.zeros(1 + len(sch), dtype=torch.int32, device=dev)
            flags[0] = B
            plan = engine.get_plan(N, 1, B, tdtype, dev, lane=0)
            _set_schedule(plan, True, True)
            reset = torch.cuda.Event(); reset.record(main)
            if plan.propagate_streamed(xdev, flags.data_ptr(), flags.data_ptr() + 4, srows, *args, state_out=rec):
                d2h = torch.cuda.Stream(device=dev)
                d2h.wait_event(reset)
                with torch.cuda.stream(d2h):
                    for ci, (r0, r1) in enumerate(sch):
                        engine._lib.check(lib.ssfm_stream_wait_geq_u32(ctypes.c_void_p(d2h.cuda_stream), ctypes.c_void_p(flags.data_ptr() + 4 + 4 * ci),
                                                                       (r1 - r0) * (N // 4096)))
                        out[r0:r1].copy_(xdev[r0:r1], non_blocking=True)
                d2h.synchronize(); main.synchronize()
                return out, engine.decode_state(rec.numpy())
    with torch.cuda.device(dev):
        ln = [(torch.cuda.Stream(device=dev), torch.empty((crows, N), dtype=torch.complex128, device=dev),
               torch.empty((crows, N), dtype=tdtype, device=dev) if tdtype != torch.complex128 else None) for _ in range(lanes)]
        torch.cuda.current_stream(dev).synchronize()
        for ci, (r0, r1) in enumerate(chunks):
            stream, gen, cast = ln[ci % lanes]
            m = r1 - r0
            with torch.cuda.stream(stream):
                engine.edfa(x, m, G, p_ase, seed, out_pol=1, out=gen[:m], first_row=first_row + r0)
                w = gen[:m]
                if cast is not None:
                    cast[:m].copy_(w)
                    w = cast[:m]
                plan = engine.get_plan(N, 1, m, tdtype, dev, lane=ci % lanes)
                _set_schedule(plan, True, True)
                plan.propagate(w, *args, state_out=rec[r0 * engine.STATE_RECORD:r1 * engine.STATE_RECORD])
                out[r0:r1].copy_(w, non_blocking=True)
        for stream, _, _ in ln:
            stream.synchronize()
    return out, engine.decode_state(rec.numpy())


def pd_lpf_batch(field, sos, *, noise=None, extra_noise=None, responsivity=1.0, r_load=50.0, i_dark=0.0,
                 sample_offset=0, sample_stride=1, device=None):
    """Square-law detection, zero-phase low-pass and sampling of a batch ``field[B, (P,) N]`` in one pass over the field
    (the chain PD -> LPF -> SAMPLER of devices.py:1514-1552 and 1871-1891).  Returns ``(signal, noise)`` float64 ``[B, m]``
    (CUDA tensors for CUDA input, NumPy arrays otherwise; ``noise`` is None when no noise input is given)."""
    torch = engine._torch()
    as_tensor = torch.is_tensor(field)
    dev = engine.require_cuda(field.device if as_tensor and field.is_cuda else device)

    def dev_c(a, dt):
        if a is None:
            return None
        t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
        return t.to(dev).to(dt).contiguous()

    s, nz = engine.pd_lpf(dev_c(field, torch.complex128), sos, dev_c(noise, torch.complex128), dev_c(extra_noise, torch.float64),
                          responsivity, r_load, i_dark, sample_offset, sample_stride)
    if as_tensor:
        return s, nz
    return s.cpu().numpy(), (None if nz is None else nz.cpu().numpy())


def psd_batch(x, fs=None, nperseg=None, *, device=None):
    """``(f, psd)`` of every row of ``x[..., N]`` exactly as ``utils.get_psd`` returns them for one signal (utils.py:2048-2079:
    Welch, ``nperseg = min(2048, N)``, Hann, 50 % overlap, scaling='spectrum', two-sided, fftshift).  NumPy in, NumPy out; CUDA
    tensor in, CUDA tensor out (``f`` is always a NumPy array)."""
    torch = engine._torch()
    as_tensor = torch.is_tensor(x)
    dev = engine.require_cuda(x.device if as_tensor and x.is_cuda else device)
    t = (x if as_tensor else torch.from_numpy(np.ascontiguousarray(x))).to(dev).to(torch.complex128).contiguous()
    nper = min(2048, t.shape[-1]) if nperseg is None else int(nperseg)
    psd = engine.welch_psd(t, nper)
    f = np.fft.fftshift(np.fft.fftfreq(nper, 1.0 / (gv.fs if fs is None else fs)))
    return f, (psd if as_tensor else psd.cpu().numpy())


_PD_MODES = {            # include_noise -> (ase, thermal, shot)        devices.py:1530-1545
    "ase-only": (1, 0, 0), "thermal-only": (0, 1, 0), "shot-only": (0, 0, 1), "ase-shot": (1, 0, 1),
    "ase-thermal": (1, 1, 0), "thermal-shot": (0, 1, 1), "all": (1, 1, 1), "none": (0, 0, 0),
}


def PD(input, BW, r=1.0, T=300.0, R_load=50.0, include_noise="all", i_dark=10e-9, Fn=0, *, device=None):
    """P-I-N photodetector -- reference signature and semantics, devices.py:1377-1556: square law, thermal / shot / ASE beat
    noise and dark current kept as the ``noise`` of the returned ``electrical_signal``, then ``LPF(output, BW)``.
    The thermal and shot samples are drawn on the host with ``np.random.normal`` in the reference's order (so a seeded script
    sees the same values); square law, noise assembly and the zero-phase filter run in one device call."""
    tic()
    try:
        if _kind(input) != "optical":
            raise TypeError("`input` must be of type 'optical_signal'.")
        for name, val in (("r", r), ("T", T), ("R_load", R_load)):
            if isinstance(val, bool) or not isinstance(val, (int, float, np.integer, np.floating)):
                raise TypeError("`%s` must be a scalar value." % name)
        if r <= 0 or r > 1:
            raise ValueError("`r` must be in the range (0,1]")
        if T < 0:
            raise ValueError("`T` must be a positive value.")
        if R_load < 0:
            raise ValueError("`R_load` must be a positive value.")
        if not isinstance(include_noise, str):
            raise TypeError("`include_noise` must be a string.")
        mode = include_noise.lower()
        torch = engine._torch()
        dev = engine.require_cuda(device)
        g = _gv_of(input)
        sig = np.asarray(input.signal, dtype=np.complex128)
        n = sig.shape[-1]
        has_ase_in = not _is_null(input.noise)
        noi = np.asarray(input.noise, dtype=np.complex128) if has_ase_in else None
        thermal = shot = None
        if "thermal" in mode or "all" in mode:                                # same draws, same order as devices.py:1521-1527
            S_T = 4 * K_BOLTZMANN * T * g.fs / 2 * 10 ** (Fn / 10) / R_load
            thermal = np.random.normal(0, S_T ** 0.5, n)
        if "shot" in mode or "all" in mode:
            tot = sig if noi is None else sig + noi
            i_mean = (r * (tot * tot.conj()).real).sum(axis=0).mean() if tot.ndim == 2 else (r * (tot * tot.conj()).real).mean()
            S_N = 2 * Q_ELECTRON * (i_mean + i_dark) * g.fs / 2
            shot = np.random.normal(0, S_N ** 0.5, n)
        if mode not in _PD_MODES:
            raise ValueError("The argument `include_noise` must be one of the following: 'ase-only','thermal-only','shot-only',"
                             "'ase-thermal','ase-shot','thermal-shot','all', 'none'.")
        use_ase, use_t, use_n = _PD_MODES[mode]
        sos = _bessel_sos(4, BW, g.fs)
        f = torch.from_numpy(np.ascontiguousarray(sig.reshape(1, -1, n))).to(dev)
        nz = torch.from_numpy(np.ascontiguousarray(noi.reshape(1, -1, n))).to(dev) if (use_ase and has_ase_in) else None
        extra = None
        if mode != "none":
            e = np.zeros(n)
            if use_n:
                e = e + shot
            if use_t:
                e = e + thermal
            extra = torch.from_numpy(e.reshape(1, n)).to(dev)
        s_f, n_f = engine.pd_lpf(f, sos, nz, extra, r, R_load, i_dark if mode != "none" else 0.0)
        klass = electrical_signal
        mod = sys.modules.get(type(input).__module__)
        if mod is not None and hasattr(mod, "electrical_signal"):
            klass = mod.electrical_signal                                       # the reference's own class for its objects
        output = klass(s_f[0].cpu().numpy()) if n_f is None else klass(s_f[0].cpu().numpy(), n_f[0].cpu().numpy())
    except Exception:
        toc()
        raise
    output.execution_time = toc()
    return output


# ---- drop-in installation -----------------------------------------------------------------------
_SAVED: dict = {}


def install(precision: str | None = None):
    """Rebind ``opticomlib.devices.{FIBER,DBP,LPF,BPF}`` (and the import-time copies
    ``opticomlib.ook.LPF`` / ``opticomlib.ppm.LPF``, reference ook.py:16, ppm.py:21) to this path.
    In-module callers (DAC->LPF, PD->LPF, MZM->BPF, EDFA->BPF) resolve the names through the module
    globals at call time, so they are redirected too."""
    global DEFAULT_PRECISION
    import importlib

    dv = importlib.import_module("opticomlib.devices")
    if precision is not None:
        _complex_dtype(precision)
        DEFAULT_PRECISION = precision
    for name, fn in (("FIBER", FIBER), ("DBP", DBP), ("LPF", LPF), ("BPF", BPF), ("DM", DM), ("PD", PD)):
        _SAVED.setdefault(("opticomlib.devices", name), getattr(dv, name))
        setattr(dv, name, fn)
    for modname in ("opticomlib.ook", "opticomlib.ppm"):
        mod = sys.modules.get(modname)
        if mod is not None and hasattr(mod, "LPF"):
            _SAVED.setdefault((modname, "LPF"), mod.LPF)
            mod.LPF = LPF


def uninstall():
    for (modname, name), fn in list(_SAVED.items()):
        mod = sys.modules.get(modname)
        if mod is not None:
            setattr(mod, name, fn)
    _SAVED.clear()
