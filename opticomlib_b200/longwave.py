"""Long waveforms: one FIBER / DBP propagation whose transform is split as N = N0 x N_l.

Two uses, one code path (BASELINE config #5):

* a single GPU and N > 2^22 samples (up to 2^30): the outer N0-point stage and the inner N_l-point stage
  both run on the same device and no data moves between them;
* one waveform spread over the G GPUs of a process group: rank g keeps columns [g N_l/G, (g+1) N_l/G) of the
  N0 x N_l sample matrix in the time domain and rows [g N0/G, (g+1) N0/G) of the (transposed-order) spectrum;
  the two layouts are exchanged with ``torch.distributed.all_to_all_single`` (NCCL over NVLink on the GPU
  box, gloo in the CPU tests of the index logic) -- twice per split step -- plus one scalar all-reduce (MAX)
  per step in adaptive mode.  Every arithmetic operation stays in the CUDA kernels behind the C-ABI
  (``ssfm_long_*`` in include/ssfm_b200.h); this module only sequences stages and moves bytes.

The statements reproduced are the same as for short waveforms (opticomlib/devices.py:1155-1196); the
reference itself has no multi-device path.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import _lib, engine

INNER_LOG2 = 18          # preferred inner transform length N_l = 2^18 (512 x 512 two-pass kernels)


def _torch():
    import torch

    return torch


def split_sizes(n_global: int, n_ranks: int = 1, n_outer: int | None = None):
    """(N0, N_l) for a waveform of ``n_global`` samples on ``n_ranks`` devices."""
    if n_global & (n_global - 1) or n_global < (1 << 12):
        raise ValueError("long waveforms need a power-of-two length >= 2^12, got %d" % n_global)
    m = n_global.bit_length() - 1
    if n_outer is None:
        lo = max(4, m - 22, int(math.log2(max(n_ranks, 1))))
        hi = min(11, m - 8)
        n_outer = 1 << min(max(m - INNER_LOG2, lo), hi)
    n_inner = n_global // n_outer
    if n_outer % n_ranks or n_inner % n_ranks or n_inner // n_ranks < 32:
        raise ValueError("cannot split %d samples as %d x %d over %d ranks" % (n_global, n_outer, n_inner, n_ranks))
    return n_outer, n_inner


def fixed_step_count(length, h, real) -> int:
    """Number of steps of the fixed-h controller (devices.py:1159-1162, 1173, 1195-1196) in the compute real type."""
    R = np.float32 if real in (np.float32, "fp32") else np.float64
    L, hk, z = R(length), R(h), R(0)
    hk = L if L < hk else hk
    steps = 0
    while z < L:
        z = R(z + hk)
        steps += 1
        rem = R(L - z)
        hk = rem if rem < hk else hk
    return steps


def local_columns(x_full, n_outer: int, n_ranks: int, rank: int):
    """This rank's time-domain share of a full waveform x[N]: columns of the N0 x N_l matrix, as [N0, N_l/G]."""
    n = x_full.shape[-1]
    n_inner = n // n_outer
    w = n_inner // n_ranks
    return x_full.reshape(n_outer, n_inner)[:, rank * w:(rank + 1) * w]


class CudaStages:
    """The stage kernels of one rank behind the C-ABI (``ssfm_long_*``): outer tables, stash, controller, inner plan."""

    def __init__(self, n_global, n_outer, ranks, rank, cdtype, device):
        torch = _torch()
        self.lib = _lib.load()
        self.device = device
        code = _lib.SSFM_C64 if cdtype == torch.complex64 else _lib.SSFM_C128
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(self.lib.ssfm_long_plan_create(ctypes.byref(h), n_global, n_outer, ranks, rank, code, device.index))
        self.handle = h

    def _stream(self):
        return ctypes.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def begin(self, field, prm):
        _lib.check(self.lib.ssfm_long_begin(self.handle, field.data_ptr(), ctypes.byref(prm), self._stream()))

    def pmax(self, value=None):
        v = ctypes.c_double(0.0 if value is None else value)
        _lib.check(self.lib.ssfm_long_pmax(self.handle, ctypes.byref(v), 0 if value is None else 1, self._stream()))
        return v.value

    def ctrl(self, init):
        _lib.check(self.lib.ssfm_long_ctrl(self.handle, 1 if init else 0, self._stream()))

    def outer(self, field, stage):
        _lib.check(self.lib.ssfm_long_outer(self.handle, field.data_ptr(), stage, self._stream()))

    def inner(self, rows):
        _lib.check(self.lib.ssfm_long_inner(self.handle, rows.data_ptr(), self._stream()))

    # ---- exchange fused into the kernels (peer memory over NVLink, CUDA IPC) --------------------------
    def p2p_export(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        _lib.check(self.lib.ssfm_long_p2p_export(self.handle, buf))
        return buf.raw

    def p2p_import(self, handles: bytes):
        _lib.check(self.lib.ssfm_long_p2p_import(self.handle, ctypes.c_char_p(handles)))

    def p2p_copy(self, field, to_internal):
        _lib.check(self.lib.ssfm_long_p2p_copy(self.handle, field.data_ptr(), 1 if to_internal else 0, self._stream()))

    def xbar(self):
        _lib.check(self.lib.ssfm_long_xbar(self.handle, self._stream()))

    def sync(self):
        _torch().cuda.current_stream(self.device).synchronize()

    def state(self, want_log=False):
        steps = np.empty(1, np.int32); z = np.empty(1, np.float64); hn = np.empty(1, np.float64); done = np.empty(1, np.int32)
        _lib.check(self.lib.ssfm_get_state(self.handle, steps.ctypes.data, z.ctypes.data, hn.ctypes.data, done.ctypes.data))
        log = None
        if want_log:
            cap = int(max(1, steps.max()))
            log = np.zeros((1, cap), np.float64)
            _lib.check(self.lib.ssfm_get_step_log(self.handle, log.ctypes.data, cap))
        return engine.StepInfo(steps, z, hn, done.astype(bool), log)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ssfm_plan_destroy(self.handle)
            self.handle = None


class LongPlan:
    """Sequencer of one rank: stages (``CudaStages``; the CPU tests inject a NumPy model of the same stages to check
    the sequencing and the exchange under gloo) + the two layout exchanges + the scalar max all-reduce."""

    def __init__(self, n_global, complex_dtype, device=None, group=None, n_outer=None, stages=None, fused_exchange=True):
        torch = _torch()
        self.group = group
        self.ranks, self.rank = 1, 0
        if group is not None:
            import torch.distributed as dist
            self.ranks, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n = int(n_global)
        self.n_outer, self.n_inner = split_sizes(self.n, self.ranks, n_outer)
        self.cols = self.n_inner // self.ranks            # columns of the sample matrix held in the time domain
        self.rows = self.n_outer // self.ranks            # rows of the spectrum held in the frequency domain
        self.cdtype = torch.complex64 if complex_dtype in (torch.complex64, np.complex64, "fp32") else torch.complex128
        self.real = np.float32 if self.cdtype == torch.complex64 else np.float64
        if stages is None:
            self.device = engine.require_cuda(device)
            stages = CudaStages(self.n, self.n_outer, self.ranks, self.rank, self.cdtype, self.device)
        else:
            self.device = torch.device("cpu")
        self.stages = stages
        self._buf = None                                   # exchange buffers (only with more than one rank)
        # Several GPUs: let the kernels store straight into the peers' buffers (CUDA IPC over NVLink) instead of an NCCL
        # all-to-all plus a re-layout copy.  Needs all ranks on one node; falls back to the collective otherwise.
        self.fused = False
        if self.ranks > 1 and fused_exchange and isinstance(stages, CudaStages) and self.ranks <= 8:
            import torch.distributed as dist
            try:
                mine = stages.p2p_export()
                ok = 1
            except Exception:
                mine, ok = b"\0" * 64, 0
            gathered = [None] * self.ranks
            dist.all_gather_object(gathered, (ok, mine), group=group)
            if all(g[0] for g in gathered):
                try:
                    stages.p2p_import(b"".join(g[1] for g in gathered))
                    ok = 1
                except Exception:
                    ok = 0
                flags = [None] * self.ranks
                dist.all_gather_object(flags, ok, group=group)
                self.fused = all(flags)
            if not self.fused and self.rank == 0:
                import warnings
                warnings.warn("opticomlib_b200.longwave: peer-memory exchange could not be set up on every rank (CUDA IPC); "
                              "falling back to the NCCL all-to-all", RuntimeWarning)

    def close(self):
        if getattr(self, "stages", None) is not None:
            self.stages.close()
            self.stages = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the exchange between the two layouts --------------------------------------------------
    def _buffers(self, like):
        torch = _torch()
        if self._buf is None:
            self._buf = (torch.empty_like(like), torch.empty((self.rows, self.n_inner), dtype=like.dtype, device=like.device))
        return self._buf

    def _to_rows(self, field):
        """[N0][N_l/G] (time layout, outer transform done) -> [N0/G][N_l] (this rank's rows)."""
        if self.ranks == 1:
            return field
        if self.fused:                                     # the outer kernel already stored into the owners' rows buffers
            self.stages.xbar()
            return field
        import torch.distributed as dist
        torch = _torch()
        recv, rows = self._buffers(field)
        dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(field), group=self.group)
        # recv[s] = block of my rows that rank s held: [G][N0/G][N_l/G] -> [N0/G][G][N_l/G]
        rows.view(self.rows, self.ranks, self.cols).copy_(recv.view(self.ranks, self.rows, self.cols).permute(1, 0, 2))
        return rows

    def _to_columns(self, rows, field):
        if self.ranks == 1:
            return
        if self.fused:
            self.stages.xbar()
            return
        import torch.distributed as dist
        torch = _torch()
        send, _ = self._buffers(field)
        send.view(self.ranks, self.rows, self.cols).copy_(rows.view(self.rows, self.ranks, self.cols).permute(1, 0, 2))
        dist.all_to_all_single(torch.view_as_real(field), torch.view_as_real(send), group=self.group)

    # ---- one propagation ------------------------------------------------------------------------
    def propagate(self, field, dt, length, alpha=0.0, beta_2=0.0, beta_3=0.0, gamma=0.0, phi_max=0.01, h=None,
                  want_log=False, on_step=None) -> engine.StepInfo:
        """In place on ``field``: CUDA tensor [N0, N_l/G] (this rank's columns of the sample matrix).

        ``on_step(field, state)``: called after every split step with the field back in the time domain (the trajectory of
        ``return_steps=True``, devices.py:1184-1186, 1201-1202); it forces the open/close stages also for a fixed step."""
        cb = None if on_step is None else (lambda fields, st: on_step(fields[0], st))
        return propagate_together([self], [field], dt, length, alpha, beta_2, beta_3, gamma, phi_max, h, want_log, cb)

    def state(self, want_log=False) -> engine.StepInfo:
        return self.stages.state(want_log)


def propagate_together(plans, fields, dt, length, alpha=0.0, beta_2=0.0, beta_3=0.0, gamma=0.0, phi_max=0.01, h=None,
                       want_log=False, on_step=None) -> engine.StepInfo:
    """Propagate the rows ``fields[i]`` (one per plan: the POLARISATIONS of one waveform) in lock step with ONE step-size
    sequence: the reference takes the max of |A|^2 over both polarisations (devices.py:1156, 1194), so after every close stage
    the per-plan maxima are combined (and all-reduced over the ranks) before every plan's controller runs."""
    torch = _torch()
    p0 = plans[0]
    on_cuda = isinstance(p0.stages, CudaStages)
    for pl, field in zip(plans, fields):
        if field.dtype != pl.cdtype or field.is_cuda != on_cuda or not field.is_contiguous():
            raise ValueError("field must be a contiguous %s tensor of dtype %s" % ("CUDA" if on_cuda else "CPU", pl.cdtype))
        if tuple(field.shape) != (pl.n_outer, pl.cols):
            raise ValueError("field must have shape (%d, %d), got %s" % (pl.n_outer, pl.cols, tuple(field.shape)))
    prm = _lib.FiberParams(float(dt), float(length), float(alpha), float(beta_2), float(beta_3), float(gamma),
                           float(phi_max), math.nan if h is None else float(h))
    R = p0.real
    fixed = h is not None
    single = (not fixed) and ((R(beta_2) == 0 and R(beta_3) == 0) or R(gamma) == 0)
    pairs = list(zip(plans, fields))

    def combine_max():
        if p0.ranks == 1 and len(plans) == 1:
            return
        vals = [pl.stages.pmax() for pl in plans]
        val = float("inf") if any(math.isnan(v) for v in vals) else max(vals)    # NaN must win the max, as in numpy
        if p0.ranks > 1:
            import torch.distributed as dist
            t = torch.tensor([val], dtype=torch.float64, device=p0.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=p0.group)
            val = float(t.item())
        val = float("nan") if math.isinf(val) and val > 0 else val
        for pl in plans:
            pl.stages.pmax(val)

    ctx = torch.cuda.device(p0.device) if on_cuda else _Null()
    with ctx:
        for pl, f in pairs:
            if pl.fused:
                pl.stages.p2p_copy(f, True)                     # the time-domain field lives in the library's IPC buffer
            pl.stages.begin(f, prm)
        if not fixed and not single:
            combine_max()
        for pl, f in pairs:
            pl.stages.ctrl(True)
        if p0.stages.state().done[0]:
            return p0.stages.state(want_log)
        for pl, f in pairs:
            if pl.fused:
                pl.stages.xbar()                                # nobody stores into a peer before every peer has loaded its field
            pl.stages.outer(f, 0)
        n_fixed = fixed_step_count(length, h, R) if fixed else 0
        done_steps = 0
        while True:
            for pl, f in pairs:
                rows = pl._to_rows(f)
                pl.stages.inner(rows)
                pl._to_columns(rows, f)
            done_steps += 1
            if fixed and on_step is None:
                for pl, f in pairs:
                    pl.stages.outer(f, 1)                       # end of this step (+ start of the next one)
                if done_steps >= n_fixed:
                    break
            else:
                for pl, f in pairs:
                    pl.stages.outer(f, 2)
                if not fixed:
                    combine_max()
                for pl, f in pairs:
                    pl.stages.ctrl(False)
                st = p0.stages.state()
                if on_step is not None:
                    for pl, f in pairs:
                        if pl.fused:
                            pl.stages.p2p_copy(f, False)
                    on_step(fields, st)
                if st.done[0]:
                    break
                for pl, f in pairs:
                    pl.stages.outer(f, 0)
        for pl, f in pairs:
            if pl.fused:
                pl.stages.p2p_copy(f, False)
        for pl, f in pairs:
            pl.stages.sync()
    info = p0.stages.state(want_log)
    if not info.done[0]:
        raise RuntimeError("long-waveform propagation ended before z reached the fibre length (controller out of step)")
    return info


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_PLANS: dict = {}


def get_long_plan(n_global, complex_dtype, device=None, group=None, n_outer=None, fused_exchange=True, pol=0) -> LongPlan:
    torch = _torch()
    dev = engine.require_cuda(device)
    cd = torch.complex64 if complex_dtype in (torch.complex64, np.complex64, "fp32") else torch.complex128
    key = (int(n_global), cd, dev.index, id(group) if group is not None else None, n_outer, bool(fused_exchange), int(pol))
    pl = _PLANS.pop(key, None)
    if pl is None:
        while len(_PLANS) >= 4:                             # long plans own O(N) device memory; least recently used goes first.
            _PLANS.pop(next(iter(_PLANS)))                  # Dropped, not destroyed: LongPlan.__del__ closes it with the last reference
        pl = LongPlan(n_global, cd, dev, group, n_outer, fused_exchange=fused_exchange)
    _PLANS[key] = pl
    return pl


def clear_plans():
    while _PLANS:
        _PLANS.popitem()[1].close()


def fiber_long(field, dt, length, alpha=0.0, beta_2=0.0, beta_3=0.0, gamma=0.0, phi_max=0.01, h=None, *,
               precision="fp32", device=None, group=None, n_outer=None, want_log=False, gather=True, fused_exchange=True,
               return_steps=False):
    """Propagate ONE waveform ``field[N]`` or ``field[P, N]`` (P = 1 | 2 polarisations sharing one step-size sequence; NumPy
    array or tensor, the same on every rank of ``group``).

    With ``group=None`` the whole waveform lives on this process's GPU (any power-of-two N in [2^12, 2^30]); with a
    process group its columns are spread over the ranks and the result is gathered back on every rank
    (``gather=False`` returns this rank's [N0, N_l/G] share instead).  ``fused_exchange`` (default): the kernels store
    straight into the peers' buffers over NVLink (CUDA IPC); ``False``: NCCL all-to-all + re-layout copy.
    Returns ``(out, StepInfo)``; with ``return_steps=True`` (one rank only) ``(z, A_z)`` like devices.py:1201-1202: positions
    ``z[steps+1]`` (float64, ``z[0] = 0``) and the field after every step ``A_z[steps+1, N]`` (host array).
    """
    torch = _torch()
    dev = engine.require_cuda(device)
    tdtype = torch.complex64 if precision in ("fp32", "float32") else torch.complex128
    as_numpy = not torch.is_tensor(field)
    x = torch.from_numpy(np.ascontiguousarray(field)) if as_numpy else field
    if x.ndim not in (1, 2) or (x.ndim == 2 and x.shape[0] not in (1, 2)):
        raise ValueError("fiber_long takes one waveform of shape [N] or [P, N] with P = 1 or 2 polarisations")
    xs = [x] if x.ndim == 1 else [x[p] for p in range(x.shape[0])]
    plans = [get_long_plan(xs[0].shape[0], tdtype, dev, group, n_outer, fused_exchange, pol=p) for p in range(len(xs))]
    p0 = plans[0]
    mine = []
    for xp in xs:                                                  # one plan per polarisation, propagated in lock step
        m = local_columns(xp, p0.n_outer, p0.ranks, p0.rank).to(dev).to(tdtype).contiguous()
        mine.append(m.clone() if m.data_ptr() == xp.data_ptr() else m)

    def whole(parts):                                              # [P][N0, cols] -> tensor shaped like the input (this rank's share)
        flat = [m.reshape(-1) for m in parts]
        return flat[0] if x.ndim == 1 else torch.stack(flat)

    if return_steps:
        if p0.ranks > 1:
            raise NotImplementedError("return_steps is available for one rank only")
        real = np.float32 if tdtype == torch.complex64 else np.float64
        z_list, snaps = [0.0], [whole(mine).cpu()]                 # snapshots go to the host: a trajectory of a long waveform is large
        propagate_together(plans, mine, dt, length, alpha, beta_2, beta_3, gamma, phi_max, h,
                           on_step=lambda fs, st: (z_list.append(real(st.z[0])), snaps.append(whole(fs).cpu())))
        return np.array(z_list, dtype=np.float64), torch.stack(snaps).numpy()
    info = propagate_together(plans, mine, dt, length, alpha, beta_2, beta_3, gamma, phi_max, h, want_log=want_log)
    if not gather:
        return (mine[0] if x.ndim == 1 else torch.stack(mine)), info
    if p0.ranks > 1:
        import torch.distributed as dist
        full = []
        for m in mine:
            parts = [torch.empty_like(m) for _ in range(p0.ranks)]
            dist.all_gather(parts, m, group=group)
            full.append(torch.cat(parts, dim=1))
        out = whole(full)
    else:
        out = whole(mine)
    return (out.cpu().numpy() if as_numpy else out), info


def dbp_long(field, dt, length, alpha=0.0, beta_2=0.0, beta_3=0.0, gamma=0.0, phi_max=0.01, h=None, **kw):
    """devices.py:1280-1283 for long waveforms: FIBER with negated parameters."""
    return fiber_long(field, dt, length, -alpha, -beta_2, -beta_3, -gamma, phi_max, h, **kw)
