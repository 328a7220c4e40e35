#!/usr/bin/env python
"""bench.py -- SSFM sample*steps/s of the FIBER hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--precision fp64]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm on the host CPU cores

One "step" = one full FIBER propagation (adaptive split-step loop, devices.py:1155-1196) of the
workload's batch of synthetic PRBS-driven OOK waveforms.  Default workload: BASELINE config #3,
a Monte-Carlo batch of 4096 waveforms x 2^16 samples (EDFA ASE realisations), sharded by rows over
the ranks (strong scaling, no data-path collective; rows are independent).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ssfm_sample_steps_per_s"
UNIT = "sample*steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg1", "cfg2", "cfg3", "cfg5"])
    ap.add_argument("--log2n", type=int, default=26, help="cfg5: log2 of the waveform length")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="cfg5 on several GPUs: kernels store into peer memory over NVLink (fused) or NCCL all-to-all")
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--rows", type=int, default=0, help="override the number of waveforms (whole job)")
    ap.add_argument("--chunk", type=int, default=-1, help="waveforms propagated together (-1 = auto)")
    ap.add_argument("--schedule", default="persistent", choices=["persistent", "multilaunch"],
                    help="persistent = one kernel per propagation (default); multilaunch = two kernels per step")
    ap.add_argument("--teams", type=int, default=0, help="persistent schedule: cap on waveforms in flight (0 = auto)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary measurements")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the CPU baseline sample")
    return ap.parse_args()


def workload(name, rows_override=0):
    from opticomlib_b200 import workloads as wl
    base, dt, kw = wl.config_input("cfg1" if name == "cfg3" else name)
    c = wl.CONFIGS[name]
    rows = rows_override or c.get("rows", 1)
    return dict(name=name, base=base, dt=dt, fiber=kw, rows=rows, n=base.size,
                gain_db=c.get("gain_db"), nf_db=c.get("nf_db"), fs=c["R"] * c["sps"], sps=c["sps"], rate=c["R"])


DESCR = {
    "cfg1": "BASELINE config #1: OOK 10 Gb/s PRBS7, 2^16 samples, 50 km SSMF",
    "cfg2": "BASELINE config #2: single 2^20-sample OOK, 100 km SSMF, beta_3, phi_max control, 20 dBm",
    "cfg3": "BASELINE config #3: Monte-Carlo batch of 4096 waveforms x 2^16 samples (EDFA ASE realisations) through FIBER",
    "cfg5": "BASELINE config #5: single 2^26-sample long-haul waveform, 100 km spans at h = 1 km (one bench step = one span "
            "= 100 split steps), transform split N0 x N_l over the GPUs (all-to-all over NVLink)",
}


# ---------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# CPU legs: the reference algorithm on host cores (oracle port; bit-identical to the reference
# for float32 -- tests/test_oracle_vs_reference.py)
# ---------------------------------------------------------------------------------------------
def _cpu_row(args):
    row, dt, kw, real = args
    from oracle.ssfm_oracle import oracle_fiber
    t = time.perf_counter()
    with np.errstate(all="ignore"):
        o = oracle_fiber(row, dt, real=np.float32 if real == "fp32" else np.float64, **kw)
    return o["steps"] * row.shape[-1], time.perf_counter() - t


def _ref_row(args):
    """One row through the UNMODIFIED reference FIBER (installed under baseline/_ref; import shim for matplotlib / pympler only).
    return_steps=True is how the reference reports its step count; the per-step snapshot copies cost < 1 % of a step."""
    row, dt, kw, sps, rate = args
    from oracle.ref_shim import import_reference
    import_reference()
    from opticomlib import gv, optical_signal
    from opticomlib.devices import FIBER
    gv(sps=sps, R=rate)
    assert abs(gv.dt / dt - 1) < 1e-12
    t = time.perf_counter()
    with np.errstate(all="ignore"):
        z, _ = FIBER(optical_signal(row), return_steps=True, **kw)
    return (len(z) - 1) * row.shape[-1], time.perf_counter() - t


def reference_available():
    try:
        from oracle.ref_shim import reference_root
        return reference_root() is not None
    except Exception:
        return False


def cpu_sample(w, precision, budget_s, cores=None, use_reference=False):
    """Time the oracle (or, with use_reference, the unmodified reference) on a bounded sample of the workload's rows with a
    process pool."""
    from concurrent.futures import ProcessPoolExecutor
    from opticomlib_b200 import workloads as wl
    cores = cores or os.cpu_count() or 1
    if w["rows"] > 1:
        one = wl.ase_rows(w["base"], [0], w["fs"], w["gain_db"], w["nf_db"])[0]
    else:
        one = w["base"]
    fn = _ref_row if use_reference else _cpu_row
    mk = (lambda r: (r, w["dt"], w["fiber"], w["sps"], w["rate"])) if use_reference else (lambda r: (r, w["dt"], w["fiber"], precision))
    units, t1 = fn(mk(one))                                             # calibrate on one row, one core
    if w["rows"] == 1:
        return dict(value=units / t1, cores=1, rows=1, seconds=t1, one_core=units / t1)
    nrows = int(max(cores, min(w["rows"], cores * max(1, int(budget_s / max(t1, 1e-3))))))
    rows = wl.ase_rows(w["base"], range(nrows), w["fs"], w["gain_db"], w["nf_db"])
    t0 = time.perf_counter()
    with ProcessPoolExecutor(max_workers=cores) as ex:
        res = list(ex.map(fn, [mk(rows[i]) for i in range(nrows)]))
    wall = time.perf_counter() - t0
    return dict(value=sum(r[0] for r in res) / wall, cores=cores, rows=nrows, seconds=wall, one_core=units / t1)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(a.workload, a.rows)
    cores = os.cpu_count() or 1
    per = []
    real_ref = reference_available()                                     # baseline/_ref (pip --no-deps install of the reference)
    for _ in range(max(1, a.warmup) if a.warmup < 2 else 1):
        cpu_sample(w, "fp32", 2.0, cores, real_ref)
    budget = max(2.0, min(a.cpu_seconds, 120.0 / max(1, a.steps)))
    for _ in range(a.steps):
        per.append(cpu_sample(w, "fp32", budget, cores, real_ref))
    val = float(np.mean([p["value"] for p in per]))
    ms = float(np.mean([p["seconds"] for p in per]) * 1e3)
    sample = ("%d of %d rows per step, %s, %d worker processes" % (
        per[0]["rows"], w["rows"],
        "the UNMODIFIED reference opticomlib.devices.FIBER (float32/complex64 as shipped; baseline/_ref, NumPy single-threaded "
        "per process)" if real_ref else
        "reference algorithm as shipped (float32/complex64, devices.py:1137-1196) via the oracle port (baseline/_ref absent)",
        per[0]["cores"]))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": DESCR[a.workload], "rows": w["rows"], "samples_per_row": w["n"], **w["fiber"]},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": per[0]["cores"], "kind": "reference" if real_ref else "port", "sample": sample,
                         "one_core": per[0]["one_core"]},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1 and not os.environ.get("SSFM_NO_NUMA_BIND"):
        from opticomlib_b200.scheduler import bind_to_gpu_numa_node
        numa = bind_to_gpu_numa_node(local)                          # before any pinned allocation: host buffers NUMA-local
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from opticomlib_b200 import engine, devices

    if a.workload == "cfg5":
        return run_cfg5(a, torch, dist, dev, rank, world, local)

    w = workload(a.workload, a.rows)
    n, rows_total = w["n"], w["rows"]
    from opticomlib_b200.scheduler import row_shard
    rows = max(1, row_shard(rows_total, world, rank).count)            # strong scaling: rows sharded over ranks
    tdtype = torch.complex128 if a.precision == "fp64" else torch.complex64
    csize = 16 if a.precision == "fp64" else 8

    # ---- synthetic inputs, resident in HBM ---------------------------------------------------
    base = torch.from_numpy(w["base"]).to(dev)
    if rows_total > 1:
        G = 10 ** (w["gain_db"] / 10)
        p_ase = 10 ** (w["nf_db"] / 10) * 6.62607015e-34 * (299792458.0 / 1550e-9) * (G - 1) * w["fs"]
        gen = torch.Generator(device=dev); gen.manual_seed(1000 + rank)
        x0 = torch.empty((rows, n), dtype=torch.complex128, device=dev)
        for r0 in range(0, rows, 256):                                   # bounded temporaries
            r1 = min(rows, r0 + 256)
            nz = torch.randn((r1 - r0, n, 2), dtype=torch.float64, device=dev, generator=gen)
            x0[r0:r1] = (G ** 0.5) * base + (p_ase / 4) ** 0.5 * torch.view_as_complex(nz)
        del nz
    else:
        x0 = base.reshape(1, n).clone()
    x0 = x0.to(tdtype)
    work = torch.empty_like(x0)
    plan = engine.get_plan(n, 1, rows, tdtype, dev)
    chunk = a.chunk if a.chunk >= 0 else 0
    plan.set_option("chunk_waveforms", chunk)
    plan.set_option("persistent", 1 if a.schedule == "persistent" else 0)
    plan.set_option("teams", a.teams)
    fiber = w["fiber"]

    def one_step():
        work.copy_(x0)                                                 # restore the input (untimed, device to device)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        info = plan.propagate(work, w["dt"], **fiber)
        e1.record(); e1.synchronize()
        return e0.elapsed_time(e1), info

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 0)):
        one_step()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = engine.launch_count()
    ms_total, units, kern_ms, kern_units = 0.0, 0, 0.0, 0
    t_wall = time.perf_counter()
    for _ in range(a.steps):
        ms, info = one_step()
        ms_total += ms
        units += info.sample_steps(n)
        kind, teams, kms1 = plan.last_timing()
        if kind == 2:                                                   # the propagation was ONE launch of k_wf
            kern_ms += kms1
            kern_units += info.sample_steps(n)
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = engine.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    steps_per_row = float(info.steps.mean())

    agg = torch.tensor([ms_total, float(units), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = agg.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = agg.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, units, launches = float(mx[0]), float(sm[1]), float(sm[2])
    value = units / (ms_total * 1e-3)

    # ---- per-kernel device times (CUDA events on the launching stream) for the roofline object -----------
    persistent = kern_units > 0
    kms = [0.0, 0.0, 0.0]
    if not persistent:
        work.copy_(x0)
        kms = plan.time_step_kernels(work, w["dt"], reps=5, **{**fiber, "h": 0.01})
    rows_timed = rows

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timing ----
    xh = torch.empty(x0.shape, dtype=torch.complex128, pin_memory=True)
    xh.copy_(x0.to(torch.complex128))
    out_h = torch.empty(x0.shape, dtype=tdtype, pin_memory=True)
    del x0, work                                                        # the host path stages its own chunks
    engine.clear_plans(); torch.cuda.empty_cache()
    sched = dict(persistent=(a.schedule == "persistent"))
    single_launch = devices.HOST_SINGLE_LAUNCH and a.schedule == "persistent" and tdtype == torch.complex128 and 4096 <= n <= (1 << 20)
    devices.fiber_batch(xh, w["dt"], precision=a.precision, out=out_h, **sched, **fiber)   # warm
    barrier()
    e2e_units, t0 = 0, time.perf_counter()
    for _ in range(a.steps):
        _, info_h = devices.fiber_batch(xh, w["dt"], precision=a.precision, out=out_h, **sched, **fiber)
        e2e_units += info_h.sample_steps(n)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    agg = torch.tensor([t_e2e, float(e2e_units)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = agg.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = agg.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t_e2e, e2e_units = float(mx[0]), float(sm[1])
    e2e_val = e2e_units / t_e2e
    h2d = int(xh.numel() * 16) * world
    d2h = int(out_h.numel() * csize) * world
    e2e_ms = t_e2e * 1e3 / a.steps

    # ---- parity of what was just timed (outside the timed region): rows of the e2e output against the oracle ----
    e2e_parity = None
    if rank == 0:
        try:
            from oracle.ssfm_oracle import oracle_fiber, rel_l2
            pick = sorted(set([0, 1, 17 % rows, rows - 1]))
            worst, same_steps = 0.0, True
            for b in pick:
                with np.errstate(all="ignore"):
                    ref = oracle_fiber(xh[b].numpy(), w["dt"], real=np.float64 if a.precision == "fp64" else np.float32, **fiber)
                worst = max(worst, rel_l2(out_h[b].numpy(), ref["out"]))
                same_steps &= int(info_h.steps[b]) == int(ref["steps"])
            e2e_parity = {"rows_checked": pick, "max_rel_l2": worst, "steps_equal": bool(same_steps),
                          "tolerance": 1e-10 if a.precision == "fp64" else 1e-4,
                          "against": "oracle/ssfm_oracle.py (restatement of devices.py:1137-1196, pinned to the reference)"}
        except Exception as e:
            e2e_parity = {"error": repr(e)}
    del xh, out_h
    engine.clear_plans(); torch.cuda.empty_cache()

    # ---- the other BASELINE configurations (all ranks take part when world > 1) ----
    extra = {}
    if not a.no_extra:
        try:
            extra = secondary(torch, engine, dev, a) if world == 1 else {}
        except Exception as e:                                          # secondary numbers never break the line
            extra = {"error": repr(e)}
        if a.workload == "cfg3":
            try:
                extra["e2e_edfa_on_device"] = extra_generated(torch, dist, dev, rank, world, a, w, rows, tdtype, csize)
            except Exception as e:
                extra["e2e_edfa_on_device"] = {"error": repr(e)}
            barrier()
        for name, fn in (("cfg4_receiver_1024x2^18_fp64", extra_cfg4), ("cfg5_2^26_fp64", extra_cfg5)):
            if name.startswith("cfg5") and world == 1:
                continue                                                # one GPU: secondary() already ran the 2^26 waveform
            try:
                extra[name] = fn(torch, dist, dev, rank, world)
            except Exception as e:
                extra[name] = {"error": repr(e)}
            barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, timed live with CUDA events on the launching stream ----
    peak, peak_src = measured_peaks()
    step_bytes = 4 * csize                                             # 2 reads + 2 writes per sample*step
    per_gpu = value / world
    traffic_tab = {}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            traffic_tab = json.load(open(tr))
        except Exception:
            pass
    if persistent:
        # the whole propagation is ONE launch of k_wf: algorithmic bytes per launch = 4 x sizeof(C) x the
        # sample*steps that launch advanced; duration = CUDA events around the launch on its stream (rank 0)
        per_launch_units = kern_units / a.steps
        alg_bytes = step_bytes * per_launch_units
        launch_ms = kern_ms / a.steps
        achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
        # DRAM bytes of one launch: an ncu capture (dram__bytes_read.sum + dram__bytes_write.sum) of THIS command at THIS size when
        # profiles/traffic.json holds one (rows per GPU must match), else the capture's bytes per sample*step scaled to this launch
        # and labelled as such.  One read + one write of the field is the floor for a batch that does not fit in L2.
        cap = traffic_tab.get("k_wf_%s_launch" % a.precision) or {}
        traffic, traffic_src = None, None
        if cap.get("rows") == rows and cap.get("samples_per_row") == n:
            traffic, traffic_src = cap["dram_bytes"], "measured: " + cap.get("source", "ncu capture of the same command and size")
        elif cap.get("dram_bytes") and cap.get("sample_steps"):
            traffic = cap["dram_bytes"] / cap["sample_steps"] * per_launch_units
            traffic_src = "extrapolated from a %d-row capture (%s)" % (cap.get("rows", 0), cap.get("source", "ncu"))
        roofline = {
            "bound": "hbm", "kernel": "k_wf", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "field_bytes_1r_1w": 2 * csize * rows * n,
            "peak_source": peak_src,
            "kernel_ms": {"k_wf": launch_ms}, "kernel_share_of_step": {"k_wf": launch_ms / (ms_total / a.steps)},
            "teams_in_flight": teams,
            "note": "one persistent propagation = two concurrent launches of k_wf (16-CTA clusters; clusters of 2 CTAs with 8 tiles "
                    "each in the CTA slots those leave), timed as one with CUDA events on the launching stream; achieved = 64 B "
                    "(fp64) / 32 B (fp32) x sample*steps of the propagation / that duration.  The waveforms in flight stay "
                    "L2-resident, so DRAM traffic is far BELOW the algorithmic bytes (one field read + one write per propagation) "
                    "and the kernel is bound by the SM -- FP64 pipe plus L1TEX, see DESIGN.md section 3d -- not by HBM: frac can "
                    "exceed what a DRAM-streaming schedule could reach",
            "algorithmic_bytes_per_launch": alg_bytes,
            "step": {"bytes_per_sample_step": step_bytes, "achieved": per_gpu * step_bytes / 1e9,
                     "frac": per_gpu * step_bytes / 1e9 / peak, "unit": "GB/s per GPU"},
        }
    else:
        # multi-launch schedule: a step is k_row + k_col_mid (k_col_fwd only opens the propagation)
        names = ["k_col_fwd(first step only)", "k_row", "k_col_mid"]
        dom = 1 + int(np.argmax(kms[1:]))
        samples_launch = rows_timed * n
        alg_bytes = 2 * csize * samples_launch                         # 1 field read + 1 field write per launch
        achieved = alg_bytes / (kms[dom] * 1e-3) / 1e9
        roofline = {
            "bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
            "kernel_ms": dict(zip(names, kms)), "kernel_share_of_step": {k: v / sum(kms[1:]) for k, v in zip(names[1:], kms[1:])},
            "note": "achieved = 1 field read + 1 field write per launch / CUDA-event time of that kernel; k_col_mid also moves "
                    "the real-valued Kerr-phase stash (+1 read +1 write of R per sample, not counted as algorithmic)",
            "algorithmic_bytes_per_launch": alg_bytes,
            "step": {"bytes_per_sample_step": step_bytes, "achieved": per_gpu * step_bytes / 1e9,
                     "frac": per_gpu * step_bytes / 1e9 / peak, "unit": "GB/s per GPU"},
        }
        t1 = traffic_tab.get("%s_%s_bytes_per_sample" % (names[dom], a.precision))
        if t1 is not None:
            roofline["traffic"] = t1 * samples_launch

    cpu = None
    if not a.no_extra and world == 1:                                   # CPU baseline: one-GPU runs only
        c = cpu_sample(w, a.precision, a.cpu_seconds)
        cpu = {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": "port",
               "sample": "%d of %d rows, oracle port of devices.py:1137-1196 in %s, %d worker processes, %.1f s"
                         % (c["rows"], w["rows"], a.precision, c["cores"], c["seconds"]),
               "one_core": c["one_core"], "cupy_path": cupy_note()}
    for k, v in list(extra.items()):                                    # every extra carries its fraction of the HBM roofline
        if isinstance(v, dict) and "value" in v and "roofline" not in v:
            b = v.get("bytes_per_unit", 64 if "fp32" not in k else 32)
            v["roofline"] = {"bound": "hbm", "achieved": v["value"] / max(1, v.get("n_gpus", 1)) * b / 1e9, "peak": peak, "unit": "GB/s per GPU",
                             "frac": v["value"] / max(1, v.get("n_gpus", 1)) * b / 1e9 / peak, "bytes_per_unit": b}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if a.precision == "fp64" else "f32", "data": "synthetic",
        "config": {"workload": DESCR[a.workload], "rows": rows_total, "rows_per_gpu": rows, "samples_per_row": n,
                   "steps_per_row_mean": steps_per_row, "parallelism": "rows sharded x%d, no data-path collective" % world,
                   "schedule": a.schedule, "chunk_waveforms": chunk, "l2": "inputs larger than L2 (%.0f MiB per GPU); input restored by an "
                   "untimed device copy before each timed propagation" % (rows * n * csize / 2 ** 20), **fiber},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "exposed_copy_ms": e2e_ms - ms_total / a.steps,
                "chunks_per_gpu": (-(-rows // max(1, devices.HOST_SINGLE_CHUNK_BYTES // (n * csize), -(-rows // 256))) if single_launch
                                   else -(-rows // devices.host_chunk_rows(rows, 1, n, tdtype))),
                "numa_node_of_rank0": (numa[0] if numa else None),
                "api": ("opticomlib_b200.fiber_batch(pinned host complex128 -> pinned host complex128): ONE persistent launch per "
                        "propagation that adopts a waveform once the host-to-device copy of its ~32 MiB chunk has been flagged "
                        "(ssfm_propagate_streamed; stream memory operations), chunks copied back by a third stream as the kernel "
                        "counts them finished; exposed_copy_ms = e2e ms per step - device-resident ms per step" if single_launch else
                        "opticomlib_b200.fiber_batch(pinned host complex128 -> pinned host %s), rows streamed in chunks over %d "
                        "streams, one launch per chunk (H2D / propagate / D2H overlapped, one enqueueing host thread); exposed_copy_ms "
                        "= e2e ms per step - device-resident ms per step (the first chunk's H2D and the last chunk's D2H)"
                        % ("complex128" if csize == 16 else "complex64", devices.HOST_LANES))},
        "e2e_parity": e2e_parity,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "wall_s_timed_region": t_wall,
        "extra": extra,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_cfg5(a, torch, dist, dev, rank, world, local):
    """One long waveform (2^log2n samples) through 100-km spans, fixed h = 1 km; columns of the N0 x N_l sample
    matrix spread over the ranks, two all-to-all exchanges per split step."""
    from opticomlib_b200 import engine, longwave as lw, workloads as wl
    n = 1 << a.log2n
    c = wl.CONFIGS["cfg5"]
    fiber = dict(c["fiber"])
    dt = 1.0 / (c["R"] * c["sps"])
    tdtype = torch.complex128 if a.precision == "fp64" else torch.complex64
    csize = 16 if a.precision == "fp64" else 8
    group = dist.group.WORLD if world > 1 else None
    plan = lw.get_long_plan(n, tdtype, dev, group, None, a.exchange == "fused")
    # synthetic PRBS23 NRZ field, built directly in this rank's layout [N0][N_l/G] (the reference DAC FIR would be 2^26 taps)
    bits = torch.from_numpy(wl.prbs(c["order"], n // c["sps"]).astype(np.float64)).to(dev)
    na = torch.arange(plan.n_outer, device=dev, dtype=torch.int64)[:, None]
    nb = torch.arange(plan.rank * plan.cols, (plan.rank + 1) * plan.cols, device=dev, dtype=torch.int64)[None, :]
    drive = bits[(na * plan.n_inner + nb) // c["sps"]] * 5.0 - 2.5
    carrier = (10 ** (c["p0_dbm"] / 10) * 1e-3) ** 0.5
    g = np.pi / 2 / 5.0 * (drive - 2.5)
    eta = 2 * (10 ** (-26.0 / 10)) ** 0.5
    x0 = (carrier * (10 ** (-3.0 / 10)) ** 0.5 * torch.complex(torch.cos(g), eta / 2 * torch.sin(g))).to(tdtype).contiguous()
    del bits, drive, g, na, nb
    work = torch.empty_like(x0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_span():
        work.copy_(x0)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        info = plan.propagate(work, dt, **fiber)
        e1.record(); e1.synchronize()
        return e0.elapsed_time(e1), info

    for _ in range(max(a.warmup, 0)):
        one_span()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = engine.launch_count()
    ms_total, units = 0.0, 0
    for _ in range(a.steps):
        ms, info = one_span()
        ms_total += ms
        units += int(info.steps[0]) * n
    barrier()
    launches = engine.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    agg = torch.tensor([ms_total, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = agg.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = agg.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, launches = float(mx[0]), float(sm[1])
    value = units / (ms_total * 1e-3)

    # end to end: host (pinned) -> device -> host of this rank's share around one span
    xh = torch.empty(x0.shape, dtype=tdtype, pin_memory=True); xh.copy_(x0)
    oh = torch.empty_like(xh)
    barrier()
    t0 = time.perf_counter()
    e2e_units = 0
    for _ in range(a.steps):
        work.copy_(xh, non_blocking=True)
        info = plan.propagate(work, dt, **fiber)
        oh.copy_(work, non_blocking=True)
        torch.cuda.synchronize()
        e2e_units += int(info.steps[0]) * n
    barrier()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); t_e2e = float(t[0])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    step_bytes = 4 * csize
    per_gpu = value / world
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if a.precision == "fp64" else "f32", "data": "synthetic",
        "config": {"workload": DESCR["cfg5"], "samples": n, "n_outer": plan.n_outer, "n_inner": plan.n_inner,
                   "split_steps_per_bench_step": int(info.steps[0]),
                   "parallelism": "columns of the %d x %d sample matrix over %d rank(s); 2 exchanges per split step (%s)" %
                                  (plan.n_outer, plan.n_inner, world, "none: one rank" if world == 1 else
                                   ("fused: kernels store into peer memory over NVLink" if plan.fused else "NCCL all-to-all + re-layout copy")),
                   "l2": "inputs larger than L2 (%.0f MiB per GPU)" % (x0.numel() * csize / 2 ** 20), **fiber},
        "e2e": {"value": e2e_units / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(x0.numel() * csize) * world,
                "d2h_bytes_per_step": int(x0.numel() * csize) * world,
                "api": "LongPlan.propagate on this rank's share, pinned host -> device -> pinned host around each span"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm" if world == 1 else "nvlink", "kernel": "split step of a long waveform: k_col_mid (outer) + "
                     "k_col_fwd, k_row, k_col_inv (inner)", "achieved": per_gpu * step_bytes / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": per_gpu * step_bytes / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "note": "achieved = 64 B (fp64) / 32 B (fp32) per sample*step x this GPU's throughput; the staged "
                             "transform moves 4 reads + 4 writes of the field per split step (twice the ideal) and, with more "
                             "than one rank, (G-1)/G of the field crosses NVLink twice per split step"},
        "cpu_baseline": None,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cupy_note():
    """The reference's GPU path is CuPy dispatch (devices.py:1114-1119); BASELINE.md section 4 asks for it 'where it installs
    offline'."""
    try:
        import cupy  # noqa: F401
        return "cupy importable: not timed (the reference's CuPy path is not the optimisation target)"
    except Exception as e:
        return "unavailable offline: import cupy fails (%s); no wheel in /opt/wheelhouse, no network" % type(e).__name__


def _events(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def extra_generated(torch, dist, dev, rank, world, a, w, rows, tdtype, csize):
    """Config #3 end to end with the EDFA noise realisations generated ON the device: the host sends one
    waveform (1 MiB) and receives the propagated batch; generation, propagation and the D2H copies of different chunks
    overlap (opticomlib_b200.edfa_fiber_batch).  Same rows per GPU, same fibre, same metric as the headline."""
    from opticomlib_b200 import devices, engine
    import opticomlib_b200 as ob
    ob.gv(sps=w["sps"], R=w["rate"])
    n = w["n"]
    out_h = torch.empty((rows, n), dtype=tdtype, pin_memory=True)
    base = torch.from_numpy(w["base"]).pin_memory()
    from opticomlib_b200.scheduler import row_shard
    first = row_shard(w["rows"], world, rank).start                  # this rank's rows of the global batch: global noise indices
    kw = dict(seed=1000, precision=a.precision, out=out_h, first_row=first, **w["fiber"])
    devices.edfa_fiber_batch(base, rows, w["gain_db"], w["nf_db"], w["dt"], **kw)      # warm
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    units, t0 = 0, time.perf_counter()
    for _ in range(a.steps):
        _, info = devices.edfa_fiber_batch(base, rows, w["gain_db"], w["nf_db"], w["dt"], **kw)
        units += info.sample_steps(n)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    agg = torch.tensor([t, float(units)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = agg.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = agg.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t, units = float(mx[0]), float(sm[1])
    del out_h
    engine.clear_plans(); torch.cuda.empty_cache()
    return {"value": units / t, "unit": UNIT, "n_gpus": world, "ms_per_step": t * 1e3 / a.steps, "bytes_per_unit": 4 * csize,
            "h2d_bytes_per_step": int(base.numel() * 16) * world, "d2h_bytes_per_step": int(rows * n * csize) * world,
            "api": "opticomlib_b200.edfa_fiber_batch(one pinned complex128 waveform -> pinned host batch): ASE realisations from the "
                   "extension's Philox generator on the device (reference EDFA devices.py:921-936), then FIBER, then D2H"}


def extra_cfg4(torch, dist, dev, rank, world):
    """BASELINE config #4 at full size: 1024 received frames x 2^18 samples (fp64), frames sharded over the ranks, receiver
    chain BPF(40 GHz) -> 10 x [x 10^(-16/20); DBP 80 km at h = 10 km] -> PD square law + LPF(7.5 GHz), every stage timed with
    CUDA events on the launching stream after one warm-up pass (max over ranks).  Rooflines: filters against 1 read + 1 write
    of the complex128 frames (32 B/sample), DBP against 64 B/sample*step."""
    from opticomlib_b200 import devices, engine, workloads as wl
    from scipy import signal as sg
    c4 = wl.CFG4_RX
    fs = wl.CONFIGS["cfg4"]["R"] * wl.CONFIGS["cfg4"]["sps"]
    frames_total, n4 = wl.CONFIGS["cfg4"]["rows"], 1 << 18
    frames = max(1, frames_total // world)
    base4 = torch.from_numpy(wl.ook_field(15, 4096, 64, 0.0)).to(dev)
    gen = torch.Generator(device=dev); gen.manual_seed(2000 + rank)
    rx0 = base4.repeat(frames, 1) * (1 + 0.01 * torch.rand((frames, 1), device=dev, dtype=torch.float64, generator=gen))
    sos_b = sg.bessel(4, c4["bpf_bw"] / 2, "low", fs=fs, output="sos", norm="mag")
    sos_l = sg.bessel(4, c4["lpf_bw"], "low", fs=fs, output="sos", norm="mag")
    y = torch.empty_like(rx0)
    gain = 10 ** (-c4["span_loss_db"] / 20)
    best = None
    for it in range(3):                                               # pass 0 warms plans, tables and the allocator
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        engine.filtfilt_sos(rx0, sos_b, out=y)                        # received frames -> filtered frames (as devices.BPF returns a new signal)
        ev[1].record()
        nsteps = 0
        for _ in range(c4["spans"]):
            y.mul_(gain)
            _, info = devices.dbp_batch(y, 1.0 / fs, precision="fp64", inplace=True, **c4["dbp"])
            nsteps += int(info.steps.sum())
        ev[2].record()
        z = devices.pd_lpf_batch(y, sos_l, responsivity=1.0)          # |.|^2 fused into the filter's first pass
        ev[3].record(); ev[3].synchronize()
        cur = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)] + [nsteps]
        if it > 0 and (best is None or sum(cur[:3]) < sum(best[:3])):
            best = cur
        del z
    t = torch.tensor(best[:3], dtype=torch.float64, device=dev)
    u = torch.tensor([float(best[3]) * n4], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    bpf_ms, dbp_ms, lpf_ms = [float(v) for v in t]
    peak, _ = measured_peaks()
    samples = frames * n4
    out = {
        "frames": frames_total, "frames_per_gpu": frames, "n_gpus": world, "samples_per_frame": n4, "unit": UNIT,
        "value": float(u[0]) / (dbp_ms * 1e-3), "bytes_per_unit": 64,
        "stage_ms": {"bpf": bpf_ms, "dbp_10_spans": dbp_ms, "pd_lpf": lpf_ms},
        "steps_per_frame": best[3] / frames,
        "roofline_per_stage": {
            "bpf": {"ideal_bytes": 32 * samples, "achieved_GBps": 32 * samples / (bpf_ms * 1e-3) / 1e9, "frac": 32 * samples / (bpf_ms * 1e-3) / 1e9 / peak},
            "dbp": {"ideal_bytes_per_sample_step": 64, "achieved_GBps": float(u[0]) / world * 64 / (dbp_ms * 1e-3) / 1e9,
                    "frac": float(u[0]) / world * 64 / (dbp_ms * 1e-3) / 1e9 / peak},
            "pd_lpf": {"ideal_bytes": 24 * samples, "achieved_GBps": 24 * samples / (lpf_ms * 1e-3) / 1e9, "frac": 24 * samples / (lpf_ms * 1e-3) / 1e9 / peak,
                       "note": "ideal = read the complex128 field (16 B) + write the real float64 photocurrent (8 B)"},
        },
        "timing": "CUDA events on the launching stream, best of 2 after 1 warm-up pass, max over ranks",
    }
    del rx0, y
    engine.clear_plans(); torch.cuda.empty_cache()
    return out


def extra_cfg5(torch, dist, dev, rank, world):
    """BASELINE config #5 on the ranks of this run: two 100-km spans (h = 1 km: 100 split steps each) of one 2^26-sample
    waveform with the exchange fused into the kernels (peer stores over NVLink), the same with the NCCL all-to-all, and -- in
    this process, on the same distributed code path -- the 2^20-sample parity check against the oracle."""
    from opticomlib_b200 import longwave as lw, workloads as wl
    n = 1 << 26
    c = wl.CONFIGS["cfg5"]
    fiber = dict(c["fiber"])
    dt = 1.0 / (c["R"] * c["sps"])
    group = dist.group.WORLD if world > 1 else None
    peak, _ = measured_peaks()
    out = {"samples": n, "n_gpus": world, "unit": UNIT, "bytes_per_unit": 64}
    for mode in (("fused", True), ("nccl", False)):
        if world == 1 and mode[0] == "nccl":
            continue
        plan = lw.get_long_plan(n, torch.complex128, dev, group, None, mode[1])
        gen = torch.Generator(device=dev); gen.manual_seed(5 + rank)
        x0 = torch.view_as_complex(torch.randn((plan.n_outer, plan.cols, 2), dtype=torch.float64, device=dev, generator=gen) * 0.02).contiguous()
        work = torch.empty_like(x0)
        spans = 3 if mode[0] == "fused" else 2
        ms_best, steps = None, 0
        for it in range(spans):
            work.copy_(x0)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = _events(torch)
            e0.record(); info = plan.propagate(work, dt, **fiber); e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1)
            steps = int(info.steps[0])
            if it > 0:
                ms_best = ms if ms_best is None else min(ms_best, ms)
        t = torch.tensor([ms_best], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        rec = {"value": steps * n / (ms * 1e-3), "ms_per_span": ms, "ms_per_split_step": ms / steps, "split_steps": steps,
               "exchange": ("kernels store into peer memory over NVLink (CUDA IPC), flag barrier between the GPUs" if plan.fused else
                            "NCCL all_to_all_single + re-layout copy") if world > 1 else "none (one rank)"}
        if world > 1:
            link = 2 * 16 * (world - 1) / world                       # bytes per sample*step leaving (and entering) each GPU
            per_gpu = rec["value"] / world
            rec["nvlink_roofline"] = {"bytes_per_sample_step_per_direction": link, "achieved_GBps_per_direction": per_gpu * link / 1e9,
                                      "peak_GBps_per_direction": 900.0, "frac": per_gpu * link / 1e9 / 900.0,
                                      "peak_source": "NVLink 5 nominal, 900 GB/s per direction per GPU"}
        rec["hbm_roofline_frac"] = rec["value"] / world * 64 / 1e9 / peak
        if mode[0] == "fused":
            out.update(rec)
        else:
            out["nccl_exchange"] = rec
        del x0, work
        lw.clear_plans(); torch.cuda.empty_cache()
    # parity of the distributed path at 2^20 samples against the oracle (full length, same code path)
    try:
        from oracle.ssfm_oracle import oracle_fiber, rel_l2
        n20 = 1 << 20
        rng = np.random.default_rng(9)
        tt = np.arange(n20) / n20
        env = np.sqrt(20e-3) * (0.55 + 0.45 * np.sign(np.sin(2 * np.pi * 37 * tt + 0.3)))
        x = np.convolve(env, np.ones(9) / 9, mode="same") * np.exp(2j * np.pi * 3 * tt) + 1e-4 * (rng.standard_normal(n20) + 1j * rng.standard_normal(n20))
        kw = dict(length=0.9, alpha=0.2, beta_2=-21.27, beta_3=0.127, gamma=1.3, h=0.3)
        res, info = lw.fiber_long(x, 1 / 640e9, precision="fp64", group=group, **kw)
        if rank == 0:
            with np.errstate(all="ignore"):
                ref = oracle_fiber(x, 1 / 640e9, real=np.float64, **kw)
            out["parity_2^20"] = {"rel_l2": rel_l2(res, ref["out"]), "steps_equal": int(info.steps[0]) == int(ref["steps"]), "tolerance": 1e-10,
                                  "path": "fiber_long over the %d rank(s) of this run, fused exchange, against oracle/ssfm_oracle.py" % world}
        lw.clear_plans(); torch.cuda.empty_cache()
    except Exception as e:
        out["parity_2^20"] = {"error": repr(e)}
    return out


def secondary(torch, engine, dev, a):
    """Other single-GPU numbers reported next to the headline (not bench lines of their own)."""
    from opticomlib_b200 import workloads as wl
    out = {}
    # BASELINE config #1 (one 2^16-sample waveform, 8 adaptive steps): launch-latency dominated, reported only
    x, dt, kw = wl.config_input("cfg1")
    for prec, td in (("fp64", torch.complex128), ("fp32", torch.complex64)):
        x0 = torch.from_numpy(x).to(dev).to(td).reshape(1, -1)
        work = torch.empty_like(x0)
        plan = engine.get_plan(x0.shape[1], 1, 1, td, dev)
        best = None
        for i in range(5):
            work.copy_(x0); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); info = plan.propagate(work, dt, **kw); e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None or ms < best else best
        out["cfg1_%s" % prec] = {"value": info.sample_steps(x0.shape[1]) / (best * 1e-3), "unit": UNIT,
                                 "steps": int(info.steps[0]), "ms": best}
    x, dt, kw = wl.config_input("cfg2")
    for prec, td in (("fp64", torch.complex128), ("fp32", torch.complex64)):
        x0 = torch.from_numpy(x).to(dev).to(td).reshape(1, -1)
        work = torch.empty_like(x0)
        plan = engine.get_plan(x0.shape[1], 1, 1, td, dev)
        plan.set_option("chunk_waveforms", 0)
        best = None
        for i in range(3):
            work.copy_(x0); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); info = plan.propagate(work, dt, **kw); e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None or ms < best else best
        out["cfg2_%s" % prec] = {"value": info.sample_steps(x0.shape[1]) / (best * 1e-3), "unit": UNIT,
                                 "steps": int(info.steps[0]), "ms": best}
    if a.workload == "cfg3":                                            # the same batch in the other precision
        other = "fp32" if a.precision == "fp64" else "fp64"
        td = torch.complex64 if other == "fp32" else torch.complex128
        w = workload("cfg3", min(a.rows or 4096, 1024))
        base = torch.from_numpy(w["base"]).to(dev)
        x0 = ((10.0 ** 0.5) * base).to(td).repeat(w["rows"], 1)
        x0 = x0 * (1 + 0.01 * torch.rand((w["rows"], 1), device=dev, dtype=torch.float64)).to(td)
        work = torch.empty_like(x0)
        plan = engine.get_plan(w["n"], 1, w["rows"], td, dev)
        plan.set_option("chunk_waveforms", 0)
        best = None
        for i in range(3):
            work.copy_(x0); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); info = plan.propagate(work, w["dt"], **w["fiber"]); e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None or ms < best else best
        out["cfg3_%s_%drows" % (other, w["rows"])] = {"value": info.sample_steps(w["n"]) / (best * 1e-3), "unit": UNIT, "ms": best}
    # a length that is not a power of two (60000 samples: chirp-z transforms of 2^17 points), 64 rows, fp64
    try:
        n_odd = 60000
        base = torch.from_numpy(wl.config_input("cfg1")[0][:n_odd]).to(dev)
        x0 = ((10.0 ** 0.5) * base).repeat(64, 1).contiguous()
        work = torch.empty_like(x0)
        plan = engine.get_plan(n_odd, 1, 64, torch.complex128, dev)
        kw1 = wl.config_input("cfg1")[2]
        best = None
        for i in range(2):
            work.copy_(x0); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); info = plan.propagate(work, 1.0 / 640e9, **kw1); e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None or ms < best else best
        out["arbitrary_length_60000_fp64_64rows"] = {"value": info.sample_steps(n_odd) / (best * 1e-3), "unit": UNIT, "ms": best,
                                                     "steps_per_row_mean": float(info.steps.mean())}
        engine.clear_plans()
    except Exception as e:
        out["arbitrary_length_60000_fp64_64rows"] = {"error": repr(e)}
    # BASELINE config #5's waveform (2^26 samples) on this one GPU: 20 fixed steps of 1 km through the staged transform
    try:
        from opticomlib_b200 import longwave as lw
        n5 = 1 << 26
        plan5 = lw.get_long_plan(n5, torch.complex128, dev)
        gen = torch.Generator(device=dev); gen.manual_seed(5)
        x5 = torch.view_as_complex(torch.randn((plan5.n_outer, plan5.cols, 2), dtype=torch.float64, device=dev, generator=gen) * 0.02).contiguous()
        c5 = dict(wl.CONFIGS["cfg5"]["fiber"]); c5["length"] = 20.0
        best = None
        for i in range(2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); info = plan5.propagate(x5, 1.0 / 640e9, **c5); e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None or ms < best else best
        out["cfg5_2^26_fp64_1gpu_20steps"] = {"value": int(info.steps[0]) * n5 / (best * 1e-3), "unit": UNIT, "ms": best,
                                              "ms_per_split_step": best / int(info.steps[0])}
        del x5
        lw.clear_plans(); torch.cuda.empty_cache()
    except Exception as e:
        out["cfg5_2^26_fp64_1gpu_20steps"] = {"error": repr(e)}
    return out


if __name__ == "__main__":
    main()
