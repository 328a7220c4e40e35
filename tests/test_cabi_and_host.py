"""CPU-side checks: the C-ABI library loads and exports exactly what include/ssfm_b200.h declares,
the host-side containers mirror the reference surface, and the product refuses to run without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from opticomlib_b200 import build, _lib
    build.build()                                      # nvcc cross-compiles without a GPU
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from opticomlib_b200 import _lib
    header = open(os.path.join(ROOT, "include", "ssfm_b200.h")).read()
    declared = set(re.findall(r"SSFM_API\s+[\w\s\*]+?\b(ssfm_\w+)\s*\(", header))
    assert declared, "no declarations found in the header"
    assert declared == set(_lib.PROTOTYPES), "ctypes prototypes and header disagree"
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ssfm_abi_version() == 1
    m = re.search(r"#define\s+SSFM_ABI_VERSION\s+(\d+)", header)
    assert int(m.group(1)) == lib.ssfm_abi_version()


def test_struct_layout_matches_header():
    from opticomlib_b200 import _lib
    assert ctypes.sizeof(_lib.FiberParams) == 8 * 8
    assert [f[0] for f in _lib.FiberParams._fields_] == [
        "dt_s", "length_km", "alpha_db_km", "beta2_ps2_km", "beta3_ps3_km", "gamma_w_km", "phi_max_rad", "h_km"]


def test_argument_validation_needs_no_gpu(lib):
    from opticomlib_b200 import _lib
    h = ctypes.c_void_p()
    for bad_n in (1, 3 << 20, 1 << 23):                          # too short; not a power of two above 2^21; above 2^22
        assert lib.ssfm_plan_create(ctypes.byref(h), bad_n, 1, 1, _lib.SSFM_C64, 0) == _lib.SSFM_ERR_UNSUPPORTED
        assert b"power of two" in lib.ssfm_last_error()
    assert lib.ssfm_plan_create(ctypes.byref(h), 4096, 3, 1, _lib.SSFM_C64, 0) == _lib.SSFM_ERR_INVALID
    assert b"n_pol" in lib.ssfm_last_error()
    assert lib.ssfm_plan_create(ctypes.byref(h), 4096, 1, 0, _lib.SSFM_C64, 0) == _lib.SSFM_ERR_INVALID
    assert lib.ssfm_plan_create(ctypes.byref(h), 4096, 1, 1, 7, 0) == _lib.SSFM_ERR_INVALID
    assert lib.ssfm_plan_destroy(None) == 0
    # streamed batches: null arguments are refused; without a driver the stream memory operations report "unsupported"
    # (the host path then pipelines one launch per chunk) -- nothing may crash
    assert lib.ssfm_propagate_streamed(None, None, None, None, None, 1, None) == _lib.SSFM_ERR_INVALID
    assert b"null" in lib.ssfm_last_error()
    for fn in (lib.ssfm_stream_write_u32, lib.ssfm_stream_wait_geq_u32):
        assert fn(None, None, 0) in (_lib.SSFM_ERR_INVALID, _lib.SSFM_ERR_UNSUPPORTED)
    with pytest.raises(ValueError):
        _lib.check(_lib.SSFM_ERR_INVALID)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import opticomlib_b200 as ob
    s = ob.optical_signal(np.ones(1024, complex))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ob.FIBER(s, length=1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ob.LPF(np.ones(100), BW=1e9)
    with pytest.raises(TypeError):                                  # type check comes first, as in the reference
        ob.FIBER(np.ones(8), length=1.0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "opticomlib_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_signal_containers_mirror_reference_surface():
    import opticomlib_b200 as ob
    ob.gv(sps=16, R=1e9)
    assert ob.gv.fs == 16e9 and ob.gv.dt == 1 / 16e9
    e = ob.electrical_signal([1.0, 2.0, 3.0], [0.1, 0.1, 0.1])
    assert e.size == 3 and e.ndim == 1 and e.type is ob.electrical_signal
    assert np.allclose(e.to_numpy(), [1.1, 2.1, 3.1])
    c = e[:]
    c.signal[0] = 9
    assert e.signal[0] == 1.0                                          # slicing copies (typing.py:1366)
    o = ob.optical_signal(np.ones(8, complex))
    assert o.n_pol == 1 and o.size == 8 and o.noise is ob.NULL
    assert np.array_equal(o.to_numpy(), np.ones(8, complex))           # NULL + x -> x
    o2 = ob.optical_signal(np.ones((2, 8), complex), 0.5 * np.ones((2, 8), complex))
    assert o2.n_pol == 2 and o2.size == 8 and o2.to_numpy().shape == (2, 8)
    assert np.allclose(o.w(), 2 * np.pi * np.fft.fftfreq(8) * ob.gv.fs)   # typing_test.py:1244-1251
    assert ob.optical_signal(np.ones(8), n_pol=2).signal.shape == (2, 8)
    with pytest.raises(ValueError):
        ob.optical_signal(np.ones((3, 8)))
    with pytest.raises(ValueError):
        ob.electrical_signal(np.ones((2, 8)))


def test_reference_objects_are_accepted_by_duck_typing(have_reference):
    if not have_reference:
        pytest.skip("reference tree not present")
    from oracle.ref_shim import import_reference
    import_reference()
    import opticomlib
    import opticomlib.devices as rdv
    from opticomlib_b200 import devices as dv
    r = opticomlib.optical_signal(np.ones(16, complex))
    assert dv._kind(r) == "optical" and dv._gv_of(r) is opticomlib.gv
    assert dv._kind(opticomlib.electrical_signal(np.ones(4))) == "electrical"
    assert dv._is_null(r.noise)
    original = rdv.FIBER
    dv.install()
    try:
        assert rdv.FIBER is dv.FIBER and rdv.LPF is dv.LPF and rdv.BPF is dv.BPF and rdv.DBP is dv.DBP
    finally:
        dv.uninstall()
    assert rdv.FIBER is original


def test_host_chunk_rule_and_length_validation():
    """Host-side logic of the end-to-end pipeline (no GPU): at least 12 chunks for large batches, a few rows per team in flight,
    never above the byte cap; lengths the kernels do not support are refused up front with the supported set spelled out."""
    import torch
    from opticomlib_b200 import devices as dv
    n = 1 << 16
    for B in (4096, 2048, 1024, 512):
        rows = dv.host_chunk_rows(B, 1, n, torch.complex128)
        assert -(-B // rows) >= 12 and rows * n * 16 <= dv.HOST_CHUNK_BYTES
        assert rows >= 36                                     # two rows per team in flight (18 teams of 16 CTAs)
    assert dv.host_chunk_rows(3, 1, n, torch.complex128) == 3 and dv.host_chunk_rows(1, 2, n, torch.complex64) == 1
    assert -(-44 // dv.host_chunk_rows(44, 1, n, torch.complex128)) == dv.HOST_LANES
    for ok in (2, 3, 1000, 60000, (1 << 21) - 1, 1 << 21, 1 << 22, 1 << 26, 1 << 30):
        dv._check_length(ok)
    for bad in (0, 1, (1 << 21) + 1, 3 << 20, (1 << 30) + 1, 1 << 31):
        with pytest.raises(ValueError, match="supports waveforms"):
            dv._check_length(bad)


def test_single_launch_chunks_cover_the_batch():
    """Copy chunks around the single streamed launch: contiguous, complete, ~32 MiB each, never more than 256."""
    from opticomlib_b200 import devices
    for B, row_bytes in ((4096, 1 << 20), (512, 1 << 20), (41, 1 << 20), (1, 1 << 20), (100000, 4096 * 16), (7, 1 << 30)):
        rows, chunks = devices.single_launch_chunks(B, row_bytes)
        assert chunks[0][0] == 0 and chunks[-1][1] == B and len(chunks) <= 256
        assert all(a1 == b0 for (_, a1), (b0, _) in zip(chunks, chunks[1:]))
        assert all(0 < r1 - r0 <= rows for r0, r1 in chunks)
        if B * row_bytes > 256 * devices.HOST_SINGLE_CHUNK_BYTES:
            assert len(chunks) in (255, 256) or rows * 256 >= B
        elif row_bytes <= devices.HOST_SINGLE_CHUNK_BYTES:
            assert rows * row_bytes <= devices.HOST_SINGLE_CHUNK_BYTES


def test_pd_argument_validation_matches_the_reference_without_a_gpu():
    """PD validates before it touches the device: same exception types and messages as devices.py:1493-1512."""
    import opticomlib_b200 as ob
    x = ob.optical_signal(np.ones(64, complex))
    with pytest.raises(TypeError, match="optical_signal"):
        ob.PD(np.ones(8), BW=1e9)
    with pytest.raises(ValueError, match=r"`r` must be in the range \(0,1\]"):
        ob.PD(x, BW=1e9, r=0.0)
    with pytest.raises(TypeError, match="`T` must be a scalar"):
        ob.PD(x, BW=1e9, T="hot")
    with pytest.raises(ValueError, match="`R_load` must be a positive"):
        ob.PD(x, BW=1e9, R_load=-1.0)
    with pytest.raises(TypeError, match="`include_noise` must be a string"):
        ob.PD(x, BW=1e9, include_noise=3)


def test_numa_binding_is_a_no_op_without_topology():
    from opticomlib_b200.scheduler import bind_to_gpu_numa_node
    import os
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None                   # no CUDA device here: nothing is changed
    assert os.sched_getaffinity(0) == before
