"""CPU checks of the boundary: the C-ABI library builds, loads and exports every symbol ``include/smcb200.h`` declares; without
a CUDA device the compute entry points refuse to run (there is no CPU fallback); the exact-scan arithmetic passes its host test."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

from pyfilter_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "smcb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smcb_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load_library()
    names = _declared_symbols()
    assert len(names) >= 20
    bound = {n for n, _, _ in _lib.SYMBOLS}
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/smcb200.h but not exported"
        assert n in bound, f"{n} has no ctypes prototype in pyfilter_b200/_lib.py"
    assert lib.smcb_version() == 200


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.smcb_config) == 88   # (the last two int32 fields fill the padding of the 8-byte aligned struct)
    assert C.sizeof(_lib.smcb_info) == 64


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a device")
def test_no_cpu_fallback():
    lib = _lib.load_library()
    assert lib.smcb_device_count() == 0
    params = (C.c_float * 3)(-1.0, 0.97, 0.2)
    cfg = _lib.smcb_config(model=2, proposal=0, algorithm=1, resampler=0, particles=1000, batch=1, n_raw_params=3,
                           params_host=C.cast(params, C.POINTER(C.c_float)), param_cols=1, ess_threshold=0.9, seed=1, history_rows=4,
                           fold_lookahead=1, exact_scan=1, reserved=0)
    h = C.c_void_p()
    assert lib.smcb_filter_create(C.byref(cfg), C.byref(h)) == -3  # SMCB_ENODEVICE
    assert b"no CPU fallback" in lib.smcb_last_error()
    assert lib.smcb_systematic(None, 10, 1, 1, 0, 1, None, 0, None, 1, 10, None) in (-1, -3)
    import pyfilter_b200 as pf

    with pytest.raises(_lib.SmcbError):
        pf.resampling.systematic(torch.zeros(8))
    with pytest.raises(_lib.SmcbError):
        pf.filters.particle.APF(pf.timeseries.build("sv_ar1"), 100).initialize()


def test_unsupported_combinations_raise():
    import pyfilter_b200 as pf

    with pytest.raises(NotImplementedError):
        pf.resampling.residual(torch.zeros(4, 2))  # one column only, like the reference (resampling.py:78-79)
    with pytest.raises(NotImplementedError):
        pf.filters.particle.APF(pf.timeseries.build("sv_ar1"), 100, resampling=lambda w: w)


def test_exact_scan_host_arithmetic(tmp_path):
    """tests/host/test_exact_scan.cpp: the tile algorithm, the probe counts and the integer transducers against the plain
    sequential definitions; Philox4x32-10 against the Random123 known answer and the keyed variant the kernels use (g++, no GPU)."""
    exe = str(tmp_path / "test_exact_scan")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "test_exact_scan.cpp")], check=True)
    out = subprocess.run([exe, "200000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout
