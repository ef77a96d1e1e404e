"""The C-ABI library loads, exports every symbol include/redtime_b200.h declares, and fails
loudly (RTRG_ENOGPU) instead of computing anything when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

import redtime_b200 as rt
from conftest import ROOT


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "redtime_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rtrg_[a-z0-9_A-Z]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    lib = rt.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n


def test_version_and_defaults():
    lib = rt.load_library()
    assert b"sm_100a" in lib.rtrg_version()
    cfg = rt.Config()
    lib.rtrg_default_config(ctypes.byref(cfg))
    # the reference's compile-time constants (redTime.cc:90-97,141-145,1285; hdr:537,664,697)
    assert (cfg.nk, cfg.kmin, cfg.kmax, cfg.z1l) == (128, 1e-3, 1.0, 10.0)
    assert (cfg.eps_abs, cfg.eps_rel) == (1e-7, 1e-2)
    assert (cfg.beta_kmin, cfg.beta_kmax, cfg.n_lnk, cfg.n_lna, cfg.a_early) == (1e-3, 1.0, 50, 100, 1e-20)


def test_invalid_arguments_are_rejected():
    lib = rt.load_library()
    assert lib.rtrg_create(None, None) == -1
    out = (ctypes.c_int * 5)()
    assert lib.rtrg_grid_info(100, 1e-3, 1.0, out, None, None) == -1  # nk must be a multiple of 16
    assert lib.rtrg_grid_info(128, 1.0, 1e-3, out, None, None) == -1
    assert lib.rtrg_num_columns(None, 0) == -1
    assert lib.rtrg_run(None, None, 0, None, None, None) == -1


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_no_cpu_fallback():
    with pytest.raises(rt.RtrgError) as e:
        rt.RedTimeB200()
    assert e.value.code == -2 and "no CPU path" in str(e.value)
    with pytest.raises(rt.RtrgError):
        rt.dfma_peak_tflops()


def test_product_sources_never_touch_the_oracle():
    """Only tests/, bench.py's cpu_baseline / reference arm and __graft_entry__ may use oracle/."""
    pkg = os.path.join(ROOT, "redtime_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle/" not in txt and "gsl_shim" not in txt, os.path.join(dirpath, f)
