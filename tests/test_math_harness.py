"""The scalar arithmetic of the CUDA kernels (redtime_b200/csrc/rtrg_math.h), compiled with g++
as TEST INFRASTRUCTURE (tests/harness/math_harness.cc), against the oracle's stage goldens:
table look-up rules, growth ODE with GSL's RK8PD step control, QAG-61 normalisation, linear
spectra and the Time-RG right-hand side.  No GPU needed; the shipped library has no such path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import redtime_b200 as rt
from conftest import ROOT

dp = C.POINTER(C.c_double)


def P(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def harness(tmp_path_factory, example1_dir):
    so = str(tmp_path_factory.mktemp("mh") / "libmh.so")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC",
                    os.path.join(ROOT, "tests", "harness", "math_harness.cc"), "-o", so], check=True)
    lib = C.CDLL(so)
    lib.mh_norm.restype = C.c_double
    lib.mh_sigv2_0.restype = C.c_double
    c = rt.read_run_dir(example1_dir)
    f = lambda k: np.ascontiguousarray(c[k], dtype=np.float64)  # noqa: E731
    keep = [f(k) for k in ("params", "k_T", "Tc_T", "Tb_T", "z_interp", "k_b", "Tc_b", "Tnu_b")]
    lib.mh_setup.argtypes = [dp, C.c_double, C.c_int, dp, dp, dp, C.c_int, dp, C.c_int, dp, dp, dp]
    assert lib.mh_setup(P(keep[0]), c["z_in"], keep[1].size, P(keep[1]), P(keep[2]), P(keep[3]), keep[4].size,
                        P(keep[4]), keep[5].size, P(keep[5]), P(keep[6]), P(keep[7])) == 0
    return lib


def rel(a, b):
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-300)))


def test_lookups_growth_and_normalisation(harness, stage_golden_1loop):
    g = stage_golden_1loop
    k = np.ascontiguousarray(g["lin_k"])
    harness.mh_lookups.argtypes = [C.c_double, dp, C.c_int] + [dp] * 6
    for iz, z in enumerate(g["lin_z"]):
        out = [np.zeros(k.size) for _ in range(6)]
        harness.mh_lookups(float(z), P(k), k.size, *[P(o) for o in out])
        D, dD, beta, Pl, Pcb, Pnu = out
        assert rel(D, g["lin_D"][iz]) < 1e-12 and rel(dD, g["lin_dD"][iz]) < 1e-12, z  # RK8PD replay + 2-D rule
        assert rel(beta, g["lin_beta"][iz]) < 1e-14, z
        assert rel(Pl, g["lin_P"][iz]) < 1e-11 and rel(Pcb, g["lin_Pcb"][iz]) < 1e-11 and rel(Pnu, g["lin_Pnu"][iz]) < 1e-11
    assert abs(harness.mh_sigv2_0() / g["sigmaV2"][-1] - 1) < 1e-11   # QAG-61 bisection sequence


def test_time_rg_right_hand_side(harness, stage_golden_full):
    g = stage_golden_full
    nk = 128
    harness.mh_rhs.argtypes = [C.c_double, C.c_int, dp, dp, dp, dp, C.c_int, dp]
    k, y = np.ascontiguousarray(g["k"]), np.ascontiguousarray(g["yp"])
    A, R = np.ascontiguousarray(g["A_yp"]), np.ascontiguousarray(g["R_yp"])
    for eta, ref in zip(g["rhs_eta"], g["rhs_dy"]):
        dy = np.zeros(41 * nk)
        harness.mh_rhs(float(eta), nk, P(k), P(y), P(A), P(R), 1, P(dy))
        # same sources as the oracle used, so every row agrees to round-off
        s = np.maximum(np.abs(ref), 1e-12 * np.abs(ref).max())
        assert np.max(np.abs(dy - ref) / s) < 1e-10, eta


def test_gsl_step_size_control(harness):
    """std_control_hadjust for control_y_new, checked on App. C's trace."""
    harness.mh_hadjust.argtypes = [C.c_double, C.c_int, dp]
    cases = [(6.827500e-06, 0.0530330490805908, 1, 0.265165245402954),      # grow, capped at 5x
             (6.318770e+00, 0.493273076774943, -1, 0.307045975536562),     # reject
             (1.027035e+00, 0.287682072451781, 0, 0.287682072451781),      # keep
             (3.458860e-01, 0.265165245402954, 1, 0.284841727047917)]
    for rmax, h, want, h_new in cases:
        hh = C.c_double(h)
        assert harness.mh_hadjust(rmax, 5, C.byref(hh)) == want
        assert abs(hh.value / h_new - 1) < 2e-6   # rmax is printed with 7 digits in the trace
