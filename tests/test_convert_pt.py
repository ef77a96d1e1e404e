"""convertPt_b200 (redtime_b200/csrc/convert_pt_main.cc) pinned on the reference's own
post-processing tool: oracle/_ref/convertPt is src/convert_pt.c compiled UNMODIFIED (gcc, as
src/Makefile:19-20 does) against the header-only spline shim oracle/gsl_shim/gsl/gsl_spline.h.
Same argv, same input files -> the k_M*.dat / pk_M*.dat files must be byte-identical for every
analysis step.  Also: the spline shim itself against scipy's natural cubic spline."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ORACLE_REF, ROOT

REF = os.path.join(ORACLE_REF, "convertPt")
OURS = os.path.join(ROOT, "redtime_b200", "convertPt_b200")
STEPS = [163, 189, 247, 300, 347, 401, 453, 499]


def write_inputs(d, n_models, nk=128, nz=33, seed=5):
    rng = np.random.default_rng(seed)
    with open(os.path.join(d, "models.dat"), "w") as f:
        f.write("# Cosmological models (1 per line)\n#\n# Columns\n#model  omega_m omega_b s8 h ns w0 wa omega_nu\n#\n")
        for m in range(n_models):
            om, onu = rng.uniform(0.12, 0.155), (0.0 if m % 2 == 0 else rng.uniform(0.001, 0.01))
            f.write("M%03d %.4f %.5f %.4f %.4f %.4f %.3f %.4f %.5f\n"
                    % (m + 1, om, rng.uniform(0.0215, 0.0235), rng.uniform(0.7, 0.9), rng.uniform(0.55, 0.85),
                       rng.uniform(0.85, 1.05), rng.uniform(-1.3, -0.7), rng.uniform(-1, 1), onu))
    k = np.exp(np.linspace(np.log(1e-3), np.log(1.0), nk))
    for m in range(n_models):
        with open(os.path.join(d, "redTime_M%03d.dat" % (m + 1)), "w") as f:
            f.write("#cosmological_parameters: opening parameter file: params_redTime.dat\n###main: eta_fin = 5.3\n")
            for iz in range(nz):
                f.write("### main: output at eta=%g\n" % (0.1 * iz))
                tab = np.abs(rng.standard_normal((nk, 17))) * 1e3 * np.exp(-3 * k[:, None]) + 1e-3
                tab[:, 0] = k
                tab[:, 1] = (0.2 + 0.8 * iz / nz) * (1 + 0.01 * np.sin(5 * k))
                for row in tab:
                    f.write("".join("%20.12g" % x for x in row) + "\n")
                f.write("\n\n")


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(OURS)), reason="convertPt binaries not built")
def test_convert_pt_files_byte_identical_to_the_reference_tool(tmp_path):
    da, db = str(tmp_path / "ref"), str(tmp_path / "ours")
    for d in (da, db):
        os.makedirs(d)
        write_inputs(d, 3)
    for step in STEPS + [1]:   # an unknown step number falls back to the first entry (convert_pt.c:150-153)
        # the reference writes ./junk.dat: run it inside its scratch directory
        subprocess.run([REF, "3", str(step), "128", os.path.join(da, "models.dat"), da], cwd=da, check=True)
        subprocess.run([OURS, "3", str(step), "128", os.path.join(db, "models.dat"), db], cwd=db, check=True)
        for m in (1, 2, 3):
            for stem in ("k", "pk"):
                name = os.path.join("STEP%d" % step, "%s_M%03d_no_interp_test.dat" % (stem, m))
                a, b = open(os.path.join(da, name), "rb").read(), open(os.path.join(db, name), "rb").read()
                assert a == b and len(a) > 128 * 8, (step, name)
    assert not os.path.exists(os.path.join(db, "junk.dat"))


@pytest.mark.skipif(not os.path.exists(OURS), reason="convertPt_b200 not built")
def test_convert_pt_rejects_what_the_reference_reads_uninitialised(tmp_path):
    d = str(tmp_path)
    write_inputs(d, 1, nz=8)     # step 499 wants block 32
    p = subprocess.run([OURS, "1", "499", "128", os.path.join(d, "models.dat"), d], stderr=subprocess.PIPE, text=True)
    assert p.returncode != 0 and "missing" in p.stderr
    assert subprocess.run([OURS, "1", "2"]).returncode != 0   # wrong argc


def test_spline_shim_is_the_natural_cubic_spline(tmp_path):
    """gsl_interp_cspline = natural cubic spline: the shim against scipy on the log-spaced k grid."""
    from scipy.interpolate import CubicSpline
    src = str(tmp_path / "s.c")
    open(src, "w").write('#include <gsl/gsl_spline.h>\n'
                         'double ev(const double*x,const double*y,int n,double q){gsl_spline*s=gsl_spline_alloc(gsl_interp_cspline,n);'
                         'gsl_interp_accel*a=gsl_interp_accel_alloc();gsl_spline_init(s,x,y,n);double r=gsl_spline_eval(s,q,a);'
                         'gsl_spline_free(s);gsl_interp_accel_free(a);return r;}\n')
    so = str(tmp_path / "s.so")
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "oracle", "gsl_shim"), src, "-o", so], check=True)
    lib = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.ev.restype = ctypes.c_double
    lib.ev.argtypes = [dp, dp, ctypes.c_int, ctypes.c_double]
    x = np.exp(np.linspace(np.log(1e-3), 0.0, 128))
    y = np.sin(3 * np.log(x)) + x
    cs = CubicSpline(x, y, bc_type="natural")
    for q in np.exp(np.linspace(np.log(1.1e-3), np.log(0.99), 57)):
        got = lib.ev(x.ctypes.data_as(dp), y.ctypes.data_as(dp), 128, float(q))
        assert abs(got - cs(q)) < 1e-11 * max(1.0, abs(cs(q)))
