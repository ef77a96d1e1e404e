"""Post-processing helpers (redtime_b200/convert.py) against a direct transcription of what
src/convert_pt.c and tests/emulator_comparison/test_models.py compute, on the reference's own
example table."""
import numpy as np

from redtime_b200 import convert

H, OM, ONU = 0.73418, 0.286233679143621, 0.00576437405571056  # examples/1_redTime/params_redTime.dat


def test_convert_pt_matches_reference_arithmetic(golden_example1):
    _, arr = golden_example1
    t = arr.reshape(7, 128, 17)
    out = convert.convert_pt(t, H, OM, ONU, i_out=3)
    f_cb = (OM - ONU) / OM
    # src/convert_pt.c:150-153,176-181 and :54
    assert np.allclose(out["k"], t[3, :, 0] * H, rtol=1e-15)
    assert np.allclose(out["pk"], t[3, :, 7] / H ** 3 * f_cb * f_cb, rtol=1e-15)
    assert np.allclose(out["pk_nu"], t[3, :, 6] / H ** 3, rtol=1e-15)
    assert np.allclose(out["D"], t[3, :, 1] / t[3, 127, 1], rtol=1e-15)   # :173, same redshift block


def test_convert_pt_python_equals_the_executable(golden_example1, tmp_path):
    """convert.convert_pt (in memory) against convertPt_b200 (files; itself byte-identical to the
    reference tool, tests/test_convert_pt.py) on a 33-block table."""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "redtime_b200", "convertPt_b200")
    if not os.path.exists(exe):
        import pytest
        pytest.skip("convertPt_b200 not built")
    _, arr = golden_example1
    t7 = arr.reshape(7, 128, 17)
    t = np.concatenate([t7] * 5)[:33]          # the HACC output list has 33 redshifts (convert_pt.c:134)
    d = str(tmp_path)
    with open(os.path.join(d, "models.dat"), "w") as f:
        f.write("#\n" * 5 + "M001 %.17g 0.0226 0.8 %.17g 0.96 -1 0 %.17g\n" % (OM * H * H, H, ONU * H * H))
    with open(os.path.join(d, "redTime_M001.dat"), "w") as f:
        for blk in t:
            f.write("### main: output\n")
            for row in blk:
                f.write("".join("%20.12g" % x for x in row) + "\n")
            f.write("\n\n")
    subprocess.run([exe, "1", "300", "128", os.path.join(d, "models.dat"), d], check=True)   # step 300 -> block 18
    out = convert.convert_pt(t, H, OM * H * H, ONU * H * H, i_out=18)
    k = np.array(open(os.path.join(d, "STEP300", "k_M001_no_interp_test.dat")).read().split(), dtype=float)
    pk = np.array(open(os.path.join(d, "STEP300", "pk_M001_no_interp_test.dat")).read().split(), dtype=float)
    assert np.allclose(k, out["k"], atol=6e-7, rtol=0) and np.allclose(pk, out["pk"], atol=6e-7, rtol=0)  # "%lf": 6 decimals


def test_emulator_regression_quantities(golden_example1):
    _, arr = golden_example1
    t = arr.reshape(7, 128, 17)
    rf = arr  # what np.loadtxt of the file gives; the test uses the last 128 rows
    k, nlin = convert.emulator_delta2(t, H, OM, ONU)
    kk = rf[-128:, 0] * H
    lin = rf[-128:, 3] / H ** 3 / (2 * np.pi ** 2) * kk ** 1.5
    nl = rf[-128:, 7] / H ** 3 / (2 * np.pi ** 2) * kk ** 1.5
    lin_nu = rf[-128:, 6] / H ** 3 / (2 * np.pi ** 2) * kk ** 1.5
    f = 1.0 - ONU / OM + np.sqrt(lin_nu / lin) * (ONU / OM)
    assert np.allclose(nlin, nl * f ** 2, rtol=1e-14) and np.allclose(k, kk, rtol=1e-15)
    k0, nl0 = convert.emulator_delta2(t, H)
    assert np.allclose(nl0, nl, rtol=1e-15)
    mx, q95 = convert.regression_metrics(k, nlin, nlin * (1 + 1e-4 * np.sin(np.arange(128))))
    assert 0 < q95 <= mx < 1.1e-4
