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
    assert np.allclose(out["D"], t[3, :, 1] / t[6, 127, 1], rtol=1e-15)


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
