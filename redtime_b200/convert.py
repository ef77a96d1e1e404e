"""Downstream of the 17-column tables: what the reference's post-processing does,
working on the in-memory tables rtrg_run returns instead of re-reading text files.

* convert_pt  -- src/convert_pt.c:124-184: pick one output redshift, go from h-units to physical
  units (k [h/Mpc] -> k h [1/Mpc], P [(Mpc/h)^3] -> P / h^3) and rescale the non-linear P_cb by
  f_cb^2 = ((Omega_m - Omega_nu) / Omega_m)^2 (the delta_cb vs delta_m convention of the N-body
  spectra, src/convert_pt.c:51-55); also returns D divided by its value at the LAST wavenumber of
  the SAME redshift block (src/convert_pt.c:173: D0 = D_h[nk_pt*(output_z+1)-1]) and the neutrino
  spectrum, as process_PT_runs does.  The file-level drop-in for the tool itself is the C++
  executable redtime_b200/convertPt_b200 (csrc/convert_pt_main.cc), pinned byte for byte on the
  reference's own source in tests/test_convert_pt.py; this function is the same arithmetic on the
  in-memory tables.
* emulator_delta2 -- tests/emulator_comparison/test_models.py:20-40: the dimensionless
  Delta^2-like quantity P k^1.5 / (2 pi^2 h^3) the regression compares, without and with the
  neutrino correction f^2 = (1 - f_nu + beta_P)^2 rebuilt from columns 4 and 7.
* regression_metrics -- the two assertions of that test (max and 95 % quantile of the relative
  difference for k h < 0.1 / Mpc).
Column numbers below are 0-based indices of the default 17-column table (rt:1670-1737).
"""
import numpy as np

COL_K, COL_D, COL_PLIN_CB, COL_PLIN_NU, COL_PNL = 0, 1, 3, 6, 7


def convert_pt(table, h, omega_m, omega_nu, i_out=-1):
    """table: [n_out, nk, >=8] of one cosmology.  Returns dict(k, pk, D, pk_nu) in physical units."""
    t = np.asarray(table)
    f_cb = (omega_m - omega_nu) / omega_m
    blk = t[i_out]
    D0 = blk[-1, COL_D]                        # D at the last wavenumber of this block (convert_pt.c:173)
    return dict(k=blk[:, COL_K] * h,
                pk=blk[:, COL_PNL] / h ** 3 * f_cb ** 2,
                D=blk[:, COL_D] / D0,
                pk_nu=blk[:, COL_PLIN_NU] / h ** 3)


def emulator_delta2(table, h, omega_m=None, omega_nu=0.0, i_out=-1):
    """(k [1/Mpc], P_nl k^1.5 / (2 pi^2 h^3)) of the z block i_out; with omega_nu > 0 the neutrino
    correction of test_models.py:29-40 is applied."""
    blk = np.asarray(table)[i_out]
    k = blk[:, COL_K] * h
    norm = k ** 1.5 / h ** 3 / (2 * np.pi ** 2)
    nlin = blk[:, COL_PNL] * norm
    if omega_nu > 0:
        lin, lin_nu = blk[:, COL_PLIN_CB] * norm, blk[:, COL_PLIN_NU] * norm
        beta_p = np.sqrt(lin_nu / lin) * (omega_nu / omega_m)
        nlin = nlin * (1.0 - omega_nu / omega_m + beta_p) ** 2
    return k, nlin


def regression_metrics(k, nlin, nlin_target, kmax=0.1):
    """max and 95 % quantile of |nlin / target - 1| for k < kmax (test_models.py:86-88,156-159)."""
    m = k < kmax
    d = np.abs(nlin[m] / nlin_target[m] - 1.0)
    return float(d.max()), float(np.quantile(d, 0.95))
