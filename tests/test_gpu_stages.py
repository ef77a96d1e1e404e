"""Stage-level parity of the CUDA path (through the C-ABI) against the ORACLE = the
unmodified reference sources built against the mini-GSL shim (committed fixtures
tests/golden/example1_stage_*.npz, made by tests/golden/make_golden.py).

Tolerances (floating point, FP64 everywhere):
  * look-ups / linear theory (columns 1-7 material):  <= 1e-9 relative (north star: 1e-6)
  * raw quadratures J, PZ, Jn0:                       <= 1e-10 relative
  * assembled A, R, PT, PMR and the RHS:              |diff| <= 1e-9 * sum|terms|-type floor,
    i.e. relative 1e-6 away from the low-k cancellation rows
"""
import numpy as np
import pytest

import redtime_b200 as rt

pytestmark = pytest.mark.gpu
NK = 128
JU = [8, 9, 10, 11, 12, 13, 14, 15, 56, 57, 59, 60, 61, 63]


def relerr(a, b, floor=0.0):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor)))


@pytest.fixture(scope="module")
def handle_1loop(example1_dir):
    h = rt.RedTimeB200()
    h.add_cosmology(rt.read_run_dir(example1_dir))
    h.prepare()
    yield h
    h.close()


@pytest.fixture(scope="module")
def handle_full(example1_full_dir):
    h = rt.RedTimeB200()
    h.add_cosmology(rt.read_run_dir(example1_full_dir))
    h.prepare()
    yield h
    h.close()


def test_linear_theory_lookups(handle_1loop, stage_golden_1loop):
    g, h = stage_golden_1loop, handle_1loop
    k = g["lin_k"]
    for iz, z in enumerate(g["lin_z"]):
        D, dD = h.D_dD(z, k)
        assert relerr(D, g["lin_D"][iz]) < 1e-9, ("D", z)
        assert relerr(dD, g["lin_dD"][iz]) < 1e-9, ("dD", z)
        assert relerr(h.Beta_P(1.0 / (1.0 + z), k), g["lin_beta"][iz]) < 1e-12, ("beta", z)
        for which, name in enumerate(("lin_P", "lin_Pcb", "lin_Pnu")):
            assert relerr(h.Plin(which, z, k), g[name][iz]) < 1e-8, (name, z)


def test_lookup_range_errors(handle_1loop):
    # the reference aborts (hdr:528-531, 646-649); the C-ABI returns RTRG_ERANGE
    with pytest.raises(rt.RtrgError) as e:
        handle_1loop.Beta_P(1.01, np.array([0.1]))
    assert e.value.code == -4
    with pytest.raises(rt.RtrgError) as e:
        handle_1loop.D_dD(2000.0, np.array([0.1]))
    assert e.value.code == -4
    assert relerr(handle_1loop.Beta_P(1.0005, np.array([0.1])), handle_1loop.Beta_P(1.0, np.array([0.1]))) == 0


def test_initial_state_and_normalisation(handle_1loop, stage_golden_1loop):
    g = stage_golden_1loop
    y, scal = handle_1loop.initial_state()
    assert relerr(y[:3 * NK], g["y0"][:3 * NK]) < 1e-9
    assert not y[3 * NK:].any()
    # sigma_v^2(z=0) pins the QAG-61 subdivision
    assert abs(scal[1] / g["sigmaV2"][-1] - 1) < 1e-9


def test_extrapolated_spectra(handle_1loop, stage_golden_1loop):
    g = stage_golden_1loop
    for tag in ("y0", "yp"):
        P3 = handle_1loop.extrap_P(g[tag][:3 * NK])
        ref = g["P3_" + tag]
        assert np.array_equal(P3 == 0, ref == 0)
        m = ref != 0
        assert relerr(P3[m], ref[m]) < 1e-13


def test_raw_quadratures(handle_1loop, stage_golden_1loop):
    g = stage_golden_1loop
    J, PZ, J0, Jlo = handle_1loop.integrals_raw(g["yp"][:3 * NK])
    assert relerr(PZ, g["PZ_yp"]) < 1e-12
    # kernel n=1 is the regularised J_{2,-2,0}: noisiest in the reference itself
    for n in range(7):
        s = slice(9 * n, 9 * n + 9)
        assert relerr(J[s], g["J_yp"][s]) < (1e-10 if n == 1 else 5e-12), ("J", n)
        assert relerr(J0[s], g["Jn0_yp"][s]) < 5e-12, ("Jn0", n)
    assert abs(Jlo / g["Jlo_yp"] - 1) < 5e-12


def assembled_floor(g, tag):
    """sum over terms of |coef k^p X|: the magnitude the cancelling sums are made of."""
    row, src, idx, kpw, cf = rt.assembly_terms()
    k = g["k"]
    vals = [g["J_" + tag], g["PZ_" + tag], g["Jn0_" + tag]]
    mag = np.zeros((55, NK))
    for r, s, i, p, c in zip(row, src, idx, kpw, cf):
        v = vals[s][i] if s < 3 else g["Jlo_" + tag]
        mag[r] += np.abs(c * k ** float(p) * v)
    return mag


def test_assembled_integrals(handle_1loop, stage_golden_1loop):
    g = stage_golden_1loop
    A, R, PT, PMR = handle_1loop.integrals_full(g["yp"][:3 * NK])
    mag = assembled_floor(g, "yp")
    got = np.concatenate([A[JU], R, PT, PMR])
    ref = np.concatenate([g["A_yp"][JU], g["R_yp"], g["PT_yp"], g["PMR_yp"]])
    assert np.all(np.abs(got - ref) <= 2e-11 * mag)
    # ... and at every k: 1e-7 relative + 5 x the reference's own response to a 1-ulp change of
    # ln P (tests/golden/floor_stage.npz, made by tests/golden/make_floor.py)
    import os
    from conftest import GOLDEN, FLOOR_C
    fl = np.load(os.path.join(GOLDEN, "floor_stage.npz"))
    floor = np.concatenate([fl["nk128_floor_A"][JU], fl["nk128_floor_R"], fl["nk128_floor_PT"], fl["nk128_floor_PMR"]])
    allowed = 1e-7 * np.abs(ref) + FLOOR_C * np.max(floor, axis=0, keepdims=True) + 1e-300
    assert np.all(np.abs(got - ref) <= allowed), float(np.max(np.abs(got - ref) / allowed))
    for dst, s in ((16, 8), (18, 9), (17, 10), (19, 11), (20, 12), (22, 13), (21, 14), (23, 15),
                   (58, 57), (62, 61)):
        assert np.array_equal(A[dst], A[s])


@pytest.mark.parametrize("mode", ["1loop", "full"])
def test_rhs(mode, handle_1loop, handle_full, stage_golden_1loop, stage_golden_full):
    g, h = (stage_golden_1loop, handle_1loop) if mode == "1loop" else (stage_golden_full, handle_full)
    y = g["yp"]
    for eta, ref in zip(g["rhs_eta"], g["rhs_dy"]):
        dy = h.derivatives(eta, y)
        d, r = dy.reshape(41, NK), ref.reshape(41, NK)
        assert relerr(d[:3], r[:3]) < 1e-9, ("dlnP", eta)
        # I and Q rows: the sources A, R are cancelling sums of the J integrals (FFT round-off
        # floor of the reference itself); compare against the row's own scale.
        # Measured on B200: <= 9e-7 (1-loop), <= 3e-7 (full).
        s = np.maximum(np.abs(r), np.max(np.abs(r), axis=1, keepdims=True) * 1e-6)
        assert np.max(np.abs(d[3:] - r[3:]) / s[3:]) < 3e-6, ("dI/dQ", eta)
