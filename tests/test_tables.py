"""Host-side, cosmology-independent tables of the C-ABI library against the ORACLE's stage
outputs (committed fixtures made from the unmodified reference, tests/golden/make_golden.py).
No GPU needed: these calls only build tables on the host."""
import numpy as np
import pytest

import redtime_b200 as rt

NK, NPAD = 128, 512
JU = [8, 9, 10, 11, 12, 13, 14, 15, 56, 57, 59, 60, 61, 63]  # redTime.cc:157


@pytest.fixture(scope="module")
def grid():
    return rt.grid_info(NK, 1e-3, 1.0)


def test_grid_constants(grid):
    # redTime.cc:90-110 and SURVEY B.1 / Q2: Lo=184 -> first non-zero window sample 185
    assert grid["np"] == 512 and grid["nshift"] == 192 and grid["nloMR"] == 128
    assert grid["jlo"] == 185 and grid["nsup"] == 327
    assert abs(grid["dlnk"] - np.log(1e3) / 127) < 1e-16


def test_windows_match_oracle_extrapolation(grid, stage_golden_1loop):
    WP, WC = rt.table_windows(NK)
    P3 = stage_golden_1loop["P3_yp"]
    assert np.all((WP == 0) == (P3[0] == 0))  # same support as the reference's P*WP
    assert WP[256] == 1.0 and WP[511] == 1.0 and 0 < WP[192] < 0.01  # Q2: no right taper
    assert WC[0] == 1 and WC[NPAD // 8] == 1 and WC[NPAD // 2] == 0 and 0 < WC[NPAD // 4] < 1


@pytest.mark.parametrize("n", [0, 1, 2, 4, 6, 7, 10, 13])
def test_T_bilinear_form_reproduces_J_MFHB(n, grid, stage_golden_1loop):
    """J_n(k_i;A,B) = kfac_i sum_jl a_j b_l T_n[(i-j)%np][(i-l)%np] vs redTime.cc:411-597."""
    g = stage_golden_1loop
    T, kf = rt.table_T(n, NK)
    kpad = np.exp(grid["lnk_pad_min"] + grid["dlnk"] * np.arange(NPAD))
    a = g["P3_yp"] * kpad ** 2
    ref = g["J_yp"] if n < 7 else g["Jn0_yp"]
    nn = n % 7
    worst = 0.0
    for ii in range(0, NK, 9):
        i = grid["nshift"] + ii
        idx = (i - np.arange(NPAD)) % NPAD
        Tw = T[np.ix_(idx, idx)]
        for pair in (0, 1, 5, 8):
            val = kf[i] * (a[pair // 3] @ Tw @ a[pair % 3])
            worst = max(worst, abs(val / ref[9 * nn + pair, ii] - 1))
    # the regularised J_{2,-2,0} (n=1) is the noisiest in the reference itself
    assert worst < (5e-11 if n == 1 else 2e-12), worst


def test_G_kernels_reproduce_PZ_reg(grid, stage_golden_1loop):
    """redTime.cc:689-727 as a 1-D log-q quadrature with weights G_n."""
    g = stage_golden_1loop
    P3 = g["P3_yp"]
    kpad = np.exp(grid["lnk_pad_min"] + grid["dlnk"] * np.arange(NPAD))
    pre = grid["dlnk"] / (2 * np.pi ** 2)
    for n in range(7):
        G = rt.table_G(n, NK)
        for ab in range(3):
            for ii in range(0, NK, 5):
                i = grid["nshift"] + ii
                v = pre * kpad[i] ** 3 * P3[0, i] * np.sum(P3[ab] * G[i - np.arange(NPAD) + NPAD - 1])
                assert abs(v / g["PZ_yp"][9 * n + 3 * ab, ii] - 1) < 1e-13


def assemble(g, tag="yp"):
    row, src, idx, kpw, cf = rt.assembly_terms()
    k = g["k"]
    vals = [g["J_" + tag], g["PZ_" + tag], g["Jn0_" + tag]]
    out, mag = np.zeros((55, NK)), np.zeros((55, NK))
    for r, s, i, p, c in zip(row, src, idx, kpw, cf):
        v = vals[s][i] if s < 3 else g["Jlo_" + tag]
        t = c * k ** float(p) * v
        out[r] += t
        mag[r] += np.abs(t)
    return out, mag


def test_assembly_table_reproduces_reference_outputs(stage_golden_1loop):
    """The reference's own J/PZ/Jn0 pushed through the coefficient table must give its
    A (14 unique), R (24), PTjm (9), PMRn (8) (redTime.cc:813-1279) up to the summation-order
    round-off of the cancelling terms: |diff| <= 64 ulp of sum |terms|."""
    g = stage_golden_1loop
    out, mag = assemble(g)
    ref = np.concatenate([g["A_yp"][JU], g["R_yp"], g["PT_yp"], g["PMR_yp"]])
    assert np.all(np.abs(out - ref) <= 64 * 2.3e-16 * mag + 1e-300)
    # and the symmetric copies of A (redTime.cc:968-978)
    A = g["A_yp"]
    for dst, s in ((16, 8), (18, 9), (17, 10), (19, 11), (20, 12), (22, 13), (21, 14), (23, 15),
                   (58, 57), (62, 61)):
        assert np.array_equal(A[dst], A[s])
    used = set(JU) | {16, 18, 17, 19, 20, 22, 21, 23, 58, 62}
    for j in range(64):
        if j not in used:
            assert not A[j].any()


def test_extrap_stencil_reproduces_Pab_times_window(stage_golden_1loop, tmp_path):
    """Pab(a,b,k_pad,y) * WP (rt:181-232, 772-778) as the 4-point stencil the device kernel
    k_extrap applies: ln P = sum_j w_j lnP[n0+j] + (n_s-3) dx, on the oracle's states y0 / yp."""
    from conftest import make_example1_dir
    ns = rt.read_run_dir(make_example1_dir(str(tmp_path / "M")))["params"][0]
    g = stage_golden_1loop
    n0, w, dx = rt.table_extrap(NK)
    WP, _ = rt.table_windows(NK)
    assert n0.min() == 0 and n0.max() == NK - 4
    assert np.allclose(w.sum(1), 1.0, rtol=0, atol=1e-12)     # interpolation weights
    assert np.all(dx[:192 + NK] == 0) and np.all(dx[192 + NK:] > 0)  # power-law tail above kmax only
    for tag in ("y0", "yp"):
        y = g[tag][:3 * NK].reshape(3, NK)
        lnP = (w[None] * y[:, n0[:, None] + np.arange(4)[None]]).sum(-1) + (ns - 3) * dx[None]
        P = np.where(WP > 0, np.exp(lnP) * WP, 0.0)
        ref = g["P3_" + tag]
        assert np.all((P == 0) == (ref == 0))
        m = ref != 0
        assert np.max(np.abs(P[m] / ref[m] - 1)) < 1e-13
