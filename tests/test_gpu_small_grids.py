"""Small and odd grids: nk = 16, 32, 64 (the reference's nk is a compile-time constant; here it is
configuration).  At these sizes a row block has fewer alpha-side lags than a CTA has threads
(NV = 47 at nk = 16), so one CTA of k_bilinear holds up to six row blocks and almost every warp
straddles a block boundary -- the corner of the flattened (row block, lag) work distribution that
nk >= 128 barely touches.  There is no oracle build at these sizes; the check is a direct one: the
device quadratures against a numpy contraction of the SAME host tables (rtrg_table_T / _G, which
tests/test_tables.py pins on the reference's J_MFHB) with the spectra the device extrapolated:

    J_n(k_i; ab, cd) = kfac_n(i) sum_{j,l} a_ab[j] a_cd[l] T_n[(i-j) mod np][(i-l) mod np],  a[j] = P(q_j) q_j^2
    PZ_n(k_i; ab)    = dlnk/(2 pi^2) k_i^3 P_00(k_i) sum_m P_ab(q_m) G_n[i - m]

plus the invariances the larger grids are tested for (batching, k-sharding: bit-identical)."""
import threading

import numpy as np
import pytest

import redtime_b200 as rt

pytestmark = pytest.mark.gpu


def numpy_quadratures(nk, P3):
    g = rt.grid_info(nk)
    npad, nshift, jlo = g["np"], g["nshift"], g["jlo"]
    kpad = np.exp(g["lnk_pad_min"] + g["dlnk"] * np.arange(npad))
    a = P3 * kpad ** 2
    a[:, :jlo] = 0.0                      # samples below the window never enter (WP = 0 there anyway)
    J = np.zeros((14, 9, nk))
    for n in range(14):
        T, kf = rt.table_T(n, nk)
        for i in range(nk):
            ip = nshift + i
            idx = (ip - np.arange(npad)) % npad
            M = T[np.ix_(idx, idx)]       # M[j, l] = T[(ip-j) mod np][(ip-l) mod np]
            S = M @ a.T                   # [j, cd]
            J[n, :, i] = kf[ip] * (a @ S).reshape(9)   # [ab, cd]
    PZ = np.zeros((7, 9, nk))
    for n in range(7):
        G = rt.table_G(n, nk)
        for i in range(nk):
            ip = nshift + i
            m = np.arange(jlo, npad)
            conv = (P3[:, m] * G[ip - m + npad - 1]).sum(axis=1)          # [ab]
            base = g["dlnk"] / (2 * np.pi ** 2) * kpad[ip] ** 3 * P3[0, ip] * conv
            for ab in range(3):
                for cd in range(3):
                    PZ[n, 3 * ab + cd, i] = base[ab] * (P3[cd, ip] / (P3[0, ip] + 1e-100) if cd else 1.0)
    return J, PZ


@pytest.mark.parametrize("nk", [16, 32, 64])
def test_quadratures_against_numpy_contraction_of_the_host_tables(nk, example1_dir):
    h = rt.RedTimeB200(nk=nk)
    h.add_cosmology(rt.read_run_dir(example1_dir))
    h.prepare()
    y0, _ = h.initial_state()
    x = np.linspace(0.0, 1.0, nk)
    lnP = y0[:3 * nk].copy()
    lnP[nk:2 * nk] += 0.3 * np.sin(2.0 * x + 0.3) - 0.2      # three distinct spectra
    lnP[2 * nk:] += 0.5 * np.cos(2.5 * x) - 0.6
    P3 = h.extrap_P(lnP)
    J, PZ, J0, Jlo = h.integrals_raw(lnP)
    refJ, refPZ = numpy_quadratures(nk, P3)
    scale = np.max(np.abs(refJ), axis=2, keepdims=True)      # per (kernel, pair): the J's change sign in k
    got = np.concatenate([J.reshape(7, 9, nk), J0.reshape(7, 9, nk)])
    assert np.max(np.abs(got - refJ) / scale) < 1e-11
    assert np.max(np.abs(PZ.reshape(7, 9, nk) - refPZ) / np.max(np.abs(refPZ), axis=2, keepdims=True)) < 1e-12
    # the whole run works at this size and does not depend on the batch
    single, _, _, st = h.run()
    assert not st.any() and np.isfinite(single[0]).all()
    h.close()
    hb = rt.RedTimeB200(nk=nk)
    hb.add_cosmologies([rt.read_run_dir(example1_dir)] * 5)
    hb.prepare()
    tb, _, _, stb = hb.run()
    hb.close()
    assert not stb.any() and all(np.array_equal(t, single[0]) for t in tb)


def test_kshard_on_a_small_grid_is_bit_identical(example1_full_dir):
    nk, G = 32, 2            # 4 row blocks, 2 per rank; one CTA spans row blocks of BOTH ranks
    inp = rt.read_run_dir(example1_full_dir)
    s = rt.RedTimeB200(nk=nk)
    s.add_cosmology(inp)
    s.prepare()
    ref = s.run()[0][0]
    s.close()
    group = rt.LoopbackGroup(G)
    hs = [rt.RedTimeB200(nk=nk, k_shards=G, k_rank=r) for r in range(G)]
    res, errs = [None] * G, []

    def work(r):
        try:
            hs[r].add_cosmology(inp)
            hs[r].kshard_init_loopback(group)
            hs[r].prepare()
            res[r] = hs[r].run()[0][0]
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(r,)) for r in range(G)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    assert not errs, errs
    [x.close() for x in hs]
    group.close()
    for r in range(G):
        assert np.array_equal(res[r], ref), "rank %d" % r
