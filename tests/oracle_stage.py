#!/usr/bin/env python3
"""Run the stage-level ORACLE (oracle/_ref/libredtime_stage*.so = the unmodified reference
sources + mini-GSL shim) in this process and dump its stage outputs to an .npz.

TEST INFRASTRUCTURE.  Must run as its own process with the run directory as CWD because the
reference constructs its global `cosmological_parameters C("params_redTime.dat")` at load
time and keeps function-local statics (one process = one cosmology).

usage: oracle_stage.py <run_dir> <out.npz> [--lib PATH] [--seed N] [--light]
"""
import argparse
import ctypes as C
import os
import sys
import numpy as np

SPECS_J = [(0, 0, 0), (2, -2, 0), (1, -1, 1), (0, 0, 2), (2, -2, 2), (1, -1, 3), (0, 0, 4)]
SPECS_0 = [(0, 2, 0), (0, 2, 2), (0, 2, 4), (2, 2, 0), (2, 2, 2), (2, 2, 4), (2, 2, 6)]
ZN = [0, 1, -1, 3, -3, 5, -5]


def perturbed_state(y0, nk, seed):
    """Deterministic non-trivial state: distinct P_ab and non-zero I, Q."""
    rng = np.random.default_rng(seed)
    y = y0.copy()
    x = np.linspace(0.0, 1.0, nk)
    y[0:nk] += 0.05 * np.sin(3.0 * x)
    y[nk:2 * nk] += 0.3 * np.sin(2.0 * x + 0.3) - 0.2
    y[2 * nk:3 * nk] += 0.5 * np.cos(2.5 * x) - 0.6
    y[3 * nk:] = 1e-3 * rng.standard_normal(38 * nk) * np.exp(y0[0:nk].mean())
    return y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("run_dir")
    ap.add_argument("out")
    ap.add_argument("--lib", default=None)
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--light", action="store_true", help="skip the raw J/PZ arrays")
    ap.add_argument("--floor", type=int, default=0, metavar="N",
                    help="only measure the round-off floor of the assembled integrals (SURVEY H2): N "
                         "evaluations of compute_*_full on yp with ln P moved by a random -1/0/+1 ulp; "
                         "writes floor_{A,R,PT,PMR} = max |difference to the unperturbed result|")
    a = ap.parse_args()
    here = os.path.dirname(os.path.abspath(__file__))
    lib_path = a.lib or os.path.join(here, "..", "oracle", "_ref", "libredtime_stage.so")
    lib_path = os.path.abspath(lib_path)
    out_path = os.path.abspath(a.out)
    os.chdir(a.run_dir)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)  # the reference prints its banner to stdout at load time
    ref = C.CDLL(lib_path)
    dp = C.POINTER(C.c_double)

    def P(x):
        return x.ctypes.data_as(dp)

    nk, npad, nU = ref.ref_nk(), ref.ref_np(), ref.ref_nU()
    for f in ("ref_dlnk", "ref_z_in", "ref_z_out", "ref_eta_out", "ref_sigmaV2", "ref_H_H0",
              "ref_WP", "ref_WC", "ref_Omega", "ref_Pbisj", "ref_H2_H02", "ref_dlnH_dlna"):
        getattr(ref, f).restype = C.c_double
    ref.ref_sigmaV2.argtypes = [C.c_double]
    ref.ref_H_H0.argtypes = [C.c_double]
    k = np.zeros(nk)
    ref.ref_init.argtypes = [dp]
    ref.ref_init(P(k))
    y0 = np.zeros(nU * nk)
    ref.ref_initial_y.argtypes = [dp]
    ref.ref_initial_y(P(y0))
    yp = perturbed_state(y0, nk, a.seed)
    res = dict(k=k, y0=y0, yp=yp, nk=nk, np=npad, z_in=ref.ref_z_in(),
               switches=np.array([ref.ref_switch(i) for i in range(4)]),
               z_out=np.array([ref.ref_z_out(i) for i in range(ref.ref_n_out())]),
               eta_out=np.array([ref.ref_eta_out(i) for i in range(ref.ref_n_out())]))

    ref.ref_extrap_P.argtypes = [dp, dp]
    ref.ref_compute_full.argtypes = [C.c_double, dp, dp, dp, dp, dp]
    ref.ref_derivatives.argtypes = [C.c_double, dp, dp]
    if a.floor > 0:
        # the assembled integrals are cancelling sums of FFT-built J's (redTime.cc:1182-1184);
        # their round-off floor is what the reference itself moves by when its input moves by one ulp
        def full(y):
            o = [np.zeros(64 * nk), np.zeros(24 * nk), np.zeros(9 * nk), np.zeros(8 * nk)]
            ref.ref_compute_full(0.0, P(y), *[P(x) for x in o])
            return [x.reshape(-1, nk) for x in o]
        base = full(yp)
        fl = [np.zeros_like(x) for x in base]
        rng = np.random.default_rng(a.seed + 1)
        for _ in range(a.floor):
            y = yp.copy()
            d = rng.integers(-1, 2, size=3 * nk)
            y[:3 * nk] = np.where(d > 0, np.nextafter(y[:3 * nk], np.inf),
                                  np.where(d < 0, np.nextafter(y[:3 * nk], -np.inf), y[:3 * nk]))
            for f, x, b in zip(fl, full(y), base):
                np.maximum(f, np.abs(x - b), out=f)
        os.dup2(saved, 1)
        np.savez_compressed(out_path, k=k, floor_A=fl[0], floor_R=fl[1], floor_PT=fl[2], floor_PMR=fl[3])
        return 0
    for tag, y in (("y0", y0), ("yp", yp)):
        P3 = np.zeros(3 * npad)
        ref.ref_extrap_P(P(y), P(P3))
        res["P3_" + tag] = P3.reshape(3, npad)
        A, R, PT, PMR = np.zeros(64 * nk), np.zeros(24 * nk), np.zeros(9 * nk), np.zeros(8 * nk)
        ref.ref_compute_full(0.0, P(y), P(A), P(R), P(PT), P(PMR))
        res["A_" + tag], res["R_" + tag] = A.reshape(64, nk), R.reshape(24, nk)
        res["PT_" + tag], res["PMR_" + tag] = PT.reshape(9, nk), PMR.reshape(8, nk)
    # raw bilinear / P13 terms of the perturbed state (for the round-off floor)
    if not a.light:
        ref.ref_J_MFHB.argtypes = [C.c_int] * 3 + [dp, dp, dp]
        ref.ref_PZ_reg.argtypes = [C.c_int, dp, dp, dp]
        P3 = res["P3_yp"]
        J, J0, PZ = np.zeros((63, npad)), np.zeros((63, npad)), np.zeros((63, npad))
        for iJ in range(63):
            n, ab, cd = iJ // 9, (iJ % 9) // 3, iJ % 3
            pa, pb, o = np.ascontiguousarray(P3[ab]), np.ascontiguousarray(P3[cd]), np.zeros(npad)
            ref.ref_J_MFHB(*SPECS_J[n], P(pa), P(pb), P(o))
            J[iJ] = o
            o = np.zeros(npad)
            ref.ref_J_MFHB(*SPECS_0[n], P(pa), P(pb), P(o))
            J0[iJ] = o
        for iJ in range(0, 63, 3):
            n, ab = iJ // 9, (iJ % 9) // 3
            pa, pb, o = np.ascontiguousarray(P3[ab]), np.ascontiguousarray(P3[0]), np.zeros(npad)
            ref.ref_PZ_reg(ZN[n], P(pa), P(pb), P(o))
            PZ[iJ] = o
            PZ[iJ + 1] = o * P3[1] / (P3[0] + 1e-100)
            PZ[iJ + 2] = o * P3[2] / (P3[0] + 1e-100)
        nshift = (npad - nk) // 2
        res["J_yp"], res["Jn0_yp"], res["PZ_yp"] = J[:, nshift:nshift + nk], J0[:, nshift:nshift + nk], PZ[:, nshift:nshift + nk]
        res["Jlo_yp"] = J[0, nshift - nk // 2]
    # right-hand side at a few times
    etas = np.array([0.0, 1.3, 3.9, float(res["eta_out"][-1])])
    dys = []
    for eta in etas:
        dy = np.zeros(nU * nk)
        ref.ref_derivatives(float(eta), P(yp), P(dy))
        dys.append(dy)
    res["rhs_eta"], res["rhs_dy"] = etas, np.array(dys)
    # linear theory
    ref.ref_D_dD.argtypes = [C.c_double, dp, C.c_int, dp, dp]
    ref.ref_Beta_P.argtypes = [C.c_double, dp, C.c_int, dp]
    for nm in ("ref_Plin", "ref_Plin_cb", "ref_Plin_nu"):
        getattr(ref, nm).argtypes = [C.c_double, dp, C.c_int, dp]
    kx = np.concatenate([k, np.array([1.0e-4, 2.0e-4, 3.3e-3, 0.57, 2.0, 8.0, 20.0])])
    zs = np.array([float(res["z_in"]), 10.0, 5.0, 2.02, 1.0, 0.5, 0.0])
    D, dD, B, Pl, Pcb, Pnu = ([] for _ in range(6))
    for z in zs:
        d1, d2, b, p0, p1, p2 = (np.zeros(kx.size) for _ in range(6))
        ref.ref_D_dD(float(z), P(kx), kx.size, P(d1), P(d2))
        ref.ref_Beta_P(1.0 / (1.0 + float(z)), P(kx), kx.size, P(b))
        ref.ref_Plin(float(z), P(kx), kx.size, P(p0))
        ref.ref_Plin_cb(float(z), P(kx), kx.size, P(p1))
        ref.ref_Plin_nu(float(z), P(kx), kx.size, P(p2))
        D.append(d1), dD.append(d2), B.append(b), Pl.append(p0), Pcb.append(p1), Pnu.append(p2)
    res.update(lin_k=kx, lin_z=zs, lin_D=np.array(D), lin_dD=np.array(dD), lin_beta=np.array(B),
               lin_P=np.array(Pl), lin_Pcb=np.array(Pcb), lin_Pnu=np.array(Pnu))
    res["sigmaV2"] = np.array([ref.ref_sigmaV2(float(z)) for z in zs])
    res["H_H0"] = np.array([ref.ref_H_H0(1.0 / (1.0 + float(z))) for z in zs])
    os.dup2(saved, 1)
    np.savez_compressed(out_path, **res)
    return 0


if __name__ == "__main__":
    sys.exit(main())
