#!/usr/bin/env python3
"""Print stage-by-stage and end-to-end parity numbers of the CUDA path vs the committed
oracle fixtures (diagnostic twin of tests/test_gpu_*.py; nothing asserted)."""
import gzip
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest  # noqa: E402
import redtime_b200 as rt  # noqa: E402

NK = 128
JU = [8, 9, 10, 11, 12, 13, 14, 15, 56, 57, 59, 60, 61, 63]


def rel(a, b, floor=1e-300):
    return float(np.max(np.abs(np.asarray(a) - b) / (np.abs(b) + floor)))


def stage(tag, d, g):
    print("==== stage parity,", tag)
    t0 = time.time()
    h = rt.RedTimeB200()
    t1 = time.time()
    h.add_cosmology(rt.read_run_dir(d))
    t2 = time.time()
    h.prepare()
    t3 = time.time()
    print("create %.2fs read %.2fs prepare %.2fs" % (t1 - t0, t2 - t1, t3 - t2))
    k = g["lin_k"]
    for iz, z in enumerate(g["lin_z"]):
        D, dD = h.D_dD(z, k)
        e = [rel(D, g["lin_D"][iz]), rel(dD, g["lin_dD"][iz]), rel(h.Beta_P(1 / (1 + z), k), g["lin_beta"][iz])]
        e += [rel(h.Plin(w, z, k), g[n][iz]) for w, n in enumerate(("lin_P", "lin_Pcb", "lin_Pnu"))]
        print("z=%-6g D %.1e dD %.1e beta %.1e P %.1e Pcb %.1e Pnu %.1e" % (z, *e))
    y, scal = h.initial_state()
    print("y0 rel", rel(y[:3 * NK], g["y0"][:3 * NK]), "sigv2 rel", scal[1] / g["sigmaV2"][-1] - 1, "Norm", scal[0])
    for t in ("y0", "yp"):
        P3 = h.extrap_P(g[t][:3 * NK])
        m = g["P3_" + t] != 0
        print("extrap", t, "support equal", np.array_equal(P3 == 0, ~m), "rel", rel(P3[m], g["P3_" + t][m]))
    J, PZ, J0, Jlo = h.integrals_raw(g["yp"][:3 * NK])
    for n in range(7):
        s = slice(9 * n, 9 * n + 9)
        print("n=%d J %.1e Jn0 %.1e PZ %.1e" % (n, rel(J[s], g["J_yp"][s]), rel(J0[s], g["Jn0_yp"][s]), rel(PZ[s], g["PZ_yp"][s])))
    print("Jlo", Jlo / g["Jlo_yp"] - 1)
    A, R, PT, PMR = h.integrals_full(g["yp"][:3 * NK])
    got = np.concatenate([A[JU], R, PT, PMR])
    ref = np.concatenate([g["A_yp"][JU], g["R_yp"], g["PT_yp"], g["PMR_yp"]])
    hi = g["k"] > 5.7e-3
    print("assembled rel (k>5.7e-3)", rel(got[:, hi], ref[:, hi]), " all k", rel(got, ref))
    for eta, r in zip(g["rhs_eta"], g["rhs_dy"]):
        dy = h.derivatives(eta, g["yp"]).reshape(41, NK)
        r = r.reshape(41, NK)
        s = np.maximum(np.abs(r), np.max(np.abs(r), axis=1, keepdims=True) * 1e-6)
        print("rhs eta=%.3f dlnP %.1e  dI,dQ(k>5.7e-3) %.1e  all %.1e" % (
            eta, rel(dy[:3], r[:3]), np.max(np.abs(dy[3:, hi] - r[3:, hi]) / s[3:, hi]), np.max(np.abs(dy[3:] - r[3:]) / s[3:])))
    h.close()


def e2e(tag, d, refs):
    print("==== end to end,", tag)
    h = rt.RedTimeB200()
    h.add_cosmology(rt.read_run_dir(d))
    h.prepare()
    t0 = time.time()
    tables, hdr, hdr0, status = h.run(raise_on_ode_failure=False)
    t1 = time.time()
    print("run %.3fs status %s counters %s launches %d" % (t1 - t0, status, h.counters(0), h.launch_count()))
    tab = tables[0]
    for name, ref in refs:
        ref = ref.reshape(tab.shape)
        e = np.max(np.abs(tab - ref) / (np.abs(ref) + 1e-300), axis=(0, 1))
        hi = ref[0, :, 0] > 5.7e-3
        ehi = np.max(np.abs(tab[:, hi] - ref[:, hi]) / (np.abs(ref[:, hi]) + 1e-300), axis=(0, 1))
        print("vs", name)
        print("  col err all k :", " ".join("%.1e" % x for x in e))
        print("  col err k>5.7e-3:", " ".join("%.1e" % x for x in ehi))
    print("hdr0", hdr0[0], "hdr[0]", hdr[0, 0])
    h.close()
    return tab


def main():
    with tempfile.TemporaryDirectory() as tmp:
        d1 = conftest.make_example1_dir(os.path.join(tmp, "a"))
        d2 = conftest.make_example1_dir(os.path.join(tmp, "b"), switches=[1, 0, 1, 1])
        g1 = dict(np.load(os.path.join(conftest.GOLDEN, "example1_stage_1loop.npz")))
        g2 = dict(np.load(os.path.join(conftest.GOLDEN, "example1_stage_full.npz")))
        stage("1loop", d1, g1)
        stage("full", d2, g2)

        def load(p):
            with gzip.open(os.path.join(conftest.GOLDEN, p), "rt") as f:
                return conftest.parse_tables(f.read())[1]
        gold = load("example1/example_redTime_result.dat.gz")
        o1, o2 = load("example1_oracle_1loop.dat.gz"), load("example1_oracle_full.dat.gz")
        t1 = e2e("1loop", d1, [("reference golden (genuine GSL)", gold), ("oracle", o1)])
        t2 = e2e("full TRG", d2, [("oracle", o2)])
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", "e2e_tables.npz"), t1=t1, t2=t2)
        print("FP64 pipe peak TFLOP/s: DMMA loop", rt.dmma_peak_tflops(0, 0.5), " DFMA loop", rt.dfma_peak_tflops(0, 0.5))


if __name__ == "__main__":
    main()
