#!/usr/bin/env python3
"""Micro-benchmark of the integral kernels: B cosmologies, `reps` evaluations per output-group
mix.  usage: bench_integrals.py [B] [reps] [nk]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import redtime_b200 as rt  # noqa: E402
from redtime_b200 import workload as wl  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
nk = int(sys.argv[3]) if len(sys.argv) > 3 else 128
base = wl.load_example1(16)
h = rt.RedTimeB200(nk=nk)
h.add_cosmologies(wl.make_cosmologies(B, base))
h.prepare()
g = rt.grid_info(nk)
peak = rt.dmma_peak_tflops(0, 0.3)
flop_set = nk * (2.0 * g["nsup"] ** 2 + 6.0 * g["nsup"])
print("B=%d nk=%d  measured FP64 pipe peak (DMMA loop) %.2f TFLOP/s" % (B, nk, peak))
row, src, idx, kpw, cf = rt.assembly_terms()


def sets(groups, ident):
    need, need_ab = {}, {}
    for r, s_, i in zip(row, src, idx):
        if s_ not in (0, 2):
            continue
        gi = 0 if r < 14 else 1 if r < 38 else 2 if r < 47 else 3
        if groups & (1 << gi):
            n = i // 9 + (7 if s_ == 2 else 0)
            ab, cd = (i % 9) // 3, i % 3
            if n in (0, 3, 6, 10, 11, 12, 13):  # symmetric kernels: pair (ab,cd) served by (min,max)
                cd = max(ab, cd)
            need[n] = need.get(n, 0) | (1 << cd)
            need_ab[n] = need_ab.get(n, 0) | (1 << ab)
    if groups & 16:
        need = {n: 7 for n in range(14)}
    elif not ident:  # kernels 7-9 have transposed copies: the product runs on the side with fewer spectra
        for n in (7, 8, 9):
            if n in need and bin(need_ab.get(n, 7)).count("1") < bin(need[n]).count("1"):
                need[n] = need_ab[n]
    return sum(1 if ident else bin(v).count("1") for v in need.values())


for name, groups, ident in (("every product", 31, 0), ("RHS: A+R", 3, 0), ("output: P_T,jm", 4, 0),
                            ("z1l cache: A+R, identical spectra", 3, 1)):
    h.bench_integrals(1, groups, ident)
    h.set_profiling(True)
    h.bench_integrals(reps, groups, ident)
    p = h.profile()
    h.set_profiling(False)
    n, ms = p["k_bilinear"]
    ns = sets(groups, ident)
    tf = flop_set * ns * B * reps / ms * 1e-9
    print("  %-36s %2d sets  k_bilinear %8.3f ms/eval  %6.2f TFLOP/s = %4.1f%% of peak" % (name, ns, ms / reps, tf, 100 * tf / peak))
