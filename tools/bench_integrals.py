#!/usr/bin/env python3
"""Micro-benchmark of the integral kernels: B cosmologies, `reps` full evaluations.
usage: bench_integrals.py [B] [reps] [nk]   (kernel variant via RTRG_BIL_VARIANT)"""
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
for c in wl.make_cosmologies(B, base):
    h.add_cosmology(c)
h.prepare()
h.bench_integrals(2)
h.set_profiling(True)
h.bench_integrals(reps)
p = h.profile()
g = rt.grid_info(nk)
flop = 42 * nk * (2.0 * g["nsup"] ** 2 + 6.0 * g["nsup"]) * B * reps
peak = rt.dfma_peak_tflops(0, 0.3)
n, ms = p["k_bilinear"]
print("variant=%s B=%d nk=%d: k_bilinear %.3f ms/launch, %.2f TFLOP/s algorithmic = %.1f%% of measured DFMA peak %.2f; "
      "others: %s" % (os.environ.get("RTRG_BIL_VARIANT", "0"), B, nk, ms / n, flop / ms * 1e-9, 100 * flop / ms * 1e-9 / peak, peak,
                      {k: round(v[1] / max(v[0], 1), 3) for k, v in p.items() if v[0] and k != "k_bilinear"}))
