#!/usr/bin/env python3
"""k_bilinear of ONE rank of a k-sharded single cosmology, on one GPU: the share of rank 0 of G ranks
(config 3: nk=256), for several v_split.  usage: bench_kshare.py [G] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import redtime_b200 as rt  # noqa: E402
from redtime_b200 import workload as wl  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
base = wl.load_example1(16)
for vs in (1, 2, 3, 4, 5, 6, 8):
    h = rt.RedTimeB200(nk=256, k_shards=G, k_rank=0, v_split=vs)
    h.add_cosmologies(wl.make_cosmologies(1, base))
    h.prepare()
    h.bench_integrals(3, 3, 0)
    h.set_profiling(True)
    h.bench_integrals(reps, 3, 0)
    p = h.profile()
    h.set_profiling(False)
    print("G=%d v_split=%d  per evaluation (event-timed, us): k_bilinear %.1f  k_extrap %.1f  k_pz %.1f  k_assemble %.1f"
          % (G, vs, 1e3 * p["k_bilinear"][1] / reps, 1e3 * p["k_extrap"][1] / reps, 1e3 * p["k_pz"][1] / reps,
             1e3 * p["k_assemble"][1] / reps))
    h.close()
