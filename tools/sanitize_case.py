#!/usr/bin/env python3
"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): a mixed batch of 7
cosmologies (packed kernel), one single run (per-cosmology kernel), the parity hooks and a
2-rank loopback k-shard run, all at nk=128 with subsampled tables."""
import os, sys, tempfile, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import redtime_b200 as rt
from redtime_b200 import workload as wl

base = wl.load_example1(64)
cos = wl.make_cosmologies(7, base, z_out=(1.0, 0.0))
cos[1]["switches"] = [1, 0, 1, 1]
cos[2]["switches"] = [0, 0, 1, 0]
h = rt.RedTimeB200(print_A=1, print_bias=1)
h.add_cosmologies(cos)
h.prepare()
t, hdr, hdr0, st = h.run()
print("batch status", st, [x.shape for x in t][:2])
y, _ = h.initial_state(0)
h.integrals_raw(y[:3 * 128], 0)
h.derivatives(0.3, y, 1)
h.D_dD(1.0, np.array([1e-3, 0.1, 1.0]), 0)
h.Beta_P(0.5, np.array([1e-3, 0.1, 1.0]), 0)
h.close()
h = rt.RedTimeB200()
h.add_cosmology(cos[1])
h.prepare()
print("single status", h.run()[3])
h.close()
G = 2
grp = rt.LoopbackGroup(G)
hs = [rt.RedTimeB200(k_shards=G, k_rank=r) for r in range(G)]
def work(r):
    hs[r].add_cosmology(cos[1]); hs[r].kshard_init_loopback(grp); hs[r].prepare(); hs[r].run()
th = [threading.Thread(target=work, args=(r,)) for r in range(G)]
[x.start() for x in th]; [x.join() for x in th]
print("kshard done")
