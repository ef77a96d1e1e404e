#!/usr/bin/env python3
"""Where the end-to-end time of one batch goes (host side): add / prepare / run."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import redtime_b200 as rt
from redtime_b200 import workload as wl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
base = wl.load_example1()
cos = rt.pack_cosmologies(wl.make_cosmologies(B, base, pinned=(len(sys.argv) > 2 and sys.argv[2] == "pinned")))
h = rt.RedTimeB200()
for it in range(3):
    t0 = time.perf_counter(); h.clear(); t1 = time.perf_counter()
    h.add_cosmologies(cos); t2 = time.perf_counter()
    h.prepare(); t3 = time.perf_counter()
    st = h.run_resident(); t4 = time.perf_counter()
    tabs = h.run_pinned(); t5 = time.perf_counter()
    print("iter %d: clear %.1f ms, add %.1f ms, prepare %.1f ms, run(resident) %.1f ms, run(+D2H) %.1f ms" % (
        it, 1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), 1e3*(t5-t4)))
h.set_profiling(True); h.clear(); h.add_cosmologies(cos); h.prepare(); print({k:v for k,v in h.profile().items() if v[0]})
