#!/bin/bash
# Timing of the device-side initialisation (k_growth_ode ...) for 1024 Latin-hypercube and for one
# cosmology, and the md5 of a mixed 7-cosmology run (compare across builds: results must not move).
cd "$(dirname "$0")/.."
timeout 300 python - <<'PY'
import hashlib, sys, tempfile
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, redtime_b200 as rt
from redtime_b200 import workload as wl
from conftest import make_example1_dir
t = tempfile.mkdtemp()
d1 = make_example1_dir(t + "/a"); d2 = make_example1_dir(t + "/b", switches=[1, 0, 1, 1])
h = rt.RedTimeB200()
h.add_cosmologies([rt.read_run_dir(d) for d in (d1, d2, d1, d1, d2, d1, d1)])
h.prepare()
tables, hdr, hdr0, status = h.run()
m = hashlib.md5()
for t_ in tables: m.update(np.ascontiguousarray(t_).tobytes())
print("status", [int(s) for s in status], "md5", m.hexdigest())
h.close()
base = wl.load_example1(16)
for B in (1, 1024):
    h = rt.RedTimeB200()
    cos = wl.make_cosmologies(B, base)
    h.add_cosmologies(cos)
    h.prepare()
    h.set_profiling(True)
    for _ in range(3):
        h.device_init()
    p = h.profile()
    h.set_profiling(False)
    out = h.run()
    m = hashlib.md5()
    for t_ in out[0]: m.update(np.ascontiguousarray(t_).tobytes())
    print("B=%d  k_growth_ode %.3f ms/init   (all init kernels: %s)  md5 %s" % (
        B, p["k_growth_ode"][1] / 3, {k: round(v[1] / 3, 3) for k, v in p.items() if v[0] and k in ("k_growth_tabs", "k_qag", "k_beta_reduce", "k_init_state")}, m.hexdigest()))
    h.close()
PY
