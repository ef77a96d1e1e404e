#!/usr/bin/env python3
"""Per-cosmology parity of the bench batch's first N cosmologies against oracle/_ref/redTime
(diagnostic; writes gpurun_out/diag_parity.npz).  usage: diag_parity.py [N] [subsample]"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import redtime_b200 as rt  # noqa: E402
from redtime_b200 import workload as wl  # noqa: E402
from conftest import parse_tables  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sub = int(sys.argv[2]) if len(sys.argv) > 2 else 1
base = wl.load_example1(sub)
cosmos = wl.make_cosmologies(N, base, seed=wl.SEED, total=1024)
tmp = tempfile.mkdtemp()
dirs = [wl.write_run_dir(os.path.join(tmp, "c%02d" % i), c) for i, c in enumerate(cosmos)]
env = dict(os.environ, OMP_NUM_THREADS="1")
procs = [subprocess.Popen([os.path.join(ROOT, "oracle", "_ref", "redTime")], cwd=d, env=env, stdout=subprocess.PIPE) for d in dirs]
res = {}
for rb in (1, 0):
    h = rt.RedTimeB200(reduce_beta=rb)
    h.add_cosmologies(cosmos)
    h.prepare()
    tables, hdr, hdr0, status = h.run()
    res[rb] = (tables, [h.counters(i) for i in range(N)], status.copy(), [h.rmax_history(i) for i in range(N)])
    h.close()
refs = [parse_tables(p.communicate()[0].decode())[1].reshape(8, 128, 17) for p in procs]
np.set_printoptions(linewidth=220, precision=2)
out = {}
for i in range(N):
    r = refs[i]
    line = "c%02d" % i
    for rb in (1, 0):
        t = res[rb][0][i]
        rel = np.abs(t - r) / (np.abs(r) + 1e-300)
        e17, e810 = rel[:, :, :7].max(), rel[:, :, 7:10].max()
        iz, ik, ic = np.unravel_index(np.argmax(rel[:, :, 7:10]), rel[:, :, 7:10].shape)
        cnt = res[rb][1][i]
        line += "  rb=%d: c1-7 %.1e c8-10 %.1e @z%d k%d col%d  att %d rej %d" % (rb, e17, e810, iz, ik, ic + 8, cnt["attempts"], cnt["rejected"])
    p = cosmos[i]["params"]
    line += "  w0 %.3f wa %.3f On %.5f h %.3f" % (p[7], p[8], p[5], p[2])
    print(line)
    r1, r0 = res[1][3][i], res[0][3][i]
    m = np.minimum(np.abs(r1 / 1.1 - 1), np.abs(r1 / 0.5 - 1))
    j = int(np.argmin(m))
    print("      attempt with the smallest decision margin: #%d rmax = %.15g (margin %.2e); rmax rb=1 vs rb=0 max rel diff "
          "over the common prefix %.2e" % (j, r1[j], m[j], np.max(np.abs(r1[:min(len(r1), len(r0))] / r0[:min(len(r1), len(r0))] - 1))))
    if len(r1) != len(r0) or i == 15:
        print("      rmax rb=1:", np.array2string(r1, precision=6))
        print("      rmax rb=0:", np.array2string(r0, precision=6))
    out["gpu%d" % i], out["ref%d" % i] = res[1][0][i], r
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "diag_parity.npz"), **out)
