#!/usr/bin/env python3
"""How do this library's own round-off response and its distance to the oracle compare with the
oracle's measured floor (tests/golden/floor_tables.npz)?  Example 1, 1-loop, nk = 128 / 256 / 512."""
import gzip
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import redtime_b200 as rt  # noqa: E402
from redtime_b200 import workload as wl  # noqa: E402
from conftest import GOLDEN, load_floor, local_scale, make_example1_dir, parse_tables, smooth_floor  # noqa: E402

np.set_printoptions(linewidth=200, precision=2)
tmp = tempfile.mkdtemp()
d0 = make_example1_dir(os.path.join(tmp, "base"))
dirs = [d0, wl.perturbed_run_dir(d0, os.path.join(tmp, "a"), 1, +1), wl.perturbed_run_dir(d0, os.path.join(tmp, "b"), 0, -1),
        wl.perturbed_run_dir(d0, os.path.join(tmp, "c"), 1, -1), wl.perturbed_run_dir(d0, os.path.join(tmp, "e"), 0, +1)]
VS = [int(x) for x in sys.argv[1:]] or [1]
for tag, nk, cfg in [(t, n, dict(c, v_split=v)) for v in VS for (t, n, c) in
                     (("1loop", 128, {}), ("nk256_1loop", 256, dict(nk=256)),
                      ("HIGH_ACCURACY_1loop", 512, dict(nk=512, eps_abs=1e-15, eps_rel=1e-6)))]:
    with gzip.open(os.path.join(GOLDEN, "example1_oracle_%s.dat.gz" % tag), "rt") as f:
        ref = parse_tables(f.read())[1].reshape(7, nk, 17)
    h = rt.RedTimeB200(**cfg)
    h.add_cosmologies([rt.read_run_dir(d) for d in dirs])
    h.prepare()
    tables, *_ = h.run()
    h.close()
    base = tables[0]
    gfl = np.zeros_like(base)
    for t in tables[1:]:
        np.maximum(gfl, np.abs(t - base), out=gfl)
    ofl = smooth_floor(load_floor(tag))
    gfs = smooth_floor(gfl)
    d = np.abs(base - ref)
    tol = 1e-5 * local_scale(ref)
    print("== %s v_split=%d" % (tag, cfg["v_split"]))
    for col in range(14, 17):
        m = ofl[..., col] > 0
        excess = np.maximum(d[..., col] - tol[..., col], 0)
        r1 = np.max(excess[m] / ofl[..., col][m]) if m.any() else 0.0
        lo = (ref[0, :, 0] < 4e-3)
        rat = np.median(gfs[:, lo, col] / (ofl[:, lo, col] + 1e-300))
        iz, ik = np.unravel_index(np.argmax(np.where(m, excess / (ofl[..., col] + 1e-300), 0)), excess.shape)
        print("  col %2d: max (|ours-oracle| - 1e-5 scale)/oracle_floor = %6.1f at k=%.3g; median ours_floor/oracle_floor (k<4e-3) "
              "= %.2f; rel noise there: oracle %.1e ours %.1e" % (col + 1, r1, ref[0, ik, 0], rat,
               ofl[iz, ik, col] / (abs(ref[iz, ik, col]) + 1e-300), gfs[iz, ik, col] / (abs(ref[iz, ik, col]) + 1e-300)))
