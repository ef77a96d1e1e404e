#!/usr/bin/env python3
"""Extract the parts of the reference's emulator-comparison goldens that can be reproduced
without CAMB: for the massless-neutrino models M001-M010 the columns
k, D, f = dlnD/dlna (k independent) and the header H depend only on params_redTime_M*.dat.
Writes tests/golden/emulator_M001_M010.json.  Needs /root/reference (run in the build container)."""
import json
import os
import re
import sys

REF = "/root/reference/tests/emulator_comparison/output_kmax50_klogint1000"
HERE = os.path.dirname(os.path.abspath(__file__))

out = {}
for m in range(1, 11):
    name = "M%03d" % m
    vals = [l.strip() for l in open(os.path.join(REF, "params_redTime_%s.dat" % name)) if l.strip() and not l.startswith("#")]
    params = [float(v) for v in vals[:9]]
    assert params[5] == 0.0
    z_out = [float(z) for z in vals[12].split()]
    blocks, hdrs = [], []
    for line in open(os.path.join(REF, "redTime_%s.dat" % name)):
        if line.startswith("### main: output"):
            hdrs.append([float(x) for x in re.findall(r"=([0-9.eE+-]+)", line)])
            blocks.append([])
        elif line.strip() and not line.startswith("#"):
            blocks[-1].append([float(x) for x in line.split()[:7]])
    rec = dict(params=params, z_in=float(vals[10]), z_out=z_out, H=[h[3] for h in hdrs], D=[], f=[], k=[r[0] for r in blocks[0]])
    for b in blocks:
        assert len(b) == 128
        assert max(r[1] for r in b) == min(r[1] for r in b)        # D is k independent
        assert all(r[4] == 0 and r[5] == 0 and r[6] == 0 for r in b)  # beta, dlnbeta, P_nu columns
        rec["D"].append(b[0][1])
        rec["f"].append(b[0][2])
    out[name] = rec
json.dump(out, open(os.path.join(HERE, "emulator_M001_M010.json"), "w"), indent=0)
print("wrote", len(out), "models")
