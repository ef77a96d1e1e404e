#!/bin/bash
# A/B harness for bilinear kernel variants selected by the environment variable RTRG_BIL (the
# round-1 variants tma3 / tma3s2 / ldg384 were measured and dropped, see
# profiles/r01_bilinear_experiments.txt; the shipped library ignores the variable): timing of the
# integral evaluation and an md5 of a mixed 7-cosmology run (the variants must agree bit for bit).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in ${VARIANTS:-ldg tma3 tma3s2}; do
  echo "=== RTRG_BIL=$v"
  RTRG_BIL=$v timeout 120 python tools/bench_integrals.py ${AB_B:-1024} 3 2>&1 | tail -5
  RTRG_BIL=$v timeout 120 python - <<'PY'
import hashlib, sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, redtime_b200 as rt
from conftest import make_example1_dir
import tempfile
t = tempfile.mkdtemp()
d1 = make_example1_dir(t + "/a"); d2 = make_example1_dir(t + "/b", switches=[1, 0, 1, 1])
h = rt.RedTimeB200()
h.add_cosmologies([rt.read_run_dir(d) for d in (d1, d2, d1, d1, d2, d1, d1)])
h.prepare()
tables, hdr, hdr0, status = h.run()
m = hashlib.md5()
for t_ in tables: m.update(np.ascontiguousarray(t_).tobytes())
print("status", list(status), "md5", m.hexdigest())
h1 = rt.RedTimeB200(); h1.add_cosmology(rt.read_run_dir(d1)); h1.prepare()
t1 = h1.run()[0][0]
print("single==batched:", np.array_equal(t1, tables[0]), hashlib.md5(np.ascontiguousarray(t1).tobytes()).hexdigest())
PY
done
