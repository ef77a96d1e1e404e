"""Sanitizer substitute: compute-sanitizer is closed on the GPU pool, so the library is also built
with -DRTRG_BOUNDS (libredtime_b200_bounds.so): every computed index of the kernels -- T-table
windows, shared-memory windows, input-pool offsets, partial-sum slots, output rows, the k-shard
packing -- is checked against the extent of its array by device-side asserts.  The end-to-end
paths run on that build in a subprocess (RTRG_LIBRARY); a violated assert aborts the kernel with
cudaErrorAssert and fails the run.  The asserts only read, so the results must agree with the
production build to round-off (the two builds are different code generations of the same source:
1e-14 relative differences in P(k), amplified in the cancelling columns)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import redtime_b200 as rt
from conftest import ROOT

pytestmark = pytest.mark.gpu
BOUNDS = os.path.join(ROOT, "redtime_b200", "libredtime_b200_bounds.so")

SCRIPT = r'''
import sys, threading
import numpy as np
sys.path.insert(0, sys.argv[1])
import redtime_b200 as rt
assert rt.library_path().endswith("libredtime_b200_bounds.so")
d1, d2, out = sys.argv[2], sys.argv[3], sys.argv[4]
res = {}
# mixed batch (1-loop + full Time-RG), all optional column groups, reduced-beta upload, nk = 128
h = rt.RedTimeB200(print_A=1, print_I=1, print_Q=1, print_bias=1, reduce_beta=1)
h.add_cosmologies([rt.read_run_dir(d) for d in (d1, d2, d1, d1, d2, d1, d1)])
h.prepare()
t, hdr, hdr0, st = h.run()
assert not st.any()
res["mixed0"], res["mixed1"] = t[0], t[1]
J, PZ, J0, Jlo = h.integrals_raw(np.log(np.abs(t[0][0, :, 7:10].T.ravel()) + 1e-30))
h.close()
# nk = 256 with a split of the beta-side lags, pipeline API
p = rt.Pipeline(depth=2, nk=256, v_split=2)
tk = [p.submit([rt.read_run_dir(d1), rt.read_run_dir(d2)]) for _ in range(3)]
for x in tk:
    tp, _, _, sp = p.wait(x)
    assert not sp.any()
    res["nk256"] = tp[0].copy()
    p.release(x)
p.close()
# k-sharded over two in-process ranks (loopback transport)
g = rt.LoopbackGroup(2)
hs = [rt.RedTimeB200(k_shards=2, k_rank=r) for r in range(2)]
got, errs = [None, None], []
def work(r):
    try:
        hs[r].add_cosmology(rt.read_run_dir(d2))
        hs[r].kshard_init_loopback(g)
        hs[r].prepare()
        got[r] = hs[r].run()[0][0]
    except Exception as e:
        errs.append(e)
th = [threading.Thread(target=work, args=(r,)) for r in range(2)]
[x.start() for x in th]; [x.join() for x in th]
assert not errs, errs
res["kshard"] = got[0]
[x.close() for x in hs]; g.close()
np.savez(out, **res)
print("BOUNDS_OK")
'''


@pytest.mark.skipif(not os.path.exists(BOUNDS), reason="libredtime_b200_bounds.so not built (make -C redtime_b200/csrc)")
def test_end_to_end_paths_on_the_bounds_checked_build(example1_dir, example1_full_dir, tmp_path):
    out = str(tmp_path / "bounds.npz")
    env = dict(os.environ, RTRG_LIBRARY="libredtime_b200_bounds.so")
    p = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, example1_dir, example1_full_dir, out], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert p.returncode == 0 and "BOUNDS_OK" in p.stdout, (p.stdout[-2000:], p.stderr[-3000:])
    got = dict(np.load(out))
    # the production build gives the same numbers
    h = rt.RedTimeB200(print_A=1, print_I=1, print_Q=1, print_bias=1, reduce_beta=1)
    h.add_cosmologies([rt.read_run_dir(d) for d in (example1_dir, example1_full_dir)])
    h.prepare()
    t, *_ = h.run()
    h.close()
    for a, b in ((t[0], got["mixed0"]), (t[1], got["mixed1"])):
        assert a.shape == b.shape and np.allclose(a[:, :, :10], b[:, :, :10], rtol=1e-10, atol=0)
    h = rt.RedTimeB200()
    h.add_cosmology(rt.read_run_dir(example1_full_dir))
    h.prepare()
    t, *_ = h.run()
    h.close()
    assert t[0].shape == got["kshard"].shape and np.allclose(t[0][:, :, :10], got["kshard"][:, :, :10], rtol=1e-9, atol=0)
