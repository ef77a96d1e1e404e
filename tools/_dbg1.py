import sys, os, tempfile, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import conftest, redtime_b200 as rt
np.set_printoptions(linewidth=200, precision=6)
g = dict(np.load('tests/golden/example1_stage_1loop.npz'))
with tempfile.TemporaryDirectory() as tmp:
    d = conftest.make_example1_dir(tmp)
    h = rt.RedTimeB200(); h.add_cosmology(rt.read_run_dir(d)); h.prepare()
    dy = h.derivatives(1.3, g['yp']).reshape(41,128)
    r = g['rhs_dy'][1].reshape(41,128)
    for j in range(41):
        print(j, dy[j,[0,40,100,127]], r[j,[0,40,100,127]])
