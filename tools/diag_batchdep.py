import sys, os, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import redtime_b200 as rt
from conftest import make_example1_dir
tmp = tempfile.mkdtemp()
d1 = make_example1_dir(os.path.join(tmp, "a")); d2 = make_example1_dir(os.path.join(tmp, "b"), switches=[1, 0, 1, 1])
print("library:", rt.library_path())
def run(dirs, **cfg):
    h = rt.RedTimeB200(**cfg); h.add_cosmologies([rt.read_run_dir(d) for d in dirs]); h.prepare(); t, *_ = h.run(); h.close(); return t
for cfg in (dict(), dict(print_A=1, print_I=1, print_Q=1, print_bias=1), dict(reduce_beta=1)):
    a = run([d1, d2], **cfg); b = run([d1, d2, d1, d1, d2, d1, d1], **cfg); c = run([d1] * 40, **cfg)
    for name, x, y in (("1loop B2 vs B7", a[0], b[0]), ("full B2 vs B7", a[1], b[1]), ("1loop B2 vs B40", a[0], c[0])):
        d = np.abs(x - y); bad = np.argwhere(d > 0)
        print(cfg, name, "identical" if bad.size == 0 else "DIFFER at cols %s, max rel %.2e" % (sorted(set(bad[:, 2])), np.max(d / (np.abs(y) + 1e-300))))
np.save(os.path.join(ROOT, "gpurun_out", "batchdep_%s.npy" % os.path.basename(rt.library_path())), a[0])
