#!/usr/bin/env python3
"""Measure the reference's own round-off floor (cancellation site redTime.cc:1182-1184)
and commit it as fixtures, so that the GPU parity tests can assert

    |ours - oracle| <= tol * |oracle| + c * floor(z, k, column)        at EVERY k

instead of masking the lowest wavenumbers.  TEST INFRASTRUCTURE; needs oracle/_ref (make -C oracle).

floor_tables.npz   per oracle build (example 1): max over perturbed runs of |table - unperturbed
                   table|, the perturbation being ONE ulp in sigma_8 or n_s of params_redTime.dat
                   (ln P moves by <= 1e-15: every physical change is 1e-10 below the tolerances,
                   what remains is how far the FFT/cancellation round-off of the reference moves).
floor_stage.npz    the same for compute_Aacdbef_Rlabc_PTjm_PMRn_full on the stage fixtures' state
                   yp with ln P moved by a random -1/0/+1 ulp (tests/oracle_stage.py --floor).

usage: python tests/golden/make_floor.py [tag ...]     (default: every configuration)"""
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import conftest  # noqa: E402

# tag -> (oracle binary, switches, nk, ncols)
CONFIGS = {
    "1loop": ("redTime", None, 128, 17),
    "full": ("redTime", [1, 0, 1, 1], 128, 17),
    "printall_1loop": ("redTime_printall", None, 128, 84),
    "nk256_1loop": ("redTime_nk256", None, 256, 17),
    "nk256_full": ("redTime_nk256", [1, 0, 1, 1], 256, 17),
    "hiacc_full": ("redTime_hiacc", [1, 0, 1, 1], 256, 17),
    "HIGH_ACCURACY_1loop": ("redTime_HIGH_ACCURACY", None, 512, 17),
}
# (index of the value line in params_redTime.dat, direction): n_s is line 0, sigma_8 line 1
PERTURBATIONS = [(1, +1), (1, -1), (0, +1), (0, -1)]


def perturbed_dir(dst, switches, line, direction):
    d = conftest.make_example1_dir(dst, switches=switches)
    p = os.path.join(d, "params_redTime.dat")
    src = open(p).read().split("\n")
    vals = [i for i, l in enumerate(src) if l.strip() and not l.startswith("#")]
    x = float(src[vals[line]].split()[0])
    src[vals[line]] = "%.17g" % np.nextafter(x, np.inf if direction > 0 else -np.inf)
    open(p, "w").write("\n".join(src))
    return d


def run_binary(binary, d, threads):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    out = subprocess.run([os.path.join(conftest.ORACLE_REF, binary)], cwd=d, env=env, check=True,
                         stdout=subprocess.PIPE).stdout.decode()
    return conftest.parse_tables(out)[1]


def table_floor(tag):
    binary, sw, nk, ncols = CONFIGS[tag]
    ncpu = os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as tmp:
        dirs = [conftest.make_example1_dir(os.path.join(tmp, "base"), switches=sw)]
        dirs += [perturbed_dir(os.path.join(tmp, "p%d" % i), sw, l, s) for i, (l, s) in enumerate(PERTURBATIONS)]
        with ThreadPoolExecutor(len(dirs)) as ex:
            tabs = list(ex.map(lambda d: run_binary(binary, d, max(1, ncpu // len(dirs))), dirs))
    base = tabs[0].reshape(-1, nk, ncols)
    fl = np.zeros_like(base)
    for t in tabs[1:]:
        np.maximum(fl, np.abs(t.reshape(base.shape) - base), out=fl)
    # float32, rounded up, is plenty for a magnitude
    return np.nextafter(fl.astype(np.float32), np.float32(np.inf)) * (fl > 0)


def stage_floor(lib, switches, n=6):
    with tempfile.TemporaryDirectory() as tmp:
        d = conftest.make_example1_dir(os.path.join(tmp, "a"), switches=switches)
        out = os.path.join(tmp, "floor.npz")
        cmd = [sys.executable, os.path.join(conftest.ROOT, "tests", "oracle_stage.py"), d, out, "--lib",
               os.path.join(conftest.ORACLE_REF, lib), "--floor", str(n)]
        subprocess.run(cmd, check=True, env=dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1)))
        return dict(np.load(out))


if __name__ == "__main__":
    tags = sys.argv[1:] or list(CONFIGS) + ["stage"]
    path = os.path.join(HERE, "floor_tables.npz")
    res = dict(np.load(path)) if os.path.exists(path) else {}
    for tag in tags:
        if tag == "stage":
            continue
        res[tag] = table_floor(tag)
        print(tag, "floor: max rel over cols", flush=True)
        np.savez_compressed(path, **res)
    if "stage" in tags:
        st = {}
        for name, lib in (("nk128", "libredtime_stage.so"), ("nk256", "libredtime_stage_nk256.so")):
            for key, v in stage_floor(lib, None).items():
                st["%s_%s" % (name, key)] = v if key == "k" else v.astype(np.float32) * np.float32(1.000001)
        np.savez_compressed(os.path.join(HERE, "floor_stage.npz"), **st)
    print("floors written to", HERE)
