import lzma
import os
import shutil
import subprocess
import sys
import tarfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, ROOT)
# the harness keeps the weight-table cache inside the repository (the library default is ~/.cache)
os.environ.setdefault("RTRG_CACHE_DIR", os.path.join(ROOT, ".rtrg_cache"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def make_example1_dir(dst, switches=None, z_out=None):
    """Materialise the reference's examples/1_redTime inputs from the packed fixture."""
    os.makedirs(dst, exist_ok=True)
    with lzma.open(os.path.join(GOLDEN, "example1", "camb_transfers.tar.xz")) as f:
        with tarfile.open(fileobj=f) as tar:
            tar.extractall(dst, filter="data")
    src = open(os.path.join(GOLDEN, "example1", "params_redTime.dat")).read().split("\n")
    if switches is not None or z_out is not None:
        vals = [i for i, l in enumerate(src) if l.strip() and not l.startswith("#")]
        # value lines: 0-8 floats, 9-12 switches, 13 z_in, 14 n_out, 15 z list, ...
        if switches is not None:
            for j, s in enumerate(switches):
                src[vals[9 + j]] = str(int(s))
        if z_out is not None:
            src[vals[14]] = str(len(z_out))
            src[vals[15]] = " ".join(repr(float(z)) for z in z_out)
    with open(os.path.join(dst, "params_redTime.dat"), "w") as f:
        f.write("\n".join(src))
    return dst


def oracle_available():
    return os.path.exists(os.path.join(ORACLE_REF, "redTime")) and \
        os.path.exists(os.path.join(ORACLE_REF, "libredtime_stage.so"))


def run_oracle_binary(run_dir, binary="redTime", threads=None):
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(threads or os.cpu_count() or 1)
    out = subprocess.run([os.path.join(ORACLE_REF, binary)], cwd=run_dir, env=env, check=True,
                         stdout=subprocess.PIPE).stdout.decode()
    return out


def run_oracle_stage(run_dir, out_npz, lib="libredtime_stage.so", light=False):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "oracle_stage.py"), run_dir, out_npz,
           "--lib", os.path.join(ORACLE_REF, lib)]
    if light:
        cmd.append("--light")
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    subprocess.run(cmd, check=True, env=env)
    return dict(np.load(out_npz))


def parse_tables(text):
    """stdout of redTime -> (header lines, array [n_out, nk, ncols])."""
    hdr = [l for l in text.split("\n") if l.startswith("#")]
    rows = [l.split() for l in text.split("\n") if l.strip() and not l.startswith("#")]
    arr = np.array(rows, dtype=float)
    return hdr, arr


@pytest.fixture(scope="session")
def example1_dir(tmp_path_factory):
    return make_example1_dir(str(tmp_path_factory.mktemp("example1")))


@pytest.fixture(scope="session")
def example1_full_dir(tmp_path_factory):
    # same inputs, full Time-RG (switches 1 0 1 1, what scripts/runRedTime:101 writes)
    return make_example1_dir(str(tmp_path_factory.mktemp("example1_full")), switches=[1, 0, 1, 1])


@pytest.fixture(scope="session")
def golden_example1():
    import gzip
    with gzip.open(os.path.join(GOLDEN, "example1", "example_redTime_result.dat.gz"), "rt") as f:
        return parse_tables(f.read())


def load_or_make_stage_golden(name, run_dir, lib="libredtime_stage.so"):
    """Committed stage-level golden (made by tests/golden/make_golden.py from the oracle);
    regenerated on the fly from the oracle when missing."""
    path = os.path.join(GOLDEN, name + ".npz")
    if os.path.exists(path):
        return dict(np.load(path))
    if not oracle_available():
        pytest.skip("no golden %s and oracle/_ref not built" % name)
    return run_oracle_stage(run_dir, os.path.join(run_dir, name + ".npz"), lib=lib)


@pytest.fixture(scope="session")
def stage_golden_1loop(example1_dir):
    return load_or_make_stage_golden("example1_stage_1loop", example1_dir)


@pytest.fixture(scope="session")
def stage_golden_full(example1_full_dir):
    return load_or_make_stage_golden("example1_stage_full", example1_full_dir)


# ---------------------------------------------------------------------------------------------
# parity assertions with the MEASURED round-off floor of the reference
# ---------------------------------------------------------------------------------------------
def load_floor(tag):
    """floor[z, k, col] = how far the reference's own table moves when sigma_8 or n_s moves by one
    ulp (tests/golden/make_floor.py): the cancellation noise of redTime.cc:1182-1184 & co."""
    return np.asarray(np.load(os.path.join(GOLDEN, "floor_tables.npz"))[tag], dtype=float)


def smooth_floor(floor):
    """A handful of perturbed runs samples the noise, it does not bound it: widen every entry to the
    max over the redshifts and over the row's two neighbours in k."""
    f = np.max(floor, axis=0, keepdims=True) * np.ones_like(floor)
    g = f.copy()
    g[:, 1:] = np.maximum(g[:, 1:], f[:, :-1])
    g[:, :-1] = np.maximum(g[:, :-1], f[:, 1:])
    return g


def local_scale(ref):
    """max |ref| over the row and its two neighbours on either side: the bispectrum columns P_B,j(k)
    change sign (e.g. at k = 0.095 h/Mpc), where a plain relative error measures the distance to
    the zero crossing rather than the accuracy."""
    a = np.abs(ref)
    scale = a.copy()
    for sh in (1, 2):
        scale[:, sh:] = np.maximum(scale[:, sh:], a[:, :-sh])
        scale[:, :-sh] = np.maximum(scale[:, :-sh], a[:, sh:])
    return scale


FLOOR_C = 5.0  # multiple of the measured floor allowed on top of the relative tolerance


def table_excess(tab, ref, floor, c=FLOOR_C, tol_hi=1e-5):
    """Per column: max over (z, k) of |tab - ref| / allowed, with the north-star tolerances
        columns 1-7   : 1e-6 |ref|
        columns 8-10  : 1e-5 |ref| + c floor(z, k, col)
        columns 11-.. : 1e-5 local_scale(ref) + c floor(z, k, col)      -- at EVERY k
    The floor of columns 8-10 is 1e-12 |ref| for example 1 and most cosmologies; it matters for the
    few that sit on an accept/reject boundary of GSL's step controller, where a 1-ulp change of
    sigma_8 makes the reference take a different step sequence and moves its own P(k) by 1e-4
    (measured: member 15 of the bench batch, tools/diag_parity.py).  A value <= 1 passes."""
    d = np.abs(tab - ref)
    allowed = np.empty_like(d)
    allowed[..., :7] = 1e-6 * np.abs(ref[..., :7])
    allowed[..., 7:10] = 1e-5 * np.abs(ref[..., 7:10]) + c * smooth_floor(floor)[..., 7:10]
    allowed[..., 10:] = tol_hi * local_scale(ref)[..., 10:] + c * smooth_floor(floor)[..., 10:]
    return np.max(d / (allowed + 1e-300), axis=(0, 1))


def assert_table_parity(tab, ref, floor, c=FLOOR_C, tol_hi=1e-5, what=""):
    assert tab.shape == ref.shape, (tab.shape, ref.shape)
    ex = table_excess(tab, ref, floor, c, tol_hi)
    if not np.all(ex <= 1.0):
        col = int(np.argmax(ex))
        d = np.abs(tab - ref)[..., col]
        sf = smooth_floor(floor)[..., col]
        al = (1e-6 * np.abs(ref[..., col]) if col < 7 else (1e-5 * np.abs(ref[..., col]) if col < 10 else
                                                             tol_hi * local_scale(ref)[..., col]) + c * sf)
        iz, ik = np.unravel_index(np.argmax(d / (al + 1e-300)), d.shape)
        raise AssertionError("%s: excess per column %s; worst: column %d at z index %d, k = %.4g: |diff| = %.3e, "
                             "|ref| = %.3e, floor = %.3e (diff / floor = %.1f)"
                             % (what, np.round(ex, 3), col + 1, iz, ref[iz, ik, 0], d[iz, ik], abs(ref[iz, ik, col]),
                                sf[iz, ik], d[iz, ik] / (sf[iz, ik] + 1e-300)))
    return ex


def oracle_tables_with_floor(dirs, binary="redTime", threads_total=None):
    """Run the oracle binary on every run directory AND on two copies with sigma_8 / n_s moved by
    one ulp (all processes concurrently, one OpenMP thread each unless few): returns
    [(hdr_lines, table[n_out*nk, ncols], floor like the table)] -- the on-the-spot version of
    make_floor.py for cosmologies that have no committed floor fixture."""
    import tempfile
    from redtime_b200.workload import perturbed_run_dir
    jobs = []
    tmp = tempfile.mkdtemp(prefix="rtfloor")
    for i, d in enumerate(dirs):
        jobs.append([d] + [perturbed_run_dir(d, os.path.join(tmp, "c%d_p%d" % (i, j)), line, sgn)
                           for j, (line, sgn) in enumerate(((1, +1), (0, -1)))])
    n_proc = sum(len(v) for v in jobs)
    ncpu = threads_total or os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS=str(max(1, ncpu // n_proc)))
    procs = [[subprocess.Popen([os.path.join(ORACLE_REF, binary)], cwd=v, env=env, stdout=subprocess.PIPE)
              for v in variants] for variants in jobs]
    out = []
    for ps in procs:
        txt = [p.communicate()[0].decode() for p in ps]
        assert all(p.returncode == 0 for p in ps)
        hdr, base = parse_tables(txt[0])
        fl = np.zeros_like(base)
        for t in txt[1:]:
            np.maximum(fl, np.abs(parse_tables(t)[1] - base), out=fl)
        out.append((hdr, base, fl))
    shutil.rmtree(tmp, ignore_errors=True)
    return out


def gpu_table_with_floor(run_dir, **cfg):
    """This library's table for run_dir and ITS OWN round-off response: the same 1-ulp changes of
    sigma_8 / n_s the oracle's floor was measured with, as three cosmologies of one batch.

    Why both floors: a table entry that is a 1e8-fold cancellation carries round-off noise in ANY
    double-precision implementation.  At nk = 128 this library is the quieter one (0.3 x the
    reference's response, tools/diag_floor.py); the dense quadrature accumulates N^2 products, the
    reference's FFTs N log N, so the ratio grows with the grid: 3 x at nk = 256 and 16 x at nk = 512
    in columns 15-17 below k = 4e-3 h/Mpc (1.5e-4 relative there against the reference's 5e-6).
    The distance between two noisy numbers is bounded by the sum of their noise amplitudes."""
    import tempfile
    import redtime_b200 as rt
    from redtime_b200.workload import perturbed_run_dir
    tmp = tempfile.mkdtemp(prefix="rtgfloor")
    dirs = [run_dir, perturbed_run_dir(run_dir, os.path.join(tmp, "a"), 1, +1),
            perturbed_run_dir(run_dir, os.path.join(tmp, "b"), 0, -1)]
    h = rt.RedTimeB200(**cfg)
    h.add_cosmologies([rt.read_run_dir(d) for d in dirs])
    h.prepare()
    tables, hdr, hdr0, status = h.run()
    cnt = h.counters(0)
    h.close()
    shutil.rmtree(tmp, ignore_errors=True)
    assert not status.any()
    floor = np.maximum(np.abs(tables[1] - tables[0]), np.abs(tables[2] - tables[0]))
    return tables[0], floor, cnt
