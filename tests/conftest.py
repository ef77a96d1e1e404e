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
