"""The drop-in boundary at process level and parity on cosmologies other than example 1.

  * `redTime_b200` honours the reference's contract (scripts/runRedTime:196-229): no argv, reads
    ./params_redTime.dat from the CWD, prints the tables to stdout.
  * Latin-hypercube cosmologies (config 4 generator, all PRINT* column groups = 84 columns) run
    through the batch API and, from the same run directories, through the oracle binary
    `oracle/_ref/redTime_printall` (the unmodified reference with PRINTA=PRINTI=PRINTQ=PRINTBIAS=1)
    on the host cores of the GPU box."""
import os
import subprocess

import numpy as np
import pytest

import redtime_b200 as rt
from redtime_b200 import workload as wl
from conftest import ORACLE_REF, ROOT, parse_tables, oracle_tables_with_floor, assert_table_parity

pytestmark = pytest.mark.gpu


def test_executable_reproduces_golden_stdout(example1_dir, golden_example1):
    exe = os.path.join(ROOT, "redtime_b200", "redTime_b200")
    p = subprocess.run([exe], cwd=example1_dir, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    hdr, arr = parse_tables(p.stdout)
    ghdr, gold = golden_example1
    assert hdr == ghdr                      # banner + '###main' lines byte-identical
    assert arr.shape == gold.shape
    e = np.max(np.abs(arr - gold) / (np.abs(gold) + 1e-300), axis=0)
    assert np.all(e[:7] < 1e-6) and np.all(e[7:15] < 1e-5), e
    blocks = p.stdout.split("\n\n\n")       # two blank lines after every redshift block
    assert len(blocks) == 8 and blocks[-1] == ""


def test_executable_fails_cleanly_without_inputs(tmp_path):
    exe = os.path.join(ROOT, "redtime_b200", "redTime_b200")
    p = subprocess.run([exe], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode != 0 and "params_redTime.dat" in p.stderr


@pytest.mark.skipif(not os.path.exists(os.path.join(ORACLE_REF, "redTime_printall")), reason="oracle/_ref not built")
def test_latin_hypercube_cosmologies_all_columns(tmp_path):
    base = wl.load_example1(subsample=27)   # ~570 rows: the scripts' default CAMB density
    cosmos = wl.make_cosmologies(4, base, seed=wl.SEED + 1)
    dirs = [wl.write_run_dir(str(tmp_path / ("M%03d" % i)), c) for i, c in enumerate(cosmos)]
    h = rt.RedTimeB200(print_A=1, print_I=1, print_Q=1, print_bias=1)
    h.add_cosmologies([rt.read_run_dir(d) for d in dirs])  # the same bytes the oracle reads
    h.prepare()
    tables, hdr, hdr0, status = h.run()
    h.close()
    assert not status.any()
    # the oracle on the same directories, plus two runs each with sigma_8 / n_s moved by one ulp:
    # its own round-off floor for exactly these cosmologies
    orc = oracle_tables_with_floor(dirs, "redTime_printall")
    for d, tab, (rhdr, ref, floor) in zip(dirs, tables, orc):
        assert tab.shape[2] == 84
        # all 84 columns at every k (optional groups relative to the local scale: sign changes)
        assert_table_parity(tab, ref.reshape(tab.shape), floor.reshape(tab.shape), what=d)
        # header lines: eta, a, z, H, sigma_v^2 to the 12 printed digits
        p = os.path.join(d, "mine.dat")
        i = dirs.index(d)
        rt.print_result(p, 128, tab, hdr[i], hdr0[i])
        mine = [l for l in open(p).read().split("\n") if l.startswith("#")]
        assert mine == rhdr


def test_table_cache_round_trip(tmp_path, example1_dir, stage_golden_1loop, monkeypatch):
    """The cosmology-independent kernels T_n are cached on disk (RTRG_CACHE_DIR); a handle
    created from the cache must give bit-identical integrals."""
    import time
    monkeypatch.setenv("RTRG_CACHE_DIR", str(tmp_path / "cache"))
    g = stage_golden_1loop
    res, dt = [], []
    for _ in range(2):
        t0 = time.perf_counter()
        h = rt.RedTimeB200()
        dt.append(time.perf_counter() - t0)
        h.add_cosmology(rt.read_run_dir(example1_dir))
        h.prepare()
        res.append(h.integrals_raw(g["yp"][:3 * 128]))
        h.close()
    files = os.listdir(str(tmp_path / "cache"))
    assert len(files) == 1 and files[0].startswith("T_v") and "_nk128_" in files[0]
    for a, b in zip(res[0][:3], res[1][:3]):
        assert np.array_equal(a, b)
    assert res[0][3] == res[1][3]
    assert dt[1] < dt[0]
    # a damaged file (one flipped byte anywhere in the 24 MB) must be detected -- the checksum covers
    # every byte -- and rebuilt, not trusted
    path = os.path.join(str(tmp_path / "cache"), files[0])
    good = open(path, "rb").read()
    bad = bytearray(good)
    bad[len(bad) // 2 + 12345] ^= 0x10
    open(path, "wb").write(bytes(bad))
    h = rt.RedTimeB200()
    h.add_cosmology(rt.read_run_dir(example1_dir))
    h.prepare()
    again = h.integrals_raw(g["yp"][:3 * 128])
    h.close()
    for a, b in zip(res[0][:3], again[:3]):
        assert np.array_equal(a, b)
    assert open(path, "rb").read() == good      # rewritten from the rebuilt tables
    monkeypatch.setenv("RTRG_CACHE_DIR", "off")
    h = rt.RedTimeB200()
    h.close()
    assert len(os.listdir(str(tmp_path / "cache"))) == 1


def test_emulator_goldens_massless_neutrinos(tmp_path):
    """The reference's own goldens M001-M010 (massless neutrinos, w0wa; genuine GSL): k, D, f and
    the header H.  Also exercises the n_z = 0 path (Beta_P == 0) and linear-only switches."""
    import re
    from test_oracle_golden import emulator_kat, write_massless_run_dir
    kat = emulator_kat()
    dirs = [write_massless_run_dir(str(tmp_path / m), kat[m]) for m in sorted(kat)]
    h = rt.RedTimeB200()
    h.add_run_dirs(dirs)
    h.prepare()
    tables, hdr, hdr0, status = h.run()
    h.close()
    assert not status.any()
    for m, tab, hd in zip(sorted(kat), tables, hdr):
        rec = kat[m]
        assert tab.shape == (8, 128, 10)   # k, 6 linear columns, P_dd, P_dt, P_tt (print_rsd = 0)
        for i in range(8):
            assert np.max(np.abs(tab[i, :, 0] / np.array(rec["k"]) - 1)) < 1e-11   # 12 printed digits
            assert np.max(np.abs(tab[i, :, 1] / rec["D"][i] - 1)) < 1e-11, m
            assert np.max(np.abs(tab[i, :, 2] / rec["f"][i] - 1)) < 1e-11, m
            assert abs(hd[i, 3] / rec["H"][i] - 1) < 1e-11
            assert not tab[i, :, 4:7].any()


def test_cpp_batch_front_end(tmp_path, golden_example1):
    """redTimeBatch_b200 <manifest>: the C++-only replacement of the runRedTimeBatch loop."""
    from conftest import make_example1_dir
    d1 = make_example1_dir(str(tmp_path / "M001"))
    d2 = make_example1_dir(str(tmp_path / "M002"), switches=[1, 0, 1, 1])
    (tmp_path / "manifest.txt").write_text("# models\nM001\n%s   # absolute path\n" % d2)
    exe = os.path.join(ROOT, "redtime_b200", "redTimeBatch_b200")
    p = subprocess.run([exe, str(tmp_path / "manifest.txt")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=300)
    assert p.returncode == 0, p.stderr
    assert "2 models, 0 failed" in p.stdout
    hdr, arr = parse_tables(open(os.path.join(d1, "redTime_M001.dat")).read())
    ghdr, gold = golden_example1
    assert hdr == ghdr
    assert np.max(np.abs(arr[:, :10] / gold[:, :10] - 1)) < 1e-9
    assert os.path.exists(os.path.join(d2, "redTime_M002.dat"))
    # sharding a manifest over processes: first/stride
    os.remove(os.path.join(d1, "redTime_M001.dat"))
    p = subprocess.run([exe, str(tmp_path / "manifest.txt"), "1", "2"], stdout=subprocess.PIPE, text=True, timeout=300)
    assert p.returncode == 0 and "1 models" in p.stdout and not os.path.exists(os.path.join(d1, "redTime_M001.dat"))
