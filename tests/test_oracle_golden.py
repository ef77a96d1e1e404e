"""Pin the ORACLE (oracle/: the unmodified reference sources compiled against the mini-GSL
shim) to the reference's own golden vector examples/1_redTime/example_redTime_result.dat, which
was produced with genuine GSL.  The committed oracle outputs (tests/golden/example1_oracle_*.dat.gz,
made by tests/golden/make_golden.py) are checked always; when oracle/_ref is built the binary
is re-run as well."""
import gzip
import os

import numpy as np
import pytest

from conftest import GOLDEN, oracle_available, parse_tables, run_oracle_binary

NK = 128


def load(name):
    with gzip.open(os.path.join(GOLDEN, name), "rt") as f:
        return parse_tables(f.read())


def check_against_golden(hdr, arr, golden):
    ghdr, gold = golden
    assert hdr == ghdr  # all '#' lines byte-identical (eta, a, z, H, sigma_v^2 to 12 digits)
    t, g = arr.reshape(7, NK, 17), gold.reshape(7, NK, 17)
    e = np.max(np.abs(t - g) / (np.abs(g) + 1e-300), axis=(0, 1))
    assert np.all(e[:7] == 0), e          # columns 1-7: identical in all 12 printed digits
    assert np.all(e[7:10] < 1e-10), e     # P_dd, P_dt, P_tt: RKF45 driver + QAG + FAST-PT
    assert np.all(e[10:15] < 1e-5), e
    hi = g[0, :, 0] > 3.3e-3              # FFT round-off floor of columns 16-17
    ehi = np.max(np.abs(t[:, hi] - g[:, hi]) / (np.abs(g[:, hi]) + 1e-300), axis=(0, 1))
    assert np.all(ehi[15:] < 1e-5), ehi


def test_committed_oracle_output_matches_reference_golden(golden_example1):
    hdr, arr = load("example1_oracle_1loop.dat.gz")
    check_against_golden(hdr, arr, golden_example1)


def test_full_trg_oracle_output_quirks():
    hdr, arr = load("example1_oracle_full.dat.gz")
    t = arr.reshape(7, NK, 17)
    assert not t[:, :, 13:].any()                      # columns 14-17 are zeros
    _, one = load("example1_oracle_1loop.dat.gz")
    assert np.array_equal(t[:, :, :7], one.reshape(7, NK, 17)[:, :, :7])  # linear columns unchanged


@pytest.mark.skipif(not oracle_available(), reason="oracle/_ref not built (make -C oracle)")
def test_oracle_binary_reproduces_reference_golden(example1_dir, golden_example1):
    hdr, arr = parse_tables(run_oracle_binary(example1_dir))
    check_against_golden(hdr, arr, golden_example1)
    chdr, carr = load("example1_oracle_1loop.dat.gz")
    assert hdr == chdr and np.array_equal(arr, carr)   # deterministic across runs / thread counts


# ---- second known-answer set: the reference's emulator-comparison goldens (genuine GSL) -------
def emulator_kat():
    import json
    with open(os.path.join(GOLDEN, "emulator_M001_M010.json")) as f:
        return json.load(f)


def write_massless_run_dir(path, rec, switches=(0, 0, 1, 0)):
    """params of model M00x + the example-1 z=0 transfer file as a stand-in (massless neutrinos
    never open the interpolation files, hdr:523-525; columns k, D, f and the header H do not
    depend on the transfer function)."""
    from redtime_b200 import workload as wl
    base = wl.load_example1(subsample=64)
    c = dict(params=np.array(rec["params"]), switches=list(switches), z_in=rec["z_in"], z_out=np.array(rec["z_out"]),
             k_T=base["k_T"], Tc_T=base["Tc_T"], Tb_T=base["Tb_T"], z_interp=base["z_interp"], k_b=base["k_b"],
             Tc_b=base["Tc_b"], Tnu_b=base["Tnu_b"])
    return wl.write_run_dir(path, c)


@pytest.mark.skipif(not oracle_available(), reason="oracle/_ref not built (make -C oracle)")
@pytest.mark.parametrize("model", ["M001", "M004", "M010"])
def test_oracle_reproduces_emulator_goldens_D_f_H(model, tmp_path):
    """Pins the RK8PD growth ODE + 2-D table interpolation of the oracle on w0wa cosmologies other
    than example 1: 12-digit agreement with what genuine GSL printed."""
    rec = emulator_kat()[model]
    d = write_massless_run_dir(str(tmp_path / model), rec)
    hdr, arr = parse_tables(run_oracle_binary(d))
    t = arr.reshape(len(rec["z_out"]), NK, -1)
    import re
    H = [float(re.findall(r"H=([0-9.eE+-]+)", l)[0]) for l in hdr if l.startswith("### main: output")]
    assert H == rec["H"]
    for i in range(len(rec["z_out"])):
        assert np.array_equal(t[i, :, 0], np.array(rec["k"]))
        assert np.all(t[i, :, 1] == rec["D"][i]) and np.all(t[i, :, 2] == rec["f"][i])
        assert not t[i, :, 4:7].any()
