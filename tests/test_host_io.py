"""Host side of the drop-in boundary: params_redTime.dat / CAMB readers and the table printer
(reference: src/AU_cosmological_parameters.h:231-353,547-627,790-832; src/redTime.cc:1602-1741)."""
import gzip
import os

import numpy as np
import pytest

import redtime_b200 as rt
from redtime_b200 import workload as wl
from conftest import GOLDEN, make_example1_dir


def test_read_example1(example1_dir):
    d = rt.read_run_dir(example1_dir)
    assert np.allclose(d["params"], [0.87969, 0.80560, 0.73418, 0.286233679143621, 0.0430930827493416, 0.00576437405571056,
                                     2.726, -1.2147, -1.112],
                       rtol=0, atol=1e-15)
    assert d["switches"] == [1, 1, 1, 1] and d["z_in"] == 200
    assert list(d["z_out"]) == [5, 4, 3, 2, 1, 0.5, 0]
    assert list(d["z_interp"]) == [200, 100, 50, 20, 10, 5, 4, 3, 2, 1, 0.5, 0]
    assert d["k_T"].size == 15447 and d["Tc_b"].shape == (12, 15447) and d["Tnu_b"].shape == (12, 15447)
    assert np.all(np.diff(d["k_T"]) > 0) and np.array_equal(d["k_T"], d["k_b"])
    assert np.array_equal(d["Tc_T"], d["Tc_b"][-1])  # the z=0 file is also the last interpolation file


def test_missing_inputs_are_errors(tmp_path):
    with pytest.raises(rt.RtrgError):
        rt.read_run_dir(str(tmp_path))
    d = make_example1_dir(str(tmp_path / "x"))
    os.remove(os.path.join(d, "camb_transfer_z3.dat"))
    with pytest.raises(rt.RtrgError):
        rt.read_run_dir(d)


def test_massless_neutrinos_need_only_the_z0_file(tmp_path):
    """hdr:523-525: with f_nu < 1e-10 the interpolation files are never opened."""
    d = make_example1_dir(str(tmp_path / "x"))
    lines = open(os.path.join(d, "params_redTime.dat")).read().split("\n")
    vals = [i for i, l in enumerate(lines) if l.strip() and not l.startswith("#")]
    lines[vals[5]] = "0.0"
    open(os.path.join(d, "params_redTime.dat"), "w").write("\n".join(lines))
    for z in ("200", "100", "50"):
        os.remove(os.path.join(d, "camb_transfer_z%s.dat" % z))
    c = rt.read_run_dir(d)
    assert c["z_interp"].size == 0 and c["k_b"].size == 0 and c["k_T"].size == 15447


def test_written_run_dir_round_trips(tmp_path):
    base = wl.load_example1(subsample=64)
    c = wl.make_cosmologies(3, base)[2]
    d = rt.read_run_dir(wl.write_run_dir(str(tmp_path / "c"), c))
    for key in ("params", "z_out", "k_T", "Tc_T", "Tb_T", "Tc_b", "Tnu_b", "z_interp"):
        assert np.array_equal(np.asarray(d[key]), np.asarray(c[key])), key


def test_printer_reproduces_the_reference_stdout(tmp_path, golden_example1):
    """Feeding the golden's own numbers through rtrg_print_result must give back the golden
    file: banner, '###main' lines, setw(20)/setprecision(12) rows, two blank lines per block."""
    hdr_lines, arr = golden_example1
    with gzip.open(os.path.join(GOLDEN, "example1", "example_redTime_result.dat.gz"), "rt") as f:
        ref_text = f.read()
    tab = arr.reshape(7, 128, 17)
    import re
    hdr0 = np.array([float(x) for x in re.findall(r"= ([0-9.eE+-]+)", hdr_lines[1])])
    hdr = np.zeros((7, 5))
    for i, l in enumerate(hdr_lines[2:]):
        hdr[i] = [float(x) for x in re.findall(r"=([0-9.eE+-]+)", l)]
    p = str(tmp_path / "o.dat")
    rt.print_result(p, 128, tab, hdr, hdr0)
    mine = open(p).read()
    assert mine.split("\n")[:3] == ref_text.split("\n")[:3]
    # 12 significant digits survive the parse -> print round trip for all but last-digit ties
    ml, rl = mine.split("\n"), ref_text.split("\n")
    assert len(ml) == len(rl)
    same = sum(a == b for a, b in zip(ml, rl))
    assert same >= 0.99 * len(rl)
    assert all(len(a) == len(b) for a, b in zip(ml, rl))


def test_fast_parser_matches_numpy_loadtxt(example1_dir):
    """the std::from_chars ingestion must give the same doubles as a reference
    text parser (numpy) for every CAMB column the path uses."""
    d = rt.read_run_dir(example1_dir)
    z0 = np.loadtxt(os.path.join(example1_dir, "camb_transfer_z0.dat"))
    assert np.array_equal(z0[:, 0], d["k_T"]) and np.array_equal(z0[:, 1], d["Tc_T"]) and np.array_equal(z0[:, 2], d["Tb_T"])
    for zs, iz in ((".5", 10), ("200", 0), ("3", 7)):
        t = np.loadtxt(os.path.join(example1_dir, "camb_transfer_z%s.dat" % zs))
        assert np.array_equal(t[:, 1], d["Tc_b"][iz]) and np.array_equal(t[:, 5], d["Tnu_b"][iz])


def test_comment_lines_in_transfer_files(tmp_path):
    """The z=0 transfer file may carry '#' lines between rows (hdr:805-821 discards them); the
    interpolation files after the first are a plain token stream in the reference (hdr:596-622),
    where a comment silently truncates the table -- here that is an error."""
    base = wl.load_example1(subsample=256)
    c = wl.make_cosmologies(1, base)[0]
    d = wl.write_run_dir(str(tmp_path / "c"), c)
    rows = open(os.path.join(d, "camb_transfer_z0.dat")).read().split("\n")
    rows = ["# header", "# k/h  CDM  baryon ..."] + rows[:5] + ["# a comment between rows"] + rows[5:]
    open(os.path.join(d, "transfer_T.dat"), "w").write("\n".join(rows))
    par = open(os.path.join(d, "params_redTime.dat")).read().replace("camb_transfer_z0.dat", "transfer_T.dat")
    open(os.path.join(d, "params_redTime.dat"), "w").write(par)
    r = rt.read_run_dir(d)
    assert np.array_equal(r["k_T"], c["k_T"]) and np.array_equal(r["Tb_T"], c["Tb_T"])
    open(os.path.join(d, "camb_transfer_z3.dat"), "w").write("# comment\n" + open(os.path.join(d, "camb_transfer_z3.dat")).read())
    with pytest.raises(rt.RtrgError):
        rt.read_run_dir(d)


def test_camb_modern_13_column_files(tmp_path):
    """-DCAMB_MODERN in the reference (hdr:76-80): 13 columns per row, same k / CDM / baryon /
    massive-nu positions.  Here: rtrg_read_run_dir(dir, camb_modern=1), RTRG_CAMB_MODERN=1."""
    base = wl.load_example1(subsample=256)
    c = wl.make_cosmologies(1, base)[0]
    d = wl.write_run_dir(str(tmp_path / "c"), c)
    for name in os.listdir(d):
        if name.startswith("camb_transfer_z"):
            t = np.loadtxt(os.path.join(d, name))
            wide = np.hstack([t, np.full((t.shape[0], 6), 7.0)])
            np.savetxt(os.path.join(d, name), wide, fmt="%.17e")
    r = rt.read_run_dir(d, camb_modern=True)
    for key in ("k_T", "Tc_T", "Tb_T", "Tc_b", "Tnu_b"):
        assert np.array_equal(np.asarray(r[key]), np.asarray(c[key])), key
    # read as 7-column files the rows no longer line up: the k lists of the files disagree
    with pytest.raises(rt.RtrgError):
        rt.read_run_dir(d, camb_modern=False)


def test_fast_parser_equals_numpy_on_ragged_files(tmp_path):
    """The table reader converts only the columns the run consumes, takes an exact fast path for
    short decimal mantissas and a fixed-width shortcut once a row's layout is known.  Files that
    break every one of those assumptions must still give the doubles np.loadtxt (strtod) gives:
    rows of changing width, 17-digit mantissas, explicit '+' signs, huge/tiny exponents, comment
    lines between the rows of the z = 0 file, CRLF line ends."""
    rng = np.random.default_rng(11)
    n = 257
    d = str(tmp_path / "r")
    os.makedirs(d)
    k = np.sort(np.exp(rng.uniform(np.log(1e-5), np.log(50.0), n)))
    zs = ["10", "1", "0"]

    def fmt(x, i, j):
        style = (i + j) % 5
        if i < 40:
            return "%14.6E" % x                       # CAMB's fixed-width block first: the layout is learnt ...
        if style == 0:
            return "%.17e" % x                        # ... and then broken in every way
        if style == 1:
            return "+%.9g" % x
        if style == 2:
            return "   %.5E" % x
        if style == 3:
            return "%.12f" % x
        return "%25.16e" % x

    tabs = {}
    for z in zs:
        cols = np.abs(rng.standard_normal((n, 7))) * 10.0 ** rng.integers(-30, 30, size=(n, 7))
        cols[:, 0] = k
        lines = [" ".join(fmt(cols[i, j], i, j) for j in range(7)) for i in range(n)]
        txt = ("\r\n" if z == "1" else "\n").join(lines) + "\n"
        open(os.path.join(d, "camb_transfer_z%s.dat" % z), "w").write(txt)
        tabs[z] = np.loadtxt(os.path.join(d, "camb_transfer_z%s.dat" % z))
        if z == "0":   # the z = 0 transfer file proper may carry comments between its rows (hdr:805-821)
            for i in range(n - 1, 0, -50):
                lines.insert(i, "# a comment between rows")
            open(os.path.join(d, "transfer_main.dat"), "w").write("\n".join(lines) + "\n")
    p = ["# ragged", "0.96", "0.8", "0.7", "0.3", "0.045", "0.004", "2.726", "-1", "0", "1", "1", "1", "1", "200", "2",
         "1 0", "transfer_main.dat", "0", "camb_transfer_z", "3", " ".join(zs)]
    open(os.path.join(d, "params_redTime.dat"), "w").write("\n".join(p) + "\n")
    c = rt.read_run_dir(d)
    assert np.array_equal(c["k_T"], tabs["0"][:, 0]) and np.array_equal(c["Tc_T"], tabs["0"][:, 1])
    assert np.array_equal(c["Tb_T"], tabs["0"][:, 2])
    for iz, z in enumerate(zs):
        assert np.array_equal(c["Tc_b"][iz], tabs[z][:, 1]) and np.array_equal(c["Tnu_b"][iz], tabs[z][:, 5]), z
    assert np.array_equal(c["k_b"], tabs["10"][:, 0])


def test_number_format_is_printf_20_12g():
    """The table printer formats with std::to_chars(general, 12) instead of printf: every value must
    come out exactly as "%20.12g" (what setw(20) << setprecision(12) prints, rt:1670-1737)."""
    import ctypes as C
    lib = rt.load_library()
    rng = np.random.default_rng(3)
    v = np.concatenate([rng.standard_normal(20000) * 10.0 ** rng.integers(-320, 300, 20000),
                        [0.0, -0.0, 1.0, 1e-5, 9.99999999999949e-5, 9.999999999995e-5, 123456789012.0, 999999999999.5,
                         1e12, 1e11, 0.1, 1e-4, 1e-300, 5e-324, 1.7976931348623157e308, 100000.0, 1e21, 0.001,
                         np.inf, -np.inf]])
    buf = C.create_string_buffer(20 * v.size + 1)
    lib.rtrg_format_g12.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_char_p]
    assert lib.rtrg_format_g12(v.ctypes.data_as(C.POINTER(C.c_double)), int(v.size), buf) == 0
    got = buf.value.decode()
    for i, x in enumerate(v):
        assert got[20 * i:20 * i + 20] == "%20.12g" % x, (x, got[20 * i:20 * i + 20], "%20.12g" % x)


def test_vectorised_beta_row_staging_is_bit_identical(tmp_path):
    """rtrg_add_cosmologies (reduce_beta = 1) forms beta(a = 1, k_b) with host_stage.cc: packed divisions,
    interpolation weights formed once.  Same bits as cub4 / lin2 (rtrg_math.h) column by column, on the
    example's own table size and on ragged lengths (vector remainders)."""
    import ctypes as C
    import subprocess
    from conftest import ROOT
    so = str(tmp_path / "libsh.so")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC",
                    os.path.join(ROOT, "tests", "harness", "stage_harness.cc"), "-o", so], check=True)
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    lib.sh_row_cubic.argtypes = [dp, dp, C.c_size_t, C.c_double, dp, C.c_double, dp, dp]
    lib.sh_row_linear.argtypes = [dp, dp, C.c_size_t, C.c_double, C.c_double, C.c_double, C.c_double, dp, dp]
    rng = np.random.default_rng(7)
    P = lambda a: a.ctypes.data_as(dp)  # noqa: E731
    for n in (1, 2, 3, 5, 8, 13, 1001, 15447):
        tn = np.ascontiguousarray(rng.uniform(1e-3, 2.0, size=(4, n)))
        tc = np.ascontiguousarray(rng.uniform(0.5, 3.0, size=(4, n)))
        x = np.array([0.62, 0.71, 0.83, 1.0])
        fast, scalar = np.empty(n), np.empty(n)
        lib.sh_row_cubic(P(tn), P(tc), n, 0.0123, P(x), 1.0, P(fast), P(scalar))
        assert np.array_equal(fast, scalar)
        lib.sh_row_cubic(P(tn), P(tc), n, 0.0123, P(x), 0.9, P(fast), P(scalar))
        assert np.array_equal(fast, scalar) and np.all(np.isfinite(fast))
        lib.sh_row_linear(P(tn), P(tc), n, 0.0123, 0.83, 1.0, 1.0, P(fast), P(scalar))
        assert np.array_equal(fast, scalar)
