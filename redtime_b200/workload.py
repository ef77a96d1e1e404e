"""Synthetic workloads for bench.py and the batch parity tests (host side, numpy only).

SURVEY.md 8(d) configs 4/5: there is no CAMB in this environment, so every cosmology reuses the
transfer-function set of the reference's examples/1_redTime (committed, packed, under
tests/golden/example1/), multiplied by a smooth seed-derived tilt (k/0.05)^eps, |eps| <= 0.02,
so that no two cosmologies share their tables.  Cosmological parameters are Latin-hypercube
draws in the ranges of the reference's misc/convert_katrin_hypercube.py:5-6, converted to the
params_redTime.dat quantities as scripts/runRedTime:106-110 does.  Physically inconsistent but
identical for this library and for the reference binary, which is what parity and timing need.
"""
import lzma
import os
import tarfile
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLE1 = os.path.join(ROOT, "tests", "golden", "example1")
SEED = 20261018
REDSHIFTS_CE = (2.02, 1.61, 1.01, 0.66, 0.43, 0.24, 0.10, 0.0)  # tests/emulator_comparison/redshifts_ce.txt
# om_m, om_b, s_8, h, n_s, w_0, (-(w0+wa))^1/4, om_nu  (misc/convert_katrin_hypercube.py:5-6)
RANGES_LO = np.array([0.12, 0.0215, 0.7, 0.55, 0.85, -1.3, 0.3, 0.0])
RANGES_HI = np.array([0.155, 0.0235, 0.9, 0.85, 1.05, -0.7, 1.29, 0.01])
INTERP_Z_STR = ("200", "100", "50", "20", "10", "5", "4", "3", "2", "1", ".5", "0")
T_CMB = 2.726


def extract_example1(dst):
    """Materialise the example-1 run directory (params + 12+1 CAMB files) under dst."""
    os.makedirs(dst, exist_ok=True)
    with lzma.open(os.path.join(EXAMPLE1, "camb_transfers.tar.xz")) as f:
        with tarfile.open(fileobj=f) as tar:
            tar.extractall(dst, filter="data")
    with open(os.path.join(EXAMPLE1, "params_redTime.dat")) as f:
        txt = f.read()
    with open(os.path.join(dst, "params_redTime.dat"), "w") as f:
        f.write(txt)
    return dst


def read_run_dir_numpy(path):
    """params_redTime.dat + the 7-column CAMB files it names, parsed with numpy alone (the layout of
    hdr:231-353, 547-627, 790-832).  Used where the product library must not be loaded: the
    reference arm of bench.py.  Same dict as redtime_b200.read_run_dir (tests/test_host_io.py)."""
    with open(os.path.join(path, "params_redTime.dat")) as f:
        tok = [l.split("#")[0].strip() for l in f if l.split("#")[0].strip()]
    params = np.array([float(t.split()[0]) for t in tok[:9]])
    switches = [int(t.split()[0]) for t in tok[9:13]]
    z_in, n_out = float(tok[13].split()[0]), int(tok[14].split()[0])
    z_out = np.array([float(x) for x in tok[15].split()[:n_out]])
    tfile, root, n_z = tok[16].split()[0], tok[18].split()[0], int(tok[19].split()[0])
    zs = tok[20].split()[:n_z]
    t0 = np.loadtxt(os.path.join(path, tfile), usecols=(0, 1, 2))
    tabs = [np.loadtxt(os.path.join(path, "%s%s.dat" % (root, z)), usecols=(0, 1, 5)) for z in zs]
    return dict(params=params, switches=switches, z_in=z_in, z_out=z_out,
                k_T=np.ascontiguousarray(t0[:, 0]), Tc_T=np.ascontiguousarray(t0[:, 1]),
                Tb_T=np.ascontiguousarray(t0[:, 2]), z_interp=np.array([float(z) for z in zs]),
                k_b=np.ascontiguousarray(tabs[0][:, 0]) if tabs else np.zeros(0),
                Tc_b=np.array([t[:, 1] for t in tabs]) if tabs else np.zeros((0, 0)),
                Tnu_b=np.array([t[:, 2] for t in tabs]) if tabs else np.zeros((0, 0)))


def load_example1(subsample=1, use_library=True):
    """The example-1 inputs as the dict RedTimeB200.add_cosmology takes, parsed by the library's
    own C++ reader (use_library=False: by numpy, without loading the library).  subsample > 1
    keeps every subsample-th table row."""
    with tempfile.TemporaryDirectory() as tmp:
        if use_library:
            from .binding import read_run_dir
            d = read_run_dir(extract_example1(tmp))
        else:
            d = read_run_dir_numpy(extract_example1(tmp))
    if subsample > 1:
        sl = slice(None, None, subsample)
        for key in ("k_T", "Tc_T", "Tb_T", "k_b"):
            d[key] = np.ascontiguousarray(d[key][sl])
        for key in ("Tc_b", "Tnu_b"):
            d[key] = np.ascontiguousarray(d[key][:, sl])
    return d


def latin_hypercube(n, seed=SEED):
    rng = np.random.default_rng(seed)
    u = np.empty((n, 8))
    for j in range(8):
        u[:, j] = (rng.permutation(n) + rng.random(n)) / n
    return u, rng.uniform(-0.02, 0.02, size=n)


def _pinned_like(a):
    """Page-locked copy of array a (torch owns the allocation; numpy views it)."""
    import torch
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
    v = t.numpy()
    v[...] = a
    return v


def make_cosmologies(n, base, seed=SEED, switches=(1, 1, 1, 1), z_out=REDSHIFTS_CE, z_in=200.0, pinned=False,
                     total=None):
    """n cosmologies (list of add_cosmology dicts) sharing base's table shapes.  pinned=True puts
    every table in page-locked host memory (needs a CUDA device): the library then sends them to
    the GPU without a host-side copy.  total: size of the Latin-hypercube draw the n cosmologies are
    the FIRST n members of (a Latin hypercube depends on its size) -- this is how the CPU reference
    runs a sample of exactly the cosmologies of a larger GPU batch."""
    u, eps = latin_hypercube(total if total else n, seed)
    out = []
    if pinned:
        base = dict(base)
        for key in ("k_T", "k_b"):
            base[key] = _pinned_like(base[key])
    tilt_T = base["k_T"] / 0.05
    tilt_b = base["k_b"] / 0.05
    for i in range(n):
        v = RANGES_LO + u[i] * (RANGES_HI - RANGES_LO)
        om_m, om_b, s8, h, ns, w0, x, om_nu = v
        wa = -(x ** 4) - w0
        params = np.array([ns, s8, h, om_m / h ** 2, om_b / h ** 2, om_nu / h ** 2, T_CMB, w0, wa])
        tT, tb = tilt_T ** eps[i], tilt_b ** eps[i]
        c = dict(params=params, switches=list(switches), z_in=float(z_in),
                 z_out=np.array(z_out, dtype=float), k_T=base["k_T"],
                 Tc_T=base["Tc_T"] * tT, Tb_T=base["Tb_T"] * tT,
                 z_interp=base["z_interp"], k_b=base["k_b"],
                 Tc_b=base["Tc_b"] * tb[None, :], Tnu_b=base["Tnu_b"] * tb[None, :])
        if pinned:
            for key in ("Tc_T", "Tb_T", "Tc_b", "Tnu_b"):
                c[key] = _pinned_like(c[key])
        out.append(c)
    return out


def _write_camb(path, k, Tc, Tb, Tnu):
    cols = np.zeros((k.size, 7))
    cols[:, 0], cols[:, 1], cols[:, 2], cols[:, 5] = k, Tc, Tb, Tnu
    np.savetxt(path, cols, fmt="%.17e")


def write_run_dir(path, c):
    """Write cosmology dict c as a reference run directory (params_redTime.dat + 7-column
    CAMB files, hdr:231-353,76-80) so that the reference binary reads the identical numbers."""
    os.makedirs(path, exist_ok=True)
    p = c["params"]
    lines = ["# synthetic cosmology written by redtime_b200.workload"]
    lines += ["%.17g" % x for x in p]
    lines += [str(int(s)) for s in c["switches"]]
    lines += ["%.17g" % c["z_in"], str(len(c["z_out"])), " ".join("%.17g" % z for z in c["z_out"])]
    lines += ["camb_transfer_z0.dat", "0", "camb_transfer_z", str(len(INTERP_Z_STR)), " ".join(INTERP_Z_STR)]
    with open(os.path.join(path, "params_redTime.dat"), "w") as f:
        f.write("\n".join(lines) + "\n")
    assert len(INTERP_Z_STR) == c["Tc_b"].shape[0] and np.array_equal(c["k_T"], c["k_b"])
    for iz, zs in enumerate(INTERP_Z_STR[:-1]):
        _write_camb(os.path.join(path, "camb_transfer_z%s.dat" % zs), c["k_b"], c["Tc_b"][iz],
                    np.zeros_like(c["k_b"]), c["Tnu_b"][iz])
    # the z=0 file is both the transfer-function file and the last interpolation file
    assert np.array_equal(c["Tc_T"], c["Tc_b"][-1])
    _write_camb(os.path.join(path, "camb_transfer_z0.dat"), c["k_T"], c["Tc_T"], c["Tb_T"], c["Tnu_b"][-1])
    return path


def perturbed_run_dir(src, dst, line=1, direction=+1):
    """Copy run directory src to dst with ONE value of params_redTime.dat moved by one ulp (line 0 =
    n_s, 1 = sigma_8, ...).  Running the reference on both measures its own round-off floor: how far
    its tables move under a change of the input that is 1e-10 below every tolerance."""
    import shutil
    shutil.copytree(src, dst)
    p = os.path.join(dst, "params_redTime.dat")
    txt = open(p).read().split("\n")
    vals = [n for n, l in enumerate(txt) if l.strip() and not l.startswith("#")]
    x = float(txt[vals[line]].split()[0])
    txt[vals[line]] = "%.17g" % np.nextafter(x, np.inf if direction > 0 else -np.inf)
    open(p, "w").write("\n".join(txt))
    return dst
