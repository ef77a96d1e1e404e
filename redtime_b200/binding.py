"""ctypes binding of include/redtime_b200.h (plumbing only)."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MAX_OUT = 64
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

ERRORS = {-1: "EINVAL", -2: "ENOGPU", -3: "ECUDA", -4: "ERANGE", -5: "EODE", -6: "ENOMEM"}


class RtrgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rtrg error %s (%d): %s" % (ERRORS.get(code, "?"), code, msg))
        self.code = code


class Config(C.Structure):
    """Mirror of rtrg_config (same field order)."""
    _fields_ = [("nk", C.c_int), ("kmin", C.c_double), ("kmax", C.c_double), ("z1l", C.c_double),
                ("eps_abs", C.c_double), ("eps_rel", C.c_double), ("beta_kmin", C.c_double),
                ("beta_kmax", C.c_double), ("n_lnk", C.c_int), ("n_lna", C.c_int),
                ("a_early", C.c_double), ("print_A", C.c_int), ("print_I", C.c_int),
                ("print_Q", C.c_int), ("print_bias", C.c_int), ("device", C.c_int),
                ("max_attempts", C.c_int), ("k_shards", C.c_int), ("k_rank", C.c_int),
                ("v_split", C.c_int), ("reduce_beta", C.c_int)]


class _Cosmology(C.Structure):
    _fields_ = [("params", C.c_double * 9), ("switches", C.c_int * 4), ("z_in", C.c_double),
                ("n_out", C.c_int), ("z_out", _dp), ("n_T", C.c_int), ("k_T", _dp), ("Tc_T", _dp),
                ("Tb_T", _dp), ("n_z", C.c_int), ("z_interp", _dp), ("n_kb", C.c_int), ("k_b", _dp),
                ("Tc_b", _dp), ("Tnu_b", _dp)]


_lib = None


def library_path():
    """The in-tree product library; RTRG_LIBRARY selects another in-tree build of it (the debug
    build with device-side bounds asserts, libredtime_b200_bounds.so)."""
    name = os.environ.get("RTRG_LIBRARY", "libredtime_b200.so")
    return name if os.path.isabs(name) else os.path.join(_HERE, os.path.basename(name))


def load_library():
    """Load the in-tree shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RtrgError(-2, "%s not found: build it with `make -C redtime_b200/csrc` "
                            "(or __graft_entry__.build()); there is no Python/CPU fallback" % path)
    lib = C.CDLL(path)
    lib.rtrg_last_error.restype = C.c_char_p
    lib.rtrg_version.restype = C.c_char_p
    lib.rtrg_launch_count.restype = C.c_longlong
    lib.rtrg_launch_count.argtypes = [C.c_void_p]
    lib.rtrg_matvec_sets.restype = C.c_longlong
    lib.rtrg_matvec_sets.argtypes = [C.c_void_p, C.c_int]
    lib.rtrg_default_config.argtypes = [C.POINTER(Config)]
    lib.rtrg_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    lib.rtrg_destroy.argtypes = [C.c_void_p]
    lib.rtrg_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.rtrg_clear_cosmologies.argtypes = [C.c_void_p]
    lib.rtrg_add_cosmology.argtypes = [C.c_void_p, C.POINTER(_Cosmology)]
    lib.rtrg_add_cosmologies.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.POINTER(_Cosmology))]
    lib.rtrg_num_cosmologies.argtypes = [C.c_void_p]
    lib.rtrg_num_columns.argtypes = [C.c_void_p, C.c_int]
    lib.rtrg_prepare.argtypes = [C.c_void_p]
    lib.rtrg_device_init.argtypes = [C.c_void_p]
    lib.rtrg_set_profiling.argtypes = [C.c_void_p, C.c_int]
    lib.rtrg_profile_name.argtypes = [C.c_int]
    lib.rtrg_profile_name.restype = C.c_char_p
    lib.rtrg_profile_query.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong), _dp]
    lib.rtrg_bench_dfma.argtypes = [C.c_int, C.c_double, _dp]
    lib.rtrg_bench_dmma.argtypes = [C.c_int, C.c_double, _dp]
    lib.rtrg_bench_integrals.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.rtrg_kshard_nccl_id.argtypes = [C.c_char_p]
    lib.rtrg_kshard_init_nccl.argtypes = [C.c_void_p, C.c_char_p]
    lib.rtrg_kshard_loopback_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.rtrg_kshard_init_loopback.argtypes = [C.c_void_p, C.c_void_p]
    lib.rtrg_kshard_loopback_free.argtypes = [C.c_void_p]
    lib.rtrg_kshard_loopback_free.restype = None
    lib.rtrg_run.argtypes = [C.c_void_p, _dp, C.c_size_t, _dp, _dp, _ip]
    lib.rtrg_fetch_outputs.argtypes = [C.c_void_p, C.POINTER(_dp), C.POINTER(C.c_size_t), C.POINTER(_dp),
                                       C.POINTER(_dp)]
    lib.rtrg_counters.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]
    lib.rtrg_extrap_P.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    lib.rtrg_integrals_full.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp]
    lib.rtrg_integrals_raw.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _dp]
    lib.rtrg_derivatives.argtypes = [C.c_void_p, C.c_int, C.c_double, _dp, _dp]
    lib.rtrg_D_dD.argtypes = [C.c_void_p, C.c_int, C.c_double, _dp, C.c_int, _dp, _dp]
    lib.rtrg_Beta_P.argtypes = [C.c_void_p, C.c_int, C.c_double, _dp, C.c_int, _dp]
    lib.rtrg_Plin.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, _dp, C.c_int, _dp]
    lib.rtrg_initial_state.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    lib.rtrg_grid_info.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int * 5, _dp, _dp]
    lib.rtrg_table_T.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, _dp, _dp]
    lib.rtrg_table_G.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, _dp]
    lib.rtrg_table_windows.argtypes = [C.c_int, C.c_double, C.c_double, _dp, _dp]
    lib.rtrg_table_extrap.argtypes = [C.c_int, C.c_double, C.c_double, _ip, _dp, _dp]
    lib.rtrg_assembly_terms.argtypes = [_ip, _ip, _ip, _ip, _dp, C.c_int]
    lib.rtrg_read_run_dir.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.rtrg_read_run_dirs.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_void_p)]
    lib.rtrg_inputs_cosmology.argtypes = [C.c_void_p]
    lib.rtrg_inputs_cosmology.restype = C.POINTER(_Cosmology)
    lib.rtrg_free_run_inputs.argtypes = [C.c_void_p]
    lib.rtrg_print_result.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]
    _lib = lib
    return lib


def _P(a):
    return a.ctypes.data_as(_dp)


def _check(rc):
    if rc != 0:
        raise RtrgError(rc, load_library().rtrg_last_error().decode())


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---- cosmology-independent tables (host side) ------------------------------------------
def grid_info(nk=128, kmin=1e-3, kmax=1.0):
    lib = load_library()
    out = (C.c_int * 5)()
    dl, l0 = C.c_double(), C.c_double()
    _check(lib.rtrg_grid_info(nk, kmin, kmax, out, C.byref(dl), C.byref(l0)))
    return dict(np=out[0], nshift=out[1], jlo=out[2], nsup=out[3], nloMR=out[4], dlnk=dl.value,
                lnk_pad_min=l0.value)


def table_T(n, nk=128, kmin=1e-3, kmax=1.0):
    lib = load_library()
    npad = 4 * nk
    T, kf = np.zeros((npad, npad)), np.zeros(npad)
    _check(lib.rtrg_table_T(nk, kmin, kmax, n, _P(T), _P(kf)))
    return T, kf


def table_G(n, nk=128, kmin=1e-3, kmax=1.0):
    lib = load_library()
    G = np.zeros(8 * nk - 1)
    _check(lib.rtrg_table_G(nk, kmin, kmax, n, _P(G)))
    return G


def table_windows(nk=128, kmin=1e-3, kmax=1.0):
    lib = load_library()
    WP, WC = np.zeros(4 * nk), np.zeros(4 * nk)
    _check(lib.rtrg_table_windows(nk, kmin, kmax, _P(WP), _P(WC)))
    return WP, WC


def table_extrap(nk=128, kmin=1e-3, kmax=1.0):
    """Pab stencil on the padded grid: (n0[np], w[np,4], dx[np]); see rtrg_table_extrap."""
    lib = load_library()
    n0, w, dx = np.zeros(4 * nk, np.int32), np.zeros(16 * nk), np.zeros(4 * nk)
    _check(lib.rtrg_table_extrap(nk, kmin, kmax, n0.ctypes.data_as(_ip), _P(w), _P(dx)))
    return n0, w.reshape(4 * nk, 4), dx


def assembly_terms():
    lib = load_library()
    n = lib.rtrg_assembly_terms(None, None, None, None, None, 0)
    row, src, idx, kpw = (np.zeros(n, np.int32) for _ in range(4))
    cf = np.zeros(n)
    lib.rtrg_assembly_terms(row.ctypes.data_as(_ip), src.ctypes.data_as(_ip), idx.ctypes.data_as(_ip),
                            kpw.ctypes.data_as(_ip), _P(cf), n)
    return row, src, idx, kpw, cf


# ---- run-directory reader (mirrors the reference's params/CAMB parsing) -----------------
def read_run_dir(path, camb_modern=False):
    """Parse <path>/params_redTime.dat + CAMB files through the library's C++ reader and
    return a plain dict of numpy arrays (the argument of RedTimeB200.add_cosmology)."""
    lib = load_library()
    hnd = C.c_void_p()
    rc = lib.rtrg_read_run_dir(os.fsencode(path), int(camb_modern), C.byref(hnd))
    if rc != 0:
        raise RtrgError(rc, "cannot read run directory %s" % path)
    try:
        c = lib.rtrg_inputs_cosmology(hnd).contents
        def arr(p, n):
            return np.ctypeslib.as_array(p, shape=(n,)).copy() if n > 0 else np.zeros(0)
        d = dict(params=np.array(list(c.params)), switches=[int(x) for x in c.switches], z_in=c.z_in,
                 z_out=arr(c.z_out, c.n_out), k_T=arr(c.k_T, c.n_T), Tc_T=arr(c.Tc_T, c.n_T),
                 Tb_T=arr(c.Tb_T, c.n_T), z_interp=arr(c.z_interp, c.n_z), k_b=arr(c.k_b, c.n_kb),
                 Tc_b=arr(c.Tc_b, c.n_z * c.n_kb).reshape(c.n_z, max(c.n_kb, 1)) if c.n_z else np.zeros((0, 0)),
                 Tnu_b=arr(c.Tnu_b, c.n_z * c.n_kb).reshape(c.n_z, max(c.n_kb, 1)) if c.n_z else np.zeros((0, 0)))
    finally:
        lib.rtrg_free_run_inputs(hnd)
    return d


def print_result(path, nk, out, hdr, hdr0, paramfile="params_redTime.dat"):
    """Write one cosmology's tables in the reference's stdout format to `path`."""
    lib = load_library()
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    f = libc.fopen(os.fsencode(path), b"w")
    if not f:
        raise OSError("cannot open %s" % path)
    out = _f64(out)
    n_out, nk_, ncols = out.shape
    hdr, hdr0 = _f64(hdr), _f64(hdr0)
    try:
        _check(lib.rtrg_print_result(f, paramfile.encode() if paramfile else None, nk, ncols, n_out,
                                     _P(out), _P(hdr), _P(hdr0)))
    finally:
        libc.fclose(f)


def kshard_nccl_id():
    """128-byte NCCL unique id (call on rank 0, ship to the other ranks)."""
    lib = load_library()
    buf = C.create_string_buffer(128)
    _check(lib.rtrg_kshard_nccl_id(buf))
    return buf.raw


class LoopbackGroup:
    """In-process transport of the k-sharded mode: one handle per rank, one host thread each."""

    def __init__(self, nranks):
        self.lib = load_library()
        self.g = C.c_void_p()
        _check(self.lib.rtrg_kshard_loopback_create(int(nranks), C.byref(self.g)))

    def close(self):
        if self.g:
            self.lib.rtrg_kshard_loopback_free(self.g)
            self.g = None


class PackedCosmologies:
    """ctypes view of a list of cosmology dicts, reusable across add_cosmologies() calls."""

    def __init__(self, dicts):
        self.structs = [RedTimeB200._struct(d) for d in dicts]
        self.array = (C.POINTER(_Cosmology) * len(self.structs))(*[C.pointer(c) for c, _ in self.structs])
        self.n_out = [int(c.n_out) for c, _ in self.structs]


def pack_cosmologies(dicts):
    return PackedCosmologies(dicts)


class Pipeline:
    """rtrg_pipeline_*: double-buffered batches on one GPU (staging/upload/initialisation of batch
    i+1 overlap the evolution of batch i).  submit() takes PackedCosmologies (kept alive until the
    batch has been waited for) and returns a ticket; wait() returns (tables, hdr, hdr0, status) as
    views of page-locked memory that stay valid until release()."""

    def __init__(self, depth=2, **kw):
        lib = load_library()
        self.lib = lib
        self.cfg = Config()
        lib.rtrg_default_config(C.byref(self.cfg))
        for k, v in kw.items():
            if not hasattr(self.cfg, k):
                raise TypeError("unknown config field %r" % k)
            setattr(self.cfg, k, v)
        lib.rtrg_pipeline_create.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(C.c_void_p)]
        lib.rtrg_pipeline_submit.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.POINTER(_Cosmology)),
                                             C.POINTER(C.c_longlong)]
        lib.rtrg_pipeline_wait.argtypes = [C.c_void_p, C.c_longlong, C.POINTER(_dp), C.POINTER(C.c_size_t),
                                           C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_ip)]
        lib.rtrg_pipeline_columns.argtypes = [C.c_void_p, C.c_longlong, C.c_int]
        lib.rtrg_pipeline_release.argtypes = [C.c_void_p, C.c_longlong]
        lib.rtrg_pipeline_times.argtypes = [C.c_void_p, C.c_longlong, _dp]
        lib.rtrg_pipeline_destroy.argtypes = [C.c_void_p]
        self.p = C.c_void_p()
        _check(lib.rtrg_pipeline_create(C.byref(self.cfg), int(depth), C.byref(self.p)))
        self.nk = self.cfg.nk
        self._jobs = {}

    def submit(self, packed):
        if not isinstance(packed, PackedCosmologies):
            packed = PackedCosmologies(packed)
        t = C.c_longlong()
        _check(self.lib.rtrg_pipeline_submit(self.p, len(packed.structs), packed.array, C.byref(t)))
        self._jobs[t.value] = packed
        return t.value

    def wait(self, ticket, raise_on_failure=True):
        packed = self._jobs[ticket]
        po, ph, ph0, ps, n = _dp(), _dp(), _dp(), _ip(), C.c_size_t()
        rc = self.lib.rtrg_pipeline_wait(self.p, ticket, C.byref(po), C.byref(n), C.byref(ph), C.byref(ph0),
                                         C.byref(ps))
        if rc != 0 and raise_on_failure:
            _check(rc)
        B = len(packed.structs)
        out = np.ctypeslib.as_array(po, shape=(n.value,))
        hdr = np.ctypeslib.as_array(ph, shape=(B, MAX_OUT, 5))
        hdr0 = np.ctypeslib.as_array(ph0, shape=(B, 2))
        status = np.ctypeslib.as_array(ps, shape=(B,))
        tables, o = [], 0
        for i in range(B):
            ncols = self.lib.rtrg_pipeline_columns(self.p, ticket, i)
            sz = packed.n_out[i] * self.nk * ncols
            tables.append(out[o:o + sz].reshape(packed.n_out[i], self.nk, ncols))
            o += sz
        return tables, hdr, hdr0, status

    def times(self, ticket):
        """(stage0, stage1, run0, run1, fetch0, fetch1) of a finished job, seconds since creation."""
        t = (C.c_double * 6)()
        _check(self.lib.rtrg_pipeline_times(self.p, ticket, t))
        return tuple(t)

    def release(self, ticket):
        _check(self.lib.rtrg_pipeline_release(self.p, ticket))
        self._jobs.pop(ticket, None)

    def close(self):
        if getattr(self, "p", None) and self.p:
            self.lib.rtrg_pipeline_destroy(self.p)
            self.p = None

    __del__ = close


def dfma_peak_tflops(device=0, seconds=0.5):
    """Measured rate of a register-resident scalar DFMA loop (TFLOP/s)."""
    lib = load_library()
    t = C.c_double()
    _check(lib.rtrg_bench_dfma(int(device), float(seconds), C.byref(t)))
    return t.value


def dmma_peak_tflops(device=0, seconds=0.5):
    """Measured FP64 pipe peak of the device (TFLOP/s): a register-resident DMMA.8x8x4 loop, the
    instruction the bilinear kernel runs on."""
    lib = load_library()
    t = C.c_double()
    _check(lib.rtrg_bench_dmma(int(device), float(seconds), C.byref(t)))
    return t.value


class RedTimeB200:
    """One handle = one GPU + one (nk,kmin,kmax) grid; cosmologies are batched."""

    def __init__(self, **kw):
        lib = load_library()
        self.lib = lib
        self.cfg = Config()
        lib.rtrg_default_config(C.byref(self.cfg))
        for k, v in kw.items():
            if not hasattr(self.cfg, k):
                raise TypeError("unknown config field %r" % k)
            setattr(self.cfg, k, v)
        self.h = C.c_void_p()
        _check(lib.rtrg_create(C.byref(self.cfg), C.byref(self.h)))
        self.nk = self.cfg.nk
        self._keep = []
        self._nout = []

    def close(self):
        if getattr(self, "h", None) and self.h:
            self.lib.rtrg_destroy(self.h)
            self.h = None

    __del__ = close

    def set_stream(self, cuda_stream):
        _check(self.lib.rtrg_set_stream(self.h, C.c_void_p(cuda_stream)))

    def clear(self):
        _check(self.lib.rtrg_clear_cosmologies(self.h))
        self._keep, self._nout = [], []

    @staticmethod
    def _struct(d):
        c = _Cosmology()
        for i in range(9):
            c.params[i] = float(d["params"][i])
        for i in range(4):
            c.switches[i] = int(d["switches"][i])
        c.z_in = float(d["z_in"])
        keep = {k: _f64(d[k]) for k in ("z_out", "k_T", "Tc_T", "Tb_T", "z_interp", "k_b", "Tc_b", "Tnu_b")}
        c.n_out = keep["z_out"].size
        c.z_out = _P(keep["z_out"])
        c.n_T = keep["k_T"].size
        c.k_T, c.Tc_T, c.Tb_T = _P(keep["k_T"]), _P(keep["Tc_T"]), _P(keep["Tb_T"])
        c.n_z = keep["z_interp"].size
        c.z_interp = _P(keep["z_interp"])
        c.n_kb = keep["k_b"].size
        c.k_b, c.Tc_b, c.Tnu_b = _P(keep["k_b"]), _P(keep["Tc_b"]), _P(keep["Tnu_b"])
        return c, keep

    def add_cosmology(self, d):
        """d: dict with params[9], switches[4], z_in, z_out, k_T, Tc_T, Tb_T, z_interp, k_b,
        Tc_b[n_z][n_kb], Tnu_b[n_z][n_kb] (the library copies the arrays)."""
        c, keep = self._struct(d)
        _check(self.lib.rtrg_add_cosmology(self.h, C.byref(c)))
        self._nout.append(int(c.n_out))
        return len(self._nout) - 1

    def add_run_dirs(self, paths, camb_modern=False):
        """Parse the run directories (params_redTime.dat + CAMB files) on the host threads of the
        library and add them as one batch; no intermediate numpy copies."""
        n = len(paths)
        cpaths = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
        hnd = (C.c_void_p * n)()
        rc = self.lib.rtrg_read_run_dirs(n, cpaths, int(camb_modern), hnd)
        if rc != 0:
            raise RtrgError(rc, "cannot read one of the run directories")
        try:
            ptrs = [self.lib.rtrg_inputs_cosmology(hnd[i]) for i in range(n)]
            arr = (C.POINTER(_Cosmology) * n)(*ptrs)
            _check(self.lib.rtrg_add_cosmologies(self.h, n, arr))
            self._nout.extend(int(p.contents.n_out) for p in ptrs)
        finally:
            for i in range(n):
                self.lib.rtrg_free_run_inputs(hnd[i])

    def add_cosmologies(self, dicts):
        """Batch form of add_cosmology: one C-ABI call, table copies on several host threads
        (page-locked arrays are sent to the device directly, see include/redtime_b200.h).
        `dicts` may also be the result of pack_cosmologies(), which skips the struct building."""
        pack = dicts if isinstance(dicts, PackedCosmologies) else pack_cosmologies(dicts)
        _check(self.lib.rtrg_add_cosmologies(self.h, len(pack.structs), pack.array))
        self._nout.extend(pack.n_out)
        self._keep = [pack]  # page-locked buffers must outlive rtrg_prepare

    @property
    def n_cosmo(self):
        return self.lib.rtrg_num_cosmologies(self.h)

    def num_columns(self, i=0):
        return self.lib.rtrg_num_columns(self.h, i)

    def prepare(self):
        _check(self.lib.rtrg_prepare(self.h))

    def device_init(self):
        """Device-side initialisation alone, on inputs already resident in HBM."""
        _check(self.lib.rtrg_device_init(self.h))

    def run_resident(self):
        """Evolve all cosmologies; tables stay on the device (no D2H).  Returns status[]."""
        status = np.zeros(self.n_cosmo, np.int32)
        _check(self.lib.rtrg_run(self.h, None, 0, None, None, status.ctypes.data_as(_ip)))
        return status

    def kshard_init_nccl(self, unique_id):
        _check(self.lib.rtrg_kshard_init_nccl(self.h, unique_id))

    def kshard_init_auto(self, dist):
        """k-shard transport for one process per GPU under torch.distributed: rank 0's NCCL unique id
        travels through the process group; returns a description of the transport in use."""
        import torch
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if dist.get_rank() == 0:
            idt.copy_(torch.frombuffer(bytearray(kshard_nccl_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        self.kshard_init_nccl(bytes(idt.cpu().numpy().tobytes()))
        return self.kshard_transport()

    def kshard_transport(self):
        self.lib.rtrg_kshard_transport.restype = C.c_char_p
        self.lib.rtrg_kshard_transport.argtypes = [C.c_void_p]
        return self.lib.rtrg_kshard_transport(self.h).decode()

    def kshard_init_loopback(self, group):
        _check(self.lib.rtrg_kshard_init_loopback(self.h, group.g))

    def bench_integrals(self, reps=1, groups=31, identical=0):
        _check(self.lib.rtrg_bench_integrals(self.h, int(reps), int(groups), int(identical)))

    def set_profiling(self, on=True):
        _check(self.lib.rtrg_set_profiling(self.h, int(on)))

    def profile(self):
        """{kernel name: (launches, total device ms)} since profiling was switched on."""
        out = {}
        for cat in range(self.lib.rtrg_profile_categories()):
            n, ms = C.c_longlong(), C.c_double()
            self.lib.rtrg_profile_query(self.h, cat, C.byref(n), C.byref(ms))
            out[self.lib.rtrg_profile_name(cat).decode()] = (int(n.value), float(ms.value))
        return out

    def run_pinned(self, raise_on_ode_failure=True):
        """Like run(), but the returned arrays are views of page-locked buffers owned by the
        handle (no pageable copy); they are valid until the next prepare()/run_pinned()."""
        B = self.n_cosmo
        status = np.zeros(B, np.int32)
        rc = self.lib.rtrg_run(self.h, None, 0, None, None, status.ctypes.data_as(_ip))
        if rc != 0 and (rc != -5 or raise_on_ode_failure):
            _check(rc)
        po, ph, ph0, n = _dp(), _dp(), _dp(), C.c_size_t()
        _check(self.lib.rtrg_fetch_outputs(self.h, C.byref(po), C.byref(n), C.byref(ph), C.byref(ph0)))
        out = np.ctypeslib.as_array(po, shape=(n.value,))
        hdr = np.ctypeslib.as_array(ph, shape=(B, MAX_OUT, 5))
        hdr0 = np.ctypeslib.as_array(ph0, shape=(B, 2))
        tables, o = [], 0
        for i in range(B):
            sz = self._nout[i] * self.nk * self.num_columns(i)
            tables.append(out[o:o + sz].reshape(self._nout[i], self.nk, -1))
            o += sz
        return tables, hdr, hdr0, status

    def run(self, raise_on_ode_failure=True):
        """Returns (tables, hdr, hdr0, status): tables[i] has shape [n_out_i, nk, ncols_i]."""
        B = self.n_cosmo
        ncols = [self.num_columns(i) for i in range(B)]
        sizes = [self._nout[i] * self.nk * ncols[i] for i in range(B)]
        out = np.zeros(int(sum(sizes)))
        hdr = np.zeros((B, MAX_OUT, 5))
        hdr0 = np.zeros((B, 2))
        status = np.zeros(B, np.int32)
        rc = self.lib.rtrg_run(self.h, _P(out), out.size, _P(hdr), _P(hdr0), status.ctypes.data_as(_ip))
        if rc != 0 and (rc != -5 or raise_on_ode_failure):
            _check(rc)
        tables, o = [], 0
        for i in range(B):
            tables.append(out[o:o + sizes[i]].reshape(self._nout[i], self.nk, ncols[i]))
            o += sizes[i]
        return tables, hdr, hdr0, status

    def counters(self, i=0):
        c = (C.c_longlong * 4)()
        _check(self.lib.rtrg_counters(self.h, i, c))
        return dict(attempts=c[0], rejected=c[1], rhs=c[2], integral_evals=c[3])

    def rmax_history(self, i=0):
        """Error norms of the RKF45 attempts of cosmology i in the last run (rtrg_rmax_history)."""
        buf = np.zeros(128)
        self.lib.rtrg_rmax_history.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int]
        n = self.lib.rtrg_rmax_history(self.h, i, _P(buf), 128)
        if n < 0:
            _check(n)
        return buf[:n].copy()

    def matvec_sets(self, i=0):
        return int(self.lib.rtrg_matvec_sets(self.h, i))

    def launch_count(self):
        return int(self.lib.rtrg_launch_count(self.h))

    # ---- stage-level hooks (reference layouts) -----------------------------------------
    def extrap_P(self, lnP, i=0):
        lnP = _f64(lnP)
        P = np.zeros(3 * 4 * self.nk)
        _check(self.lib.rtrg_extrap_P(self.h, i, _P(lnP), _P(P)))
        return P.reshape(3, 4 * self.nk)

    def integrals_full(self, lnP, i=0):
        lnP = _f64(lnP)
        nk = self.nk
        A, R, PT, PMR = np.zeros(64 * nk), np.zeros(24 * nk), np.zeros(9 * nk), np.zeros(8 * nk)
        _check(self.lib.rtrg_integrals_full(self.h, i, _P(lnP), _P(A), _P(R), _P(PT), _P(PMR)))
        return A.reshape(64, nk), R.reshape(24, nk), PT.reshape(9, nk), PMR.reshape(8, nk)

    def integrals_raw(self, lnP, i=0):
        lnP = _f64(lnP)
        nk = self.nk
        J, PZ, J0 = np.zeros(63 * nk), np.zeros(63 * nk), np.zeros(63 * nk)
        jlo = C.c_double()
        _check(self.lib.rtrg_integrals_raw(self.h, i, _P(lnP), _P(J), _P(PZ), _P(J0), C.byref(jlo)))
        return J.reshape(63, nk), PZ.reshape(63, nk), J0.reshape(63, nk), jlo.value

    def derivatives(self, eta, y, i=0):
        y = _f64(y)
        dy = np.zeros_like(y)
        _check(self.lib.rtrg_derivatives(self.h, i, float(eta), _P(y), _P(dy)))
        return dy

    def D_dD(self, z, k, i=0):
        k = _f64(k)
        D, dD = np.zeros_like(k), np.zeros_like(k)
        _check(self.lib.rtrg_D_dD(self.h, i, float(z), _P(k), k.size, _P(D), _P(dD)))
        return D, dD

    def Beta_P(self, a, k, i=0):
        k = _f64(k)
        b = np.zeros_like(k)
        _check(self.lib.rtrg_Beta_P(self.h, i, float(a), _P(k), k.size, _P(b)))
        return b

    def Plin(self, which, z, k, i=0):
        k = _f64(k)
        P = np.zeros_like(k)
        _check(self.lib.rtrg_Plin(self.h, i, int(which), float(z), _P(k), k.size, _P(P)))
        return P

    def initial_state(self, i=0):
        y = np.zeros(41 * self.nk)
        s = np.zeros(2)
        _check(self.lib.rtrg_initial_state(self.h, i, _P(y), _P(s)))
        return y, s
