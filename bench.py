#!/usr/bin/env python3
"""bench.py -- cosmology*redshift outputs per second of the Time-RG hot path at nk=128.

  python bench.py [--gpus N] [--steps K] [--warmup W]              this library on N B200s
  python bench.py --impl reference [--gpus N] [--steps K] ...      the reference on host cores

A "step" is one pass of the whole hot path over one batch of synthetic cosmologies per GPU
(device-side linear theory tables + sigma_8 normalisation + 1-loop mode-coupling integrals +
Time-RG evolution with the batched RKF45 stepper + output tables at 8 redshifts).

  value : outputs/s with the inputs already resident in HBM (CUDA events on the library's
          stream, max over ranks; the production launch path: the evolution is one conditional
          WHILE graph); work of all ranks / that time ("weak" scaling: every rank owns its own
          --cosmologies batch, no data-path collective).
  e2e   : the same through the reference-facing C-ABI with HOST buffers, every step moving its
          inputs host->device and its tables device->host: rtrg_pipeline_* (double-buffered:
          batch i+1 is staged, uploaded and initialised while batch i evolves).
  roofline     : the dominant kernel (k_bilinear, FP64 FMA pipe): algorithmic FLOP of all its
                 launches / its CUDA-event time (extra steps with per-kernel event timing),
                 against the FP64 pipe peak (DMMA loop) measured live.
  cpu_baseline : oracle/_ref/redTime (the UNMODIFIED reference sources + mini-GSL shim) as one
                 single-thread process per host core on the FIRST cosmologies of rank 0's batch.
  parity       : those oracle tables against the GPU tables of the same cosmologies (taken from the
                 end-to-end path), outside every timed region.
  modes / kshard : short sub-records of the other BASELINE configurations -- full Time-RG
                 (switches 1 0 1 1, what scripts/runRedTime:101 writes), nk=256, and ONE nk=256
                 high-accuracy cosmology with its k rows sharded over the GPUs (strong scaling).
"""
import argparse
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the harness keeps the weight-table cache inside the repository (the library default is ~/.cache)
os.environ.setdefault("RTRG_CACHE_DIR", os.path.join(ROOT, ".rtrg_cache"))

METRIC = "cosmology*redshift outputs/sec at nk=128"
UNIT = "outputs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cosmologies", type=int, default=1024, help="cosmologies per GPU and step")
    ap.add_argument("--mode", default="1loop", choices=["1loop", "full"],
                    help="1loop = switches 1 1 1 1 (headline); full = 1 0 1 1 (full Time-RG)")
    ap.add_argument("--subsample", type=int, default=1, help="keep every n-th CAMB table row")
    ap.add_argument("--nk", type=int, default=128, help="output wavenumbers (the headline metric is nk=128)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-modes", action="store_true", help="skip the full-TRG / nk=256 / k-shard sub-records")
    ap.add_argument("--print-all", action="store_true",
                    help="PRINTA = PRINTI = PRINTQ = PRINTBIAS = 1: 84 columns per row (BASELINE configs[3])")
    ap.add_argument("--full-beta", action="store_true",
                    help="upload the full Beta_P(a,k) tables (default: the host pre-reduces them to what the run "
                         "consumes, rtrg_config.reduce_beta = 1)")
    ap.add_argument("--pageable", action="store_true",
                    help="keep the synthetic input tables in pageable host memory (the library then stages "
                         "them through its own page-locked arena); default: page-locked inputs, direct H2D")
    ap.add_argument("--ref-procs", type=int, default=0, help="reference arm: concurrent processes (0 = host cores)")
    ap.add_argument("--v-split", type=int, default=0, help="kshard: rtrg_config.v_split (0 = by GPU count)")
    ap.add_argument("--workload", default="batch", choices=["batch", "kshard"],
                    help="batch: cosmologies sharded over GPUs, no collective (headline, with the other "
                         "configurations as sub-records); kshard: only the k-sharded single cosmology")
    return ap.parse_args()


def workload_name(B, nk, mode, subsample=1):
    sw = "1 1 1 1 (1-loop)" if mode == "1loop" else "1 0 1 1 (full Time-RG)"
    return ("throughput sweep (BASELINE configs[4]): %d w0wa+massive-nu cosmologies per GPU x 8 redshifts, "
            "nk=%d, switches %s, example-1 CAMB tables (%d rows x 12 redshifts) with per-cosmology tilt, "
            "Latin-hypercube parameters seed 20261018" % (B, nk, sw, -(-15447 // subsample)))


def config_dict(a, B, l2_inputs_gb):
    """`config` of the JSON line; the reference arm prints the same dict (same workload)."""
    return {"workload": workload_name(B, a.nk, a.mode, a.subsample), "cosmologies_per_gpu": B, "redshifts": 8,
            "nk": a.nk, "mode": a.mode, "columns": 84 if a.print_all else 17,
            "l2": "256 MiB flush before every step; per-step inputs %.2f GB > L2" % l2_inputs_gb}


def metric_name(nk):
    return METRIC if nk == 128 else METRIC.replace("nk=128", "nk=%d" % nk)


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference binary on the host cores
# ------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def core_list():
    try:
        return sorted(os.sched_getaffinity(0))
    except AttributeError:
        return list(range(os.cpu_count() or 1))


def reference_binary(nk=128, hiacc=False):
    """The reference executable: nk is a compile-time constant there (redTime.cc:93)."""
    name = "redTime_hiacc" if hiacc else {128: "redTime", 256: "redTime_nk256"}.get(nk)
    if name is None:
        return None
    p = os.path.join(ROOT, "oracle", "_ref", name)
    return p if os.path.exists(p) else None


def reference_spawn(dirs, binary, cores, omp_threads=1):
    """One process per run directory, `omp_threads` OpenMP threads each, pinned round-robin to
    `cores`; returns the Popen list (stdout -> <dir>/out.dat)."""
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(omp_threads)
    use_taskset = shutil.which("taskset") is not None and omp_threads == 1
    procs = []
    for i, d in enumerate(dirs):
        cmd = [binary]
        if use_taskset:
            cmd = ["taskset", "-c", str(cores[i % len(cores)])] + cmd
        procs.append(subprocess.Popen(cmd, cwd=d, env=env, stdout=open(os.path.join(d, "out.dat"), "w"),
                                      stderr=subprocess.DEVNULL))
    return procs


def reference_step(dirs, binary, cores=None, omp_threads=1):
    """Wall seconds from first spawn to last exit."""
    t0 = time.perf_counter()
    rcs = [p.wait() for p in reference_spawn(dirs, binary, cores or core_list(), omp_threads)]
    wall = time.perf_counter() - t0
    if any(rcs):
        raise RuntimeError("reference process failed: %s" % rcs)
    return wall


def reference_dirs(n, total, mode, subsample, tmp, use_library):
    """Run directories of the FIRST n cosmologies of the `total`-member draw rank 0 benchmarks."""
    from redtime_b200 import workload as wl
    base = wl.load_example1(subsample, use_library=use_library)
    sw = (1, 1, 1, 1) if mode == "1loop" else (1, 0, 1, 1)
    cosmos = wl.make_cosmologies(n, base, seed=wl.SEED, switches=sw, total=total)
    return [wl.write_run_dir(os.path.join(tmp, "c%04d" % i), c) for i, c in enumerate(cosmos)]


def read_tables(d, nk):
    rows = [l.split() for l in open(os.path.join(d, "out.dat")) if l.strip() and not l.startswith("#")]
    arr = np.array(rows, dtype=float)
    return arr.reshape(-1, nk, arr.shape[1])


def run_reference_arm(a, rank):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, on
    this arm's config.  No product library in this process: the inputs are parsed with numpy."""
    if rank != 0:
        return
    steps = a.steps if a.steps is not None else 2
    warm = a.warmup if a.warmup is not None else 1
    binary = reference_binary(a.nk)
    if binary is None:
        emit(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/redTime not built (make -C oracle)"}))
        return
    nproc = a.ref_procs or host_cores()
    with tempfile.TemporaryDirectory() as tmp:
        dirs = reference_dirs(nproc, a.cosmologies, a.mode, a.subsample, tmp, use_library=False)
        # bytes of the two interpolation tables of one cosmology x the batch (what the GPU arm's config states)
        n_rows = sum(1 for l in open(os.path.join(dirs[0], "camb_transfer_z0.dat")) if l.strip() and not l.startswith("#"))
        l2_gb = a.cosmologies * 2 * 12 * n_rows * 8 / 1e9
        for _ in range(warm):
            reference_step(dirs, binary)
        t = [reference_step(dirs, binary) for _ in range(steps)]
    n_out = 8
    wall = float(np.sum(t))
    value = steps * nproc * n_out / wall
    sample = ("the first %d cosmologies of the %d-member batch x %d redshifts per step, one single-thread process "
              "per core" % (nproc, a.cosmologies, n_out))
    line = {"impl": "reference", "metric": metric_name(a.nk), "value": value, "unit": UNIT, "n_gpus": a.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * wall / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(a, a.cosmologies, l2_gb),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(json.dumps(line))


def cpu_baseline_headline(a, gpu_tables):
    """Bounded sample for the default bench line: ONE step of the reference on the first `cores`
    cosmologies of rank 0's batch, with the shipped 15 447-row CAMB tables and with the scripts'
    default density (every 27th row, ~580 rows).  The 15 447-row tables are then compared with the
    GPU tables of the SAME cosmologies (parity block); an untimed second reference run with sigma_8
    moved by one ulp measures how far the reference's own tables move (its round-off floor)."""
    from redtime_b200 import workload as wl
    binary = reference_binary(a.nk)
    if binary is None:
        return None, None
    nproc = host_cores()
    n_out = 8
    with tempfile.TemporaryDirectory() as tmp:
        dirs = reference_dirs(nproc, a.cosmologies, a.mode, a.subsample, os.path.join(tmp, "full"), True)
        wall = reference_step(dirs, binary)
        ref = [read_tables(d, a.nk) for d in dirs]
        floor = None
        if gpu_tables is not None:
            # sigma_8 + 1 ulp and n_s - 1 ulp: two samples of the reference's own round-off response
            pa = [wl.perturbed_run_dir(d, os.path.join(tmp, "pa%04d" % i), line=1, direction=+1) for i, d in enumerate(dirs)]
            pb = [wl.perturbed_run_dir(d, os.path.join(tmp, "pb%04d" % i), line=0, direction=-1) for i, d in enumerate(dirs)]
            reference_step(pa + pb, binary)
            floor = [np.maximum(np.abs(read_tables(x, a.nk) - r), np.abs(read_tables(y, a.nk) - r))
                     for x, y, r in zip(pa, pb, ref)]
        sparse = None
        if a.subsample == 1:
            d27 = reference_dirs(nproc, a.cosmologies, a.mode, 27, os.path.join(tmp, "s27"), True)
            w27 = reference_step(d27, binary)
            sparse = {"value": nproc * n_out / w27, "unit": UNIT, "cores": nproc,
                      "sample": "same cosmologies with every 27th CAMB row (572 rows, the density scripts/runRedTime "
                                "produces), %.1f s wall" % w27}
    cpu = {"value": nproc * n_out / wall, "unit": UNIT, "cores": nproc, "kind": "reference",
           "sample": "the first %d cosmologies of rank 0's batch x %d redshifts (one single-thread oracle/_ref/redTime "
                     "process per core), %.1f s wall" % (nproc, n_out, wall),
           "with_580_row_tables": sparse}
    parity = None
    if gpu_tables is not None:
        n = min(len(ref), len(gpu_tables))
        e17 = e810 = ehi = f810 = excess = 0.0
        n_unstable = 0
        for r, g, fl in zip(ref[:n], gpu_tables[:n], floor[:n]):
            d = np.abs(g - r)
            rel = d / (np.abs(r) + 1e-300)
            e17 = max(e17, float(rel[:, :, :7].max()))
            e810 = max(e810, float(rel[:, :, 7:10].max()))
            fr = float((fl[:, :, 7:10] / (np.abs(r[:, :, 7:10]) + 1e-300)).max())
            f810 = max(f810, fr)
            n_unstable += fr > 1e-5
            # columns 11-17: relative to the local scale (sign changes of P_B,j)
            a_ = np.abs(r)
            scale = a_.copy()
            for sh in (1, 2):
                scale[:, sh:] = np.maximum(scale[:, sh:], a_[:, :-sh])
                scale[:, :-sh] = np.maximum(scale[:, :-sh], a_[:, sh:])
            hi = r[0, :, 0] > 5.7e-3
            ehi = max(ehi, float((d[:, hi, 10:] / (scale[:, hi, 10:] + 1e-300)).max()))
            # the all-k criterion of the tests: tolerance + 5 x the floor (widened over z and k +- 1)
            fs = np.max(fl, axis=0, keepdims=True) * np.ones_like(fl)
            fw = fs.copy()
            fw[:, 1:] = np.maximum(fw[:, 1:], fs[:, :-1])
            fw[:, :-1] = np.maximum(fw[:, :-1], fs[:, 1:])
            allowed = np.empty_like(d)
            allowed[..., :7] = 1e-6 * a_[..., :7]
            allowed[..., 7:10] = 1e-5 * a_[..., 7:10] + 5 * fw[..., 7:10]
            allowed[..., 10:] = 1e-5 * scale[..., 10:] + 5 * fw[..., 10:]
            excess = max(excess, float((d / (allowed + 1e-300)).max()))
        parity = {"cols_1_7": e17, "cols_8_10": e810, "cols_11_17_hi_k": ehi, "n_compared": n,
                  "max_error_over_allowed_all_k": excess, "pass": bool(excess <= 1.0),
                  "reference_1ulp_response_cols_8_10": f810, "n_reference_unstable": int(n_unstable),
                  "what": "max relative error of the GPU tables (end-to-end path, reduce_beta=%d) against oracle/_ref/"
                          "redTime on the same %d cosmologies; tolerance 1e-6 (columns 1-7) / 1e-5 (8-10) / 1e-5 of the "
                          "local scale (11-17; cols_11_17_hi_k: k > 5.7e-3 h/Mpc).  max_error_over_allowed_all_k: every "
                          "column at EVERY k against tolerance + 5 x the reference's own response to a 1-ulp change of "
                          "sigma_8 or n_s (two more oracle runs per cosmology); n_reference_unstable counts the cosmologies whose "
                          "reference P(k) itself moves by more than 1e-5 under that change (GSL's step controller at an "
                          "accept/reject boundary) -- cols_8_10 is dominated by those" % (0 if a.full_beta else 1, n)}
    return cpu, parity


# ------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])), mx.append(float(r[1])), pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# this library
# ------------------------------------------------------------------------------------------
def flops_per_matvec_set(grid, nk):
    """Algorithmic FLOP of one (kernel n, beta-side spectrum) set of k_bilinear for one
    cosmology: for each of the nk rows an nsup x nsup matrix-vector product
    (2 nsup^2 FLOP), then 3 alpha-side dot products (2 nsup each).  A full evaluation of all
    outputs is 42 sets; the library computes only the sets the requested outputs consume."""
    nsup = grid["nsup"]
    return nk * (2.0 * nsup * nsup + 3 * 2.0 * nsup)


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_bilinear launch, parsed from the
    committed ncu --set full capture of this kernel (profiles/r02_k_bilinear_ncu.csv)."""
    path = os.path.join(ROOT, "profiles", "r02_k_bilinear_ncu.csv")
    try:
        vals = {}
        for line in open(path):
            f = [x.strip().strip('"') for x in line.split(",")]
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[f[1]]
                vals[f[0]] = float(f[2]) * mult
        if len(vals) == 2:
            return int(sum(vals.values())), path
    except (OSError, ValueError, KeyError):
        pass
    return None, path


class Comm:
    """torch.distributed plumbing of the benchmark itself (barriers, max/sum of timings): the batch
    data path has no collective."""

    def __init__(self, world, local_rank):
        import torch
        self.torch, self.world = torch, world
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist

    def sync(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def reduce(self, x, op="max"):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())


def measure_batch(a, comm, rank, world, local_rank, B, nk, mode, steps, warm, want_e2e, prof_steps, clocks=None,
                  keep_tables=0):
    """One batch workload on every rank: resident-input timing (profiling off: the production launch
    path), kernel shares from extra profiled steps, end-to-end timing through the pipeline."""
    import torch
    import redtime_b200 as rt
    from redtime_b200 import workload as wl

    base = wl.load_example1(a.subsample)
    sw = (1, 1, 1, 1) if mode == "1loop" else (1, 0, 1, 1)
    # rank 0 benchmarks the draw whose first members the CPU reference runs
    cosmos = wl.make_cosmologies(B, base, seed=wl.SEED + 7919 * rank, switches=sw, pinned=not a.pageable)
    packed = rt.pack_cosmologies(cosmos)  # ctypes views of the same buffers, built once
    n_out = len(wl.REDSHIFTS_CE)
    outputs_per_step = B * n_out
    reduce_beta = 0 if a.full_beta else 1
    extra = dict(print_A=1, print_I=1, print_Q=1, print_bias=1) if a.print_all else {}
    cfg = dict(device=local_rank, nk=nk, reduce_beta=reduce_beta, **extra)
    h = rt.RedTimeB200(**cfg)
    stream = torch.cuda.Stream()
    h.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def resident_step():
        with torch.cuda.stream(stream):
            flush.zero_()
        h.device_init()
        return h.run_resident()

    # ---- resident-input timing ("value")
    h.clear()
    h.add_cosmologies(packed)
    h.prepare()
    st = None
    for _ in range(max(warm, 1)):
        st = resident_step()
    n_failed = int(np.count_nonzero(st))
    l0 = h.launch_count()
    comm.sync()
    if clocks is not None:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        st = resident_step()
    ev1.record(stream)
    comm.sync()
    clk = clocks.stop() if clocks is not None else None
    ms_total = comm.reduce(ev0.elapsed_time(ev1))
    launches = h.launch_count() - l0
    n_failed = max(n_failed, int(np.count_nonzero(st)))
    value = world * outputs_per_step * steps / (ms_total * 1e-3)

    # ---- per-kernel CUDA-event timing (switches the graph replay off): extra, separately timed steps
    h.set_profiling(True)
    comm.sync()
    pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pv0.record(stream)
    for _ in range(prof_steps):
        resident_step()
    pv1.record(stream)
    comm.sync()
    ms_prof = pv0.elapsed_time(pv1)
    prof = h.profile()
    h.set_profiling(False)
    evals = sum(h.counters(i)["integral_evals"] for i in range(B))  # per step (last run)
    sets = sum(h.matvec_sets(i) for i in range(B))                  # per step (device_init + run)
    h.close()
    del flush

    # ---- end to end through the C-ABI with host buffers (double-buffered pipeline)
    e2e, tables_keep = None, None
    if want_e2e:
        pipe = rt.Pipeline(depth=2, **cfg)
        for _ in range(2):  # warm both handles' arenas
            t = pipe.submit(packed)
            pipe.wait(t)
            pipe.release(t)
        comm.sync()
        t0 = time.perf_counter()
        tickets = [pipe.submit(packed) for _ in range(min(2, steps))]
        check = []
        for i in range(steps):
            tables, hdr, hdr0, status = pipe.wait(tickets[i], raise_on_failure=False)
            check.append(float(tables[B // 2][-1, nk // 2, 7]))   # the host reads the step's result
            n_failed = max(n_failed, int(np.count_nonzero(status)))
            if os.environ.get("RTRG_PIPE_TIMES") and rank == 0:
                print("pipe job %d: stage %.1f-%.1f run %.1f-%.1f fetch %.1f-%.1f ms" %
                      ((i,) + tuple(1e3 * x for x in pipe.times(tickets[i]))), file=sys.stderr)
            if i == steps - 1:
                d2h = sum(t.nbytes for t in tables) + hdr.nbytes + hdr0.nbytes
                if keep_tables:
                    tables_keep = [t.copy() for t in tables[:keep_tables]]
            pipe.release(tickets[i])
            if i + 2 < steps:
                tickets.append(pipe.submit(packed))
        t_pipe = comm.reduce(time.perf_counter() - t0)
        pipe.close()
        assert len(set(check)) == 1, "pipelined batches differ"
        # serial reference point: one handle, nothing overlapped
        hs = rt.RedTimeB200(**cfg)
        for _ in range(1):
            hs.clear(), hs.add_cosmologies(packed), hs.prepare(), hs.run_pinned()
        comm.sync()
        n_ser = min(steps, 3)
        t0 = time.perf_counter()
        for _ in range(n_ser):
            hs.clear()
            hs.add_cosmologies(packed)
            hs.prepare()
            hs.run_pinned()
        t_serial = comm.reduce(time.perf_counter() - t0) / n_ser
        hs.close()
        if reduce_beta:   # k_T, Tc_T, Tb_T, a, k_b, beta(a=1,k_b), beta[n_z][nk + n_lnk + 1]
            h2d = sum(c["k_T"].nbytes * 3 + c["k_b"].nbytes * 2 + c["z_interp"].nbytes * (1 + nk + 51) for c in cosmos)
        else:
            h2d = sum(c["k_T"].nbytes * 3 + c["k_b"].nbytes + c["Tc_b"].nbytes * (1 if a.pageable else 2) +
                      c["z_interp"].nbytes for c in cosmos)
        h2d += B * (400 + 64 * 8 * 3)  # per-cosmology scalars and output redshift lists
        e2e = {"value": world * outputs_per_step * steps / t_pipe, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * t_pipe / steps,
               "serial_value": world * outputs_per_step / t_serial, "serial_ms_per_step": 1e3 * t_serial,
               "timing": "wall clock around the C-ABI calls, max over ranks",
               "inputs": ("pageable" if a.pageable else "page-locked (torch pin_memory)") + " numpy buffers holding the "
                         "full CAMB tables; " + ("the host pre-reduces the Beta_P table to what the run consumes "
                                                 "(reduce_beta=1)" if reduce_beta else "full Beta_P tables uploaded"),
               "path": "rtrg_pipeline_submit / _wait / _release, depth 2: every step stages its inputs (host threads), "
                       "sends them host->device, initialises and evolves them and copies its tables device->host "
                       "into page-locked memory, the host reads them; batch i+1's staging/H2D/initialisation overlap "
                       "batch i's evolution.  serial_value: the same calls on one handle, nothing overlapped"}

    grid = rt.grid_info(nk)
    n_bil, ms_bil = prof["k_bilinear"]
    flop_set = flops_per_matvec_set(grid, nk)
    achieved = flop_set * sets * prof_steps / (ms_bil * 1e-3) * 1e-12 if ms_bil > 0 else 0.0
    return dict(value=value, ms_total=ms_total, ms_per_step=ms_total / steps, launches=int(launches), clk=clk,
                n_failed=n_failed, e2e=e2e, prof=prof, ms_prof_per_step=ms_prof / prof_steps, sets=sets, evals=evals,
                achieved=achieved, n_bil=n_bil, ms_bil=ms_bil, flop_set=flop_set, tables=tables_keep, n_out=n_out,
                l2_inputs_gb=sum(c["Tc_b"].nbytes * 2 for c in cosmos) / 1e9)


def mode_cpu_samples(a, specs):
    """Reference samples of the sub-record workloads, run CONCURRENTLY on disjoint shares of the host
    cores (one single-thread process per core of the share)."""
    cores = core_list()
    share = max(1, len(cores) // len(specs))
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        jobs = []
        for j, (name, nk, mode, total) in enumerate(specs):
            binary = reference_binary(nk)
            if binary is None:
                continue
            mine = cores[j * share:(j + 1) * share]
            dirs = reference_dirs(len(mine), total, mode, a.subsample, os.path.join(tmp, name), True)
            jobs.append((name, mine, time.perf_counter(), reference_spawn(dirs, binary, mine)))
        for name, mine, t0, procs in jobs:
            rcs = [p.wait() for p in procs]
            wall = time.perf_counter() - t0
            if not any(rcs):
                out[name] = {"value": len(mine) * 8 / wall, "unit": UNIT, "cores": len(mine), "kind": "reference",
                             "sample": "%d cosmologies x 8 redshifts, one single-thread process per core on %d of the "
                                       "%d host cores (the other sub-record samples ran beside it), %.1f s wall"
                                       % (len(mine), len(mine), len(cores), wall)}
    return out


def files_record(a, local_rank, n_models=256, subsample=27, passes=4):
    """File-to-file throughput of the batch front-end redTimeBatch_b200 (what replaces the model loop
    of scripts/runRedTimeBatch:91-99): run directories with params_redTime.dat + 13 CAMB files in,
    redTime_<MODEL>.dat files out -- parsing, GPU pass and formatted writing overlapped chunk by
    chunk.  Tables with the scripts' default density (every 27th row of example 1: 572 rows)."""
    from redtime_b200 import workload as wl
    exe = os.path.join(ROOT, "redtime_b200", "redTimeBatch_b200")
    if not os.path.exists(exe):
        return None
    base = wl.load_example1(subsample)
    cosmos = wl.make_cosmologies(n_models, base, seed=wl.SEED, total=a.cosmologies)
    with tempfile.TemporaryDirectory() as tmp:
        names = []
        for i, c in enumerate(cosmos):
            wl.write_run_dir(os.path.join(tmp, "M%04d" % i), c)
            names.append("M%04d" % i)
        # the manifest lists every directory `passes` times, one pass per GPU chunk: the work of
        # passes x n_models models without writing that many input directories in the benchmark
        open(os.path.join(tmp, "manifest.txt"), "w").write("\n".join(names * passes) + "\n")
        env = dict(os.environ, RTRG_DEVICE=str(local_rank), RTRG_BATCH_CHUNK=str(n_models))
        best = None
        for _ in range(2):   # first run: cold page cache and CUDA context; report the second
            t0 = time.perf_counter()
            p = subprocess.run([exe, os.path.join(tmp, "manifest.txt")], env=env, stdout=subprocess.PIPE,
                               stderr=subprocess.PIPE, text=True)
            wall = time.perf_counter() - t0
            if p.returncode != 0:
                return {"error": (p.stderr or p.stdout)[-300:]}
            best = (wall, p.stdout.strip().split("\n")[-1])
        ok = all(os.path.getsize(os.path.join(tmp, nm, "redTime_%s.dat" % nm)) > 8 * 128 * 17 * 20 for nm in names)
    wall, summary = best
    inner = float(summary.split(" s wall")[0].split()[-1])
    n_total = n_models * passes
    return {"value": n_total * 8 / inner, "unit": UNIT, "models": n_total, "camb_rows": int(base["k_T"].size),
            "seconds_inside_process": inner, "seconds_process_wall": wall, "all_tables_written": bool(ok),
            "stage_times": summary,
            "what": "redTimeBatch_b200 <manifest>: %d models (%d run directories with params_redTime.dat + 13 CAMB files "
                    "of %d rows, listed %d times) -> redTime_<MODEL>.dat files, chunks of %d models; value = outputs / "
                    "seconds between manifest read and last file closed, INCLUDING the one-off start-up (CUDA context, "
                    "two handles, weight tables from the cache) that stage_times lists"
                    % (n_total, n_models, int(base["k_T"].size), passes, n_models)}


def run_b200(a, rank, world, local_rank):
    import torch
    import redtime_b200 as rt

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; redtime_b200 has no CPU path (use --impl reference)")
    torch.cuda.set_device(local_rank)
    comm = Comm(world, local_rank)
    steps = a.steps if a.steps is not None else 5
    warm = max(a.warmup if a.warmup is not None else 3, 0)
    B = a.cosmologies
    clocks = ClockSampler(local_rank) if rank == 0 else None
    n_cmp = host_cores() if (rank == 0 and world == 1 and not a.no_cpu_baseline) else 0
    m = measure_batch(a, comm, rank, world, local_rank, B, a.nk, a.mode, steps, warm, not a.no_e2e,
                      prof_steps=min(steps, 3), clocks=clocks, keep_tables=n_cmp)
    n_failed = int(comm.reduce(m["n_failed"], "sum"))

    # ---- the other BASELINE configurations as short sub-records (same code path, fewer steps)
    modes, kshard = {}, None
    if not a.no_modes and a.mode == "1loop" and a.nk == 128 and not a.print_all:
        for name, nk, mode, Bm in (("full_trg", 128, "full", 256), ("nk256", 256, "1loop", 256)):
            r = measure_batch(a, comm, rank, world, local_rank, Bm, nk, mode, 3, 1, not a.no_e2e, prof_steps=1)
            peak_m = rt.dmma_peak_tflops(local_rank, 0.2) if rank == 0 else None
            modes[name] = {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"],
                           "metric": metric_name(nk) + (", full Time-RG" if mode == "full" else ""),
                           "workload": workload_name(Bm, nk, mode, a.subsample), "cosmologies_per_gpu": Bm, "steps": 3,
                           "warmup": 1, "e2e": None if r["e2e"] is None else
                           {k: r["e2e"][k] for k in ("value", "unit", "ms_per_step", "h2d_bytes_per_step",
                                                     "d2h_bytes_per_step", "serial_value")},
                           "n_failed": int(comm.reduce(r["n_failed"], "sum")),
                           "roofline": {"kernel": "k_bilinear", "achieved": r["achieved"], "unit": "TFLOP/s",
                                        "frac": (r["achieved"] / peak_m) if peak_m else None,
                                        "share_of_step": r["ms_bil"] / max(r["ms_prof_per_step"], 1e-9),
                                        "integral_evaluations_per_step": r["evals"]}}
        kshard = kshard_record(a, comm, rank, world, local_rank, steps=5, warm=2)

    if rank != 0:
        return
    # ---- roofline of the dominant kernel
    peak = rt.dmma_peak_tflops(local_rank, 0.5)
    peak_dfma = rt.dfma_peak_tflops(local_rank, 0.3)
    sm_clk = 1.965e9
    traffic, traffic_src = ncu_traffic()
    roof = {"bound": "fp64", "kernel": "k_bilinear", "achieved": m["achieved"], "peak": peak, "unit": "TFLOP/s",
            "frac": m["achieved"] / peak if peak else None,
            "traffic": traffic,
            "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, parsed from %s (ncu --set full of "
                            "tools/bench_integrals.py 64: 64 cosmologies x 42 matvec sets = 74.3 GFLOP algorithmic; the "
                            "23.7 MB of weight tables that launch reads stream from L2, DRAM sees them once)" % os.path.relpath(traffic_src, ROOT),
            "peak_source": "register-resident DMMA.8x8x4 loop measured live by rtrg_bench_dmma (MEASURED_PEAKS.json has no "
                           "FP64 figure); a scalar DFMA loop (rtrg_bench_dfma) reaches %.2f; nominal 148 SM x 64 FMA/clk "
                           "x 2 x 1.965 GHz = %.1f TFLOP/s.  k_bilinear issues its FMAs as DMMA.8x8x4: the same FP64 "
                           "units, ncu books them under sm__pipe_tensor_subpipe_dmma_cycles_active"
                           % (peak_dfma, 148 * 64 * 2 * sm_clk * 1e-12),
            "launches": m["n_bil"], "avg_launch_ms": m["ms_bil"] / max(m["n_bil"], 1),
            "timed": "CUDA events around every launch in %d extra steps (per-kernel events switch the graph replay "
                     "off; those steps took %.2f ms each against %.2f ms for the graph-launched timed steps)"
                     % (min(steps, 3), m["ms_prof_per_step"], m["ms_per_step"]),
            "algorithmic_flop_per_matvec_set": m["flop_set"], "matvec_sets_per_step": m["sets"],
            "integral_evaluations_per_step": m["evals"],
            "share_of_step": m["ms_bil"] / min(steps, 3) / m["ms_prof_per_step"]}
    kernels = {k: {"launches": v[0], "ms": round(v[1], 3)} for k, v in m["prof"].items() if v[0]}
    cpu = parity = None
    if not a.no_cpu_baseline and world == 1:
        cpu, parity = cpu_baseline_headline(a, m["tables"])
        if modes:
            samples = mode_cpu_samples(a, [("full_trg", 128, "full", 256), ("nk256", 256, "1loop", 256)])
            for k, v in samples.items():
                modes[k]["cpu_baseline"] = v
        if kshard is not None:
            kshard["cpu_baseline"] = kshard_cpu_sample(a)
    files = files_record(a, local_rank) if (world == 1 and modes) else None
    line = {"metric": metric_name(a.nk), "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(a, B, m["l2_inputs_gb"]),
            "clocks": m["clk"], "e2e": m["e2e"], "gpu_launches": m["launches"], "n_failed": n_failed,
            "roofline": roof, "cpu_baseline": cpu, "parity": parity, "modes": modes or None, "kshard": kshard, "files": files,
            "kernel_ms_in_profiled_steps": kernels}
    emit(json.dumps(line))


# ------------------------------------------------------------------------------------------
# k-sharded single cosmology (BASELINE configs[2])
# ------------------------------------------------------------------------------------------
def kshard_cpu_sample(a):
    binary = reference_binary(256, hiacc=True)
    if binary is None:
        return None
    from redtime_b200 import workload as wl
    with tempfile.TemporaryDirectory() as tmp:
        d = wl.extract_example1(os.path.join(tmp, "c"))
        p = os.path.join(d, "params_redTime.dat")
        src = open(p).read().split("\n")
        vals = [i for i, l in enumerate(src) if l.strip() and not l.startswith("#")]
        src[vals[10]] = "0"  # switches 1 0 1 1: full Time-RG
        open(p, "w").write("\n".join(src))
        wall = reference_step([d], binary, omp_threads=host_cores())
    return {"value": 7 / wall, "unit": UNIT, "cores": host_cores(), "kind": "reference",
            "sample": "the same cosmology: oracle/_ref/redTime_hiacc (nk=256, beta clamp [1e-5,20], n_lnk=1000, "
                      "a_early=1e-50), ONE process with %d OpenMP threads, %.1f s wall" % (host_cores(), wall)}


def kshard_record(a, comm, rank, world, local_rank, steps, warm):
    import torch
    import redtime_b200 as rt
    from redtime_b200 import workload as wl

    nk = 256
    if (nk // 8) % world:
        return None
    c = wl.load_example1(a.subsample)
    c["switches"] = [1, 0, 1, 1]
    # split the beta-side lags as well once the rows per rank are few: keeps every SM busy and
    # shortens the serial chain of a CTA (rtrg_config.v_split; round-off level change only)
    v_split = a.v_split if a.v_split else (1 if world == 1 else 2 if world == 2 else 4)
    h = rt.RedTimeB200(device=local_rank, nk=nk, beta_kmin=1e-5, beta_kmax=20.0, n_lnk=1000, a_early=1e-50,
                       k_shards=world, k_rank=rank, v_split=v_split)
    stream = torch.cuda.Stream()
    h.set_stream(stream.cuda_stream)
    transport = "none (one GPU)"
    if world > 1:
        transport = h.kshard_init_auto(comm.dist)
    h.add_cosmology(c)
    h.prepare()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step():
        with torch.cuda.stream(stream):
            flush.zero_()
        return h.run_resident()

    for _ in range(warm):
        step()
    # per-kernel event timing switches the CUDA-graph replay of the rounds off, so the kernel
    # shares come from one extra, untimed step
    h.device_init()          # also zeroes the work counters
    h.set_profiling(True)
    step()
    prof = h.profile()
    h.set_profiling(False)
    sets_per_step = h.matvec_sets(0)
    l0 = h.launch_count()
    comm.sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    comm.sync()
    ms = comm.reduce(ev0.elapsed_time(ev1))
    launches = h.launch_count() - l0
    cnt = h.counters(0)
    # end to end: host buffers in, tables out (every rank ends with the full tables)
    comm.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        h.clear()
        h.add_cosmology(c)
        h.prepare()
        tables, hdr, hdr0, status = h.run()
    torch.cuda.synchronize()
    t_e2e = comm.reduce(time.perf_counter() - t0)
    h.close()
    if rank != 0:
        return None
    n_out = tables[0].shape[0]
    # parity against the committed oracle output of the same configuration
    with gzip.open(os.path.join(ROOT, "tests", "golden", "example1_oracle_hiacc_full.dat.gz"), "rt") as f:
        ref = np.array([l.split() for l in f.read().split("\n") if l.strip() and not l.startswith("#")], dtype=float)
    ref = ref.reshape(tables[0].shape)
    err = np.max(np.abs(tables[0] - ref) / (np.abs(ref) + 1e-300), axis=(0, 1)) if a.subsample == 1 else None
    grid = rt.grid_info(nk)
    n_bil, ms_bil = prof["k_bilinear"]
    peak = rt.dmma_peak_tflops(local_rank, 0.3)
    # this rank's share: matvec sets x its nk/world rows
    achieved = flops_per_matvec_set(grid, nk) / world * sets_per_step / (ms_bil * 1e-3) * 1e-12 if ms_bil else 0.0
    return {"metric": "cosmology*redshift outputs/sec, ONE nk=256 high-accuracy full-TRG cosmology, k-sharded",
            "value": n_out * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_cosmology": ms / steps, "scaling": "strong",
            "workload": "BASELINE configs[2]: example-1 inputs, nk=256, beta clamp [1e-5,20], n_lnk=1000, a_early=1e-50, "
                        "switches 1 0 1 1, k rows sharded over %d GPU(s), exchange of ln P_ab per RHS stage + max-reduce "
                        "of the error norm per attempt, v_split=%d" % (world, v_split),
            "transport": transport,
            "e2e": {"value": n_out * steps / t_e2e, "unit": UNIT, "ms_per_step": 1e3 * t_e2e / steps,
                    "h2d_bytes_per_step": int(c["k_T"].nbytes * 3 + c["k_b"].nbytes + c["Tc_b"].nbytes * 2),
                    "d2h_bytes_per_step": int(tables[0].nbytes)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "kernel": "k_bilinear", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "launches": n_bil,
                         "avg_launch_ms": ms_bil / max(n_bil, 1), "share_of_step": ms_bil / (ms / steps),
                         "note": "k_bilinear timed with CUDA events in one extra step (event timing disables the "
                                 "CUDA-graph replay the timed steps use)"},
            "counters": cnt,
            "parity_max_rel_err_cols_1_10_vs_oracle": None if err is None else float(err[:10].max()),
            "kernel_ms_in_profiled_step": {k: {"launches": v[0], "ms": round(v[1], 3)} for k, v in prof.items() if v[0]}}


def run_kshard(a, rank, world, local_rank):
    import torch
    torch.cuda.set_device(local_rank)
    comm = Comm(world, local_rank)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
    rec = kshard_record(a, comm, rank, world, local_rank, steps=a.steps if a.steps is not None else 5,
                        warm=max(a.warmup if a.warmup is not None else 3, 0))
    if rank != 0:
        return
    rec.update({"ms_per_step": rec["ms_per_cosmology"], "higher_is_better": True, "vs_baseline": None, "dtype": "f64",
                "data": "example-1 CAMB tables (the reference's examples/1_redTime)",
                "config": {"workload": rec["workload"], "l2": "256 MiB flush before every step"},
                "clocks": clocks.stop()})
    if not a.no_cpu_baseline and world == 1:
        rec["cpu_baseline"] = kshard_cpu_sample(a)
    emit(json.dumps(rec))


_REAL_STDOUT = None


def quiet_stdout():
    """NCCL and torchrun print banners on stdout; keep fd 1 for the ONE JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, (line + "\n").encode())
    else:
        print(line, flush=True)


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference_arm(a, rank)
        return
    if world == 1 and a.gpus > 1:
        # plain `python bench.py --gpus N`: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517"] + sys.argv
        sys.exit(subprocess.call(cmd))
    quiet_stdout()
    if a.workload == "kshard":
        run_kshard(a, rank, world, local_rank)
    else:
        run_b200(a, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
