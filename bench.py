#!/usr/bin/env python3
"""bench.py -- cosmology*redshift outputs per second of the Time-RG hot path at nk=128.

  python bench.py [--gpus N] [--steps K] [--warmup W]              this library on N B200s
  python bench.py --impl reference [--gpus N] [--steps K] ...      the reference on host cores

A "step" is one pass of the whole hot path over one batch of synthetic cosmologies per GPU
(device-side linear theory tables + sigma_8 normalisation + 1-loop mode-coupling integrals +
Time-RG evolution with the batched RKF45 stepper + output tables at 8 redshifts).

  value : outputs/s with the inputs already resident in HBM (CUDA events on the library's
          stream, max over ranks); work of all ranks / that time ("weak" scaling: every rank
          owns its own --cosmologies batch, no data-path collective).
  e2e   : the same through the reference-facing C-ABI with HOST buffers: rtrg_add_cosmology
          (host) -> rtrg_prepare (H2D + device init) -> rtrg_run (evolution + D2H of the tables).
  roofline     : the dominant kernel (k_bilinear, FP64 FMA pipe): algorithmic FLOP of all its
                 launches / its CUDA-event time, against the DFMA peak measured live.
  cpu_baseline : oracle/_ref/redTime (the UNMODIFIED reference sources + mini-GSL shim) as one
                 single-thread process per host core on a bounded sample of the same workload.
"""
import argparse
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
                    help="batch: cosmologies sharded over GPUs, no collective (headline); kshard: ONE "
                         "high-accuracy nk=256 full Time-RG cosmology (BASELINE configs[2]) with its k rows "
                         "sharded over the GPUs, NCCL all-gather of ln P_ab per RHS stage (strong scaling)")
    return ap.parse_args()


def workload_name(a):
    sw = "1 1 1 1 (1-loop)" if a.mode == "1loop" else "1 0 1 1 (full Time-RG)"
    return ("throughput sweep (BASELINE configs[4]): %d w0wa+massive-nu cosmologies per GPU x 8 redshifts, "
            "nk=%d, switches %s, example-1 CAMB tables (%d rows x 12 redshifts) with per-cosmology tilt, "
            "Latin-hypercube parameters seed 20261018" % (a.cosmologies, a.nk, sw, -(-15447 // a.subsample)))


def metric_name(a):
    return METRIC if a.nk == 128 else METRIC.replace("nk=128", "nk=%d" % a.nk)


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference binary on the host cores
# ------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def reference_binary(nk=128):
    """The reference executable: nk is a compile-time constant there (redTime.cc:93)."""
    name = {128: "redTime", 256: "redTime_nk256"}.get(nk)
    if name is None:
        return None
    p = os.path.join(ROOT, "oracle", "_ref", name)
    return p if os.path.exists(p) else None


def reference_step(dirs, binary):
    """One process per run directory, one OpenMP thread each, pinned round-robin to the
    allowed cores; returns wall seconds from first spawn to last exit."""
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = "1"
    try:
        cores = sorted(os.sched_getaffinity(0))
    except AttributeError:
        cores = list(range(os.cpu_count() or 1))
    use_taskset = shutil.which("taskset") is not None
    t0 = time.perf_counter()
    procs = []
    for i, d in enumerate(dirs):
        cmd = [binary]
        if use_taskset:
            cmd = ["taskset", "-c", str(cores[i % len(cores)])] + cmd
        procs.append(subprocess.Popen(cmd, cwd=d, env=env, stdout=open(os.path.join(d, "out.dat"), "w"),
                                      stderr=subprocess.DEVNULL))
    rcs = [p.wait() for p in procs]
    wall = time.perf_counter() - t0
    if any(rcs):
        raise RuntimeError("reference process failed: %s" % rcs)
    return wall


def reference_setup(a, nproc, tmp):
    from redtime_b200 import workload as wl
    base = wl.load_example1(a.subsample)
    sw = (1, 1, 1, 1) if a.mode == "1loop" else (1, 0, 1, 1)
    cosmos = wl.make_cosmologies(nproc, base, seed=wl.SEED, switches=sw)
    return [wl.write_run_dir(os.path.join(tmp, "c%04d" % i), c) for i, c in enumerate(cosmos)]


def run_reference_arm(a, rank):
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
        dirs = reference_setup(a, nproc, tmp)
        for _ in range(warm):
            reference_step(dirs, binary)
        t = [reference_step(dirs, binary) for _ in range(steps)]
    n_out = len(REDSHIFTS())
    wall = float(np.sum(t))
    value = steps * nproc * n_out / wall
    sample = "%d cosmologies x %d redshifts per step, one single-thread process per core" % (nproc, n_out)
    line = {"impl": "reference", "metric": metric_name(a), "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * wall / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(json.dumps(line))


def REDSHIFTS():
    from redtime_b200 import workload as wl
    return wl.REDSHIFTS_CE


def cpu_baseline(a):
    """Bounded sample for the default bench line: ONE step of the reference arm."""
    binary = reference_binary(a.nk)
    if binary is None:
        return None
    nproc = host_cores()
    with tempfile.TemporaryDirectory() as tmp:
        dirs = reference_setup(a, nproc, tmp)
        wall = reference_step(dirs, binary)
    n_out = len(REDSHIFTS())
    return {"value": nproc * n_out / wall, "unit": UNIT, "cores": nproc, "kind": "reference",
            "sample": "%d cosmologies x %d redshifts (one single-thread oracle/_ref/redTime process per core), "
                      "%.1f s wall" % (nproc, n_out, wall)}


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


def run_b200(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import redtime_b200 as rt
    from redtime_b200 import workload as wl

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; redtime_b200 has no CPU path (use --impl reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    steps = a.steps if a.steps is not None else 5
    warm = max(a.warmup if a.warmup is not None else 3, 0)

    B = a.cosmologies
    base = wl.load_example1(a.subsample)
    sw = (1, 1, 1, 1) if a.mode == "1loop" else (1, 0, 1, 1)
    cosmos = wl.make_cosmologies(B, base, seed=wl.SEED + 7919 * rank, switches=sw, pinned=not a.pageable)
    packed = rt.pack_cosmologies(cosmos)  # ctypes views of the same buffers, built once
    n_out = len(wl.REDSHIFTS_CE)
    outputs_per_step = B * n_out

    reduce_beta = 0 if a.full_beta else 1
    extra = dict(print_A=1, print_I=1, print_Q=1, print_bias=1) if a.print_all else {}
    h = rt.RedTimeB200(device=local_rank, nk=a.nk, reduce_beta=reduce_beta, **extra)
    stream = torch.cuda.Stream()
    h.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def upload():
        h.clear()
        h.add_cosmologies(packed)
        h.prepare()

    def e2e_step():
        upload()
        return h.run_pinned()

    def resident_step():
        with torch.cuda.stream(stream):
            flush.zero_()
        h.device_init()
        return h.run_resident()

    # ---- resident-input timing ("value")
    upload()
    for _ in range(warm):
        st = resident_step()
    if warm and st.any():
        print("bench.py: warning: %d cosmologies failed (status != 0)" % int(np.count_nonzero(st)), file=sys.stderr)
    clocks = ClockSampler(local_rank)
    h.set_profiling(True)
    l0 = h.launch_count()
    sync_all()
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        resident_step()
    ev1.record(stream)
    sync_all()
    clk = clocks.stop() if rank == 0 else None
    ms_total = reduce_max(ev0.elapsed_time(ev1))
    launches = h.launch_count() - l0
    prof = h.profile()
    h.set_profiling(False)
    evals = sum(h.counters(i)["integral_evals"] for i in range(B))  # per step (last run)
    sets = sum(h.matvec_sets(i) for i in range(B))                  # per step (device_init + run)
    value = world * outputs_per_step * steps / (ms_total * 1e-3)

    # ---- end to end through the C-ABI with host buffers
    e2e = None
    if not a.no_e2e:
        for _ in range(min(warm, 1)):
            e2e_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(steps):
            tables, hdr, hdr0, status = e2e_step()
        torch.cuda.synchronize()
        t_serial = reduce_max(time.perf_counter() - t0)
        if reduce_beta:   # k_T, Tc_T, Tb_T, a, k_b, beta(a=1,k_b), beta[n_z][nk + n_lnk + 1]
            h2d = sum(c["k_T"].nbytes * 3 + c["k_b"].nbytes * 2 + c["z_interp"].nbytes * (1 + a.nk + 51) for c in cosmos)
        else:
            h2d = sum(c["k_T"].nbytes * 3 + c["k_b"].nbytes + c["Tc_b"].nbytes * (1 if a.pageable else 2) +
                      c["z_interp"].nbytes for c in cosmos)
        h2d += B * (400 + 64 * 8 * 3)  # per-cosmology scalars and output redshift lists
        d2h = sum(t.nbytes for t in tables) + hdr.nbytes + hdr0.nbytes
        tables = [t.copy() for t in tables[B // 2:B // 2 + 1]] * B  # keep one table to compare against

        # double-buffered: a second handle stages and uploads batch i+1 (host threads + copy
        # engine) while batch i is computed and read back -- every step still moves its own
        # inputs host->device and its own tables device->host inside the timed region.  Measured
        # on one GPU per node only: with 8 ranks x 2 driver threads on the 32 host cores of the
        # box the double-buffered variant was slower than the serial one (48.6 k vs 148.7 k
        # outputs/s), so multi-rank runs report the serial number.
        t_pipe = None
        if world == 1:
            h2 = rt.RedTimeB200(device=local_rank, nk=a.nk, reduce_beta=reduce_beta, **extra)
            hs = [h, h2]
            for hh in hs:                      # warm both handles' arenas
                hh.clear()
                hh.add_cosmologies(packed)
                hh.prepare()
                hh.run_pinned()
            sync_all()

            def stage(hh):  # upload + device-side initialisation (growth ODE, QAG, 1-loop cache)
                hh.clear()
                hh.add_cosmologies(packed)
                hh.prepare()

            t0 = time.perf_counter()
            pending = threading.Thread(target=stage, args=(hs[0],))
            pending.start()
            results = []
            for i in range(steps):
                pending.join()
                cur = hs[i % 2]
                if i + 1 < steps:
                    pending = threading.Thread(target=stage, args=(hs[(i + 1) % 2],))
                    pending.start()
                tb_, _, _, st_ = cur.run_pinned()
                results.append(tb_[B // 2][-1, :, 7].copy())
            torch.cuda.synchronize()
            t_pipe = time.perf_counter() - t0
            h2.close()
            assert all(np.array_equal(x, tables[B // 2][-1, :, 7]) for x in results), "pipelined results differ"
        t_best = t_pipe if (t_pipe is not None and t_pipe < t_serial) else t_serial
        e2e = {"value": world * outputs_per_step * steps / t_best, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * t_best / steps,
               "double_buffered_value": None if t_pipe is None else world * outputs_per_step * steps / t_pipe,
               "serial_value": world * outputs_per_step * steps / t_serial, "serial_ms_per_step": 1e3 * t_serial / steps,
               "timing": "wall clock around the C-ABI calls, max over ranks",
               "inputs": ("pageable" if a.pageable else "page-locked (torch pin_memory)") + " numpy buffers holding the "
                         "full CAMB tables; " + ("the host pre-reduces the Beta_P table to what the run consumes "
                                                 "(reduce_beta=1)" if reduce_beta else "full Beta_P tables uploaded"),
               "path": "per step: rtrg_add_cosmologies (page-locked caller buffers: asynchronous H2D straight from "
                       "them; pageable ones: copied to a page-locked arena on host threads, chunked H2D overlapped) "
                       "-> rtrg_prepare -> rtrg_run -> "
                       "rtrg_fetch_outputs (D2H into page-locked memory).  value = the better of "
                       "double_buffered_value (one GPU only: batch i+1 is uploaded and initialised on a second "
                       "handle/stream while batch i evolves) and serial_value (one handle, nothing overlapped)"}

    if rank != 0:
        return
    # ---- roofline of the dominant kernel
    grid = rt.grid_info(a.nk)
    n_bil, ms_bil = prof["k_bilinear"]
    flop_set = flops_per_matvec_set(grid, a.nk)
    flop_total = flop_set * sets * steps
    peak = rt.dfma_peak_tflops(local_rank, 0.5)
    achieved = flop_total / (ms_bil * 1e-3) * 1e-12 if ms_bil > 0 else 0.0
    sm_clk = 1.965e9
    roof = {"bound": "fp64", "kernel": "k_bilinear", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak if peak else None,
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE k_bilinear launch from the committed
            # ncu --set full capture (profiles/r01_k_bilinear_final_ncu.txt: 64 cosmologies x 42 sets =
            # 74.3 GFLOP algorithmic): the 23.7 MB weight tables stream from L2, DRAM sees them once
            "traffic": 22763264 + 177664,
            "traffic_note": "bytes per launch of the ncu-captured launch (64 cosmologies x 42 matvec sets), not of "
                            "the average bench launch; algorithmic FLOP per DRAM byte there = 3.2e3",
            "peak_source": "DFMA loop measured live by rtrg_bench_dfma (MEASURED_PEAKS.json has no FP64 figure); "
                           "nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = %.1f TFLOP/s" % (148 * 64 * 2 * sm_clk * 1e-12),
            "launches": n_bil, "avg_launch_ms": ms_bil / max(n_bil, 1),
            "algorithmic_flop_per_matvec_set": flop_set, "matvec_sets_per_step": sets,
            "integral_evaluations_per_step": evals,
            "share_of_step": ms_bil / ms_total}
    kernels = {k: {"launches": v[0], "ms": round(v[1], 3)} for k, v in prof.items() if v[0]}
    cpu = None if a.no_cpu_baseline else cpu_baseline(a)
    line = {"metric": metric_name(a), "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "cosmologies_per_gpu": B, "redshifts": n_out, "nk": a.nk,
                       "mode": a.mode, "columns": 84 if a.print_all else 17, "l2": "256 MiB flush before every step; per-step inputs %.2f GB > L2"
                       % (sum(c["Tc_b"].nbytes * 2 for c in cosmos) / 1e9)},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "kernel_ms_in_timed_region": kernels}
    emit(json.dumps(line))


# ------------------------------------------------------------------------------------------
# k-sharded single cosmology (BASELINE configs[2])
# ------------------------------------------------------------------------------------------
def run_kshard(a, rank, world, local_rank):
    import gzip
    import torch
    import torch.distributed as dist
    import redtime_b200 as rt
    from redtime_b200 import workload as wl

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    steps = a.steps if a.steps is not None else 5
    warm = max(a.warmup if a.warmup is not None else 3, 0)
    nk = 256
    c = wl.load_example1(a.subsample)
    c["switches"] = [1, 0, 1, 1]
    # split the beta-side lags as well once the rows per rank are few: keeps every SM busy and
    # shortens the serial chain of a CTA (rtrg_config.v_split; round-off level change only)
    v_split = a.v_split if a.v_split else (1 if world == 1 else 2 if world == 2 else 4)
    h = rt.RedTimeB200(device=local_rank, nk=nk, beta_kmin=1e-5, beta_kmax=20.0, n_lnk=1000, a_early=1e-50,
                       k_shards=world, k_rank=rank, v_split=v_split)
    stream = torch.cuda.Stream()
    h.set_stream(stream.cuda_stream)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(rt.kshard_nccl_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        h.kshard_init_nccl(bytes(idt.cpu().numpy().tobytes()))
    h.add_cosmology(c)
    h.prepare()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

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
    clocks = ClockSampler(local_rank)
    l0 = h.launch_count()
    sync_all()
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    sync_all()
    clk = clocks.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = h.launch_count() - l0
    cnt = h.counters(0)
    # end to end: host buffers in, tables out (every rank ends with the full tables)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(steps):
        h.clear()
        h.add_cosmology(c)
        h.prepare()
        tables, hdr, hdr0, status = h.run()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    if rank != 0:
        return
    n_out = tables[0].shape[0]
    # parity against the committed oracle output of the same configuration
    with gzip.open(os.path.join(ROOT, "tests", "golden", "example1_oracle_hiacc_full.dat.gz"), "rt") as f:
        ref = np.array([l.split() for l in f.read().split("\n") if l.strip() and not l.startswith("#")], dtype=float)
    ref = ref.reshape(tables[0].shape)
    err = np.max(np.abs(tables[0] - ref) / (np.abs(ref) + 1e-300), axis=(0, 1)) if a.subsample == 1 else None
    grid = rt.grid_info(nk)
    n_bil, ms_bil = prof["k_bilinear"]
    peak = rt.dfma_peak_tflops(local_rank, 0.5)
    # this rank's share: matvec sets x its nk/world rows
    achieved = flops_per_matvec_set(grid, nk) / world * sets_per_step / (ms_bil * 1e-3) * 1e-12 if ms_bil else 0.0
    line = {"metric": "cosmology*redshift outputs/sec, ONE nk=256 high-accuracy full-TRG cosmology, k-sharded",
            "value": n_out * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "example-1 CAMB tables (the reference's examples/1_redTime)",
            "config": {"workload": "BASELINE configs[2]: single cosmology, nk=256, beta clamp [1e-5,20], n_lnk=1000, "
                                   "a_early=1e-50, switches 1 0 1 1, k rows sharded over %d GPU(s), NCCL all-gather of "
                                   "ln P_ab per RHS stage + max-reduce of the error norm per attempt, v_split=%d"
                                   % (world, v_split),
                       "l2": "256 MiB flush before every step"},
            "clocks": clk,
            "e2e": {"value": n_out * steps / t_e2e, "unit": UNIT, "ms_per_step": 1e3 * t_e2e / steps,
                    "h2d_bytes_per_step": int(c["k_T"].nbytes * 3 + c["k_b"].nbytes + c["Tc_b"].nbytes * 2),
                    "d2h_bytes_per_step": int(tables[0].nbytes)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "kernel": "k_bilinear", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": None, "launches": n_bil,
                         "avg_launch_ms": ms_bil / max(n_bil, 1), "share_of_step": ms_bil / (ms / steps),
                         "note": "k_bilinear timed with CUDA events in one extra step (event timing disables the "
                                 "CUDA-graph replay the timed steps use)"},
            "counters": cnt,
            "parity_max_rel_err_cols_1_10_vs_oracle": None if err is None else float(err[:10].max()),
            "kernel_ms_in_timed_region": {k: {"launches": v[0], "ms": round(v[1], 3)} for k, v in prof.items() if v[0]}}
    emit(json.dumps(line))


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
