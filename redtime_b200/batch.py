"""Batch front-end: what replaces the reference's sequential `for model in models` loop
(scripts/runRedTimeBatch:91-99 -> scripts/runRedTime:196-229, one `redTime > redTime_<MODEL>.dat`
process per model) with one GPU pass per rank.

A manifest is a list of run directories, each holding `params_redTime.dat` and the CAMB files it
names (exactly what scripts/runRedTime prepares).  Run directories are sharded across the ranks
of a `torch.distributed` job cosmology-major (a cosmology's redshifts share one ODE trajectory
and stay together); there is no data-path collective -- only the per-model status words are
gathered at the end so that rank 0 can report.  Every rank writes `redTime_<MODEL>.dat` next to
the inputs, byte-compatible with the reference's stdout (rtrg_print_result).

    torchrun --nproc-per-node 8 -m redtime_b200.batch manifest.txt
"""
import os
import sys


def shard(items, world, rank):
    """Round-robin, cosmology-major: item i goes to rank i % world."""
    return [x for i, x in enumerate(items) if i % world == rank]


def read_manifest(path):
    base = os.path.dirname(os.path.abspath(path))
    out = []
    with open(path) as f:
        for line in f:
            line = line.split("#")[0].strip()
            if line:
                out.append(line if os.path.isabs(line) else os.path.join(base, line))
    return out


def model_name(run_dir):
    return os.path.basename(os.path.normpath(run_dir))


def gpu_runner(run_dirs, device=0, **cfg):
    """Process a list of run directories on one GPU; returns {run_dir: status}."""
    from . import binding as rt
    if not run_dirs:
        return {}
    h = rt.RedTimeB200(device=device, **cfg)
    h.add_run_dirs(run_dirs)  # parallel parse (std::from_chars, one host thread per directory)
    h.prepare()
    tables, hdr, hdr0, status = h.run(raise_on_ode_failure=False)
    for i, d in enumerate(run_dirs):
        rt.print_result(os.path.join(d, "redTime_%s.dat" % model_name(d)), h.nk, tables[i], hdr[i], hdr0[i])
    h.close()
    return {d: int(status[i]) for i, d in enumerate(run_dirs)}


def run_batch(run_dirs, runner=gpu_runner, dist=None, **cfg):
    """Shard run_dirs over the ranks of `dist` (an initialised torch.distributed module, or None
    for a single process), run `runner` on this rank's share and return, on every rank, the
    merged {run_dir: status} of the whole job."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    mine = shard(list(run_dirs), world, rank)
    local = runner(mine, **cfg)
    if dist is None or world == 1:
        return dict(local)
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    merged = {}
    for part in gathered:
        merged.update(part)
    missing = [d for d in run_dirs if d not in merged]
    if missing:
        raise RuntimeError("batch: no result for %s" % missing)
    return merged


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        print(__doc__)
        return 2
    dirs = read_manifest(argv[0])
    dist = None
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("gloo")  # status words only; the data path has no collective
        dist = dist_mod
    res = run_batch(dirs, dist=dist, device=local_rank)
    if dist is None or dist.get_rank() == 0:
        bad = {d: s for d, s in res.items() if s}
        print("redtime_b200.batch: %d models, %d failed" % (len(res), len(bad)))
        for d, s in bad.items():
            print("  %s: status %d" % (d, s))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
