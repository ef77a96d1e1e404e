"""The N>1 host-side path on CPU: world_size-2 `gloo` jobs (no GPU needed).
  * batch front-end: run directories are sharded cosmology-major, each rank handles only its
    share, the merged status map reaches every rank (the data path has no collective);
  * bench.py --impl reference under torchrun: rank 0 alone runs and prints ONE JSON line."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT, oracle_available

WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from redtime_b200 import batch
dist.init_process_group("gloo")
rank = dist.get_rank()
seen = []
def fake_runner(dirs, **cfg):           # stands in for the GPU pass of one rank
    seen.extend(dirs)
    return {d: (7 if d.endswith("M003") else 0) for d in dirs}
dirs = ["/runs/M%03d" % i for i in range(1, 8)]
res = batch.run_batch(dirs, runner=fake_runner, dist=dist)
# one file per rank: two ranks writing to the shared stdout pipe can interleave inside a line
with open(os.path.join(sys.argv[2], "result_%d.json" % rank), "w") as f:
    json.dump({"rank": rank, "seen": seen, "res": res}, f)
dist.destroy_process_group()
'''


def torchrun(nproc, args, port, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + args
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                          timeout=timeout)


def test_shard_is_a_partition():
    from redtime_b200 import batch
    items = list(range(11))
    parts = [batch.shard(items, 4, r) for r in range(4)]
    assert sorted(sum(parts, [])) == items
    assert parts[0] == [0, 4, 8] and parts[3] == [3, 7]
    assert batch.shard(items, 1, 0) == items


def test_manifest_reader(tmp_path):
    from redtime_b200 import batch
    m = tmp_path / "manifest.txt"
    m.write_text("# models\nM001\n/abs/M002  # comment\n\n")
    assert batch.read_manifest(str(m)) == [str(tmp_path / "M001"), "/abs/M002"]
    assert batch.model_name("/x/y/M007/") == "M007"


def test_batch_front_end_world_size_2_gloo(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    p = torchrun(2, [str(w), ROOT, str(tmp_path)], 29611)
    assert p.returncode == 0, p.stderr[-2000:]
    outs = [json.load(open(tmp_path / ("result_%d.json" % r))) for r in range(2)]
    by_rank = {o["rank"]: o for o in outs}
    assert by_rank[0]["seen"] == ["/runs/M001", "/runs/M003", "/runs/M005", "/runs/M007"]
    assert by_rank[1]["seen"] == ["/runs/M002", "/runs/M004", "/runs/M006"]
    for o in outs:  # every rank holds the merged status map
        assert len(o["res"]) == 7 and o["res"]["/runs/M003"] == 7 and o["res"]["/runs/M002"] == 0


@pytest.mark.skipif(not oracle_available(), reason="oracle/_ref not built (make -C oracle)")
def test_reference_arm_under_torchrun_prints_one_line():
    p = torchrun(2, ["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                     "--ref-procs", "2", "--subsample", "128"], 29612)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.split("\n") if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "outputs/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 2
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["n_gpus"] == 2
