#!/usr/bin/env python3
"""Worker of tests/test_gpu_kshard_multi.py: one rank of a k-sharded run, one process per GPU
(torchrun).  Config 3: nk = 256, high-accuracy growth / beta settings, full Time-RG.  Every rank
writes its tables to <out>/rank<r>.npz; rank 0 also runs the cosmology unsharded.
usage: kshard_worker.py <repo root> <run dir> <out dir>"""
import os
import sys

import numpy as np

sys.path.insert(0, sys.argv[1])
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import redtime_b200 as rt  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = dict(nk=256, beta_kmin=1e-5, beta_kmax=20.0, n_lnk=1000, a_early=1e-50)
inp = rt.read_run_dir(sys.argv[2])
h = rt.RedTimeB200(device=local, k_shards=world, k_rank=rank, **cfg)
transport = h.kshard_init_auto(dist)
h.add_cosmology(inp)
h.prepare()
tables, hdr, hdr0, status = h.run()
again, *_ = h.run()               # a second run on the same handle (sequence numbers keep counting)
cnt = h.counters(0)
h.close()
out = dict(table=tables[0], again=again[0], hdr=hdr[0], status=status, attempts=cnt["attempts"], rejected=cnt["rejected"],
           transport=np.array(transport))
if rank == 0:
    s = rt.RedTimeB200(device=local, **cfg)
    s.add_cosmology(inp)
    s.prepare()
    t1, h1, _, st1 = s.run()
    c1 = s.counters(0)
    s.close()
    out.update(single=t1[0], single_hdr=h1[0], single_attempts=c1["attempts"], single_rejected=c1["rejected"])
np.savez(os.path.join(sys.argv[3], "rank%d.npz" % rank), **out)
dist.barrier()
dist.destroy_process_group()
