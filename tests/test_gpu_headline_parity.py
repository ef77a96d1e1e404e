"""Parity of the path bench.py's headline number runs, against the oracle binary DIRECTLY:
reduce_beta = 1 (the host pre-reduces the Beta_P table), page-locked caller tables sent by the
copy engine without a host copy, the full 15 447-row tilted CAMB tables, Latin-hypercube
cosmologies -- the first members of bench.py's own draw (workload.make_cosmologies(total=1024)).
The same batch also goes through the double-buffered pipeline (rtrg_pipeline_*), which must not
change a bit.

Tolerances: north star (1e-6 columns 1-7, 1e-5 columns 8-17) at EVERY k; for the mode-coupling
columns 11-17 the bound is 1e-5 of the local scale + 5 x the oracle's own round-off floor, which
is measured here for exactly these cosmologies (two more oracle runs each with sigma_8 / n_s
moved by one ulp; SURVEY H2, redTime.cc:1182-1184)."""
import os

import numpy as np
import pytest

import redtime_b200 as rt
from redtime_b200 import workload as wl
from conftest import ORACLE_REF, assert_table_parity, oracle_tables_with_floor

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(os.path.join(ORACLE_REF, "redTime")), reason="oracle/_ref not built")
def test_bench_path_reduced_beta_page_locked_full_tables(tmp_path):
    base = wl.load_example1(1)
    assert base["k_T"].size == 15447
    n = 3
    cosmos = wl.make_cosmologies(n, base, seed=wl.SEED, total=1024, pinned=True)
    # one member in pageable memory: staged through the library's own page-locked arena
    cosmos[1] = wl.make_cosmologies(n, base, seed=wl.SEED, total=1024, pinned=False)[1]
    dirs = [wl.write_run_dir(str(tmp_path / ("c%d" % i)), c) for i, c in enumerate(cosmos)]
    packed = rt.pack_cosmologies(cosmos)

    h = rt.RedTimeB200(reduce_beta=1)
    h.add_cosmologies(packed)
    h.prepare()
    tables, hdr, hdr0, status = h.run_pinned()
    tables, hdr, hdr0 = [t.copy() for t in tables], hdr.copy(), hdr0.copy()   # views of the handle's memory
    assert not status.any()
    full = rt.RedTimeB200(reduce_beta=0)   # the full Beta_P tables through the direct path
    full.add_cosmologies(packed)
    full.prepare()
    tables_full, *_ = full.run()
    full.close()

    pipe = rt.Pipeline(depth=2, reduce_beta=1)
    tickets = [pipe.submit(packed) for _ in range(3)]   # three batches, two in flight
    for t in tickets:
        tp, hp, h0p, sp = pipe.wait(t)
        assert not sp.any()
        for a, b in zip(tp, tables):
            assert np.array_equal(a, b)
        assert np.array_equal(hp[:, :8], hdr[:, :8]) and np.array_equal(h0p, hdr0)
        pipe.release(t)
    pipe.close()
    h.close()

    orc = oracle_tables_with_floor(dirs, "redTime")
    for i, (tab, (rhdr, ref, floor)) in enumerate(zip(tables, orc)):
        assert tab.shape == (8, 128, 17)
        ex = assert_table_parity(tab, ref.reshape(tab.shape), floor.reshape(tab.shape), what="cosmology %d" % i)
        assert_table_parity(tables_full[i], ref.reshape(tab.shape), floor.reshape(tab.shape), what="full beta %d" % i)
        p = str(tmp_path / ("mine%d.dat" % i))
        rt.print_result(p, 128, tab, hdr[i], hdr0[i])
        mine = [l for l in open(p).read().split("\n") if l.startswith("#")]
        assert mine[1:] == rhdr[1:]     # '###main' lines to the 12 printed digits
        print("cosmology %d: excess per column (1 = at tolerance)" % i, np.round(ex, 4))
