"""The reference's own -DHIGH_ACCURACY build (src/redTime.cc:90-94,141-145): nk = 512, RKF45
tolerances (1e-15, 1e-6).  Here that is plain run-time configuration.  Oracle output:
oracle/_ref/redTime_HIGH_ACCURACY (the unmodified sources compiled with that flag) on example 1."""
import gzip
import os

import numpy as np
import pytest

import redtime_b200 as rt
from conftest import GOLDEN, parse_tables, load_floor, assert_table_parity

pytestmark = pytest.mark.gpu


def test_high_accuracy_build_of_the_reference(example1_dir):
    with gzip.open(os.path.join(GOLDEN, "example1_oracle_HIGH_ACCURACY_1loop.dat.gz"), "rt") as f:
        hdr, ref = parse_tables(f.read())
    ref = ref.reshape(7, 512, 17)
    h = rt.RedTimeB200(nk=512, eps_abs=1e-15, eps_rel=1e-6)
    h.add_cosmology(rt.read_run_dir(example1_dir))
    h.prepare()
    tables, hd, hd0, status = h.run()
    cnt = h.counters(0)
    h.close()
    assert not status.any()
    tab = tables[0]
    assert tab.shape == ref.shape
    assert cnt["attempts"] > 100          # the tight tolerance takes hundreds of steps
    # every k and column; the floor is larger at np = 2048 and was measured with this oracle build
    assert_table_parity(tab, ref, load_floor("HIGH_ACCURACY_1loop"), what="HIGH_ACCURACY")
