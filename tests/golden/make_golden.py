#!/usr/bin/env python3
"""Regenerate the committed stage-level goldens from the oracle (oracle/_ref must be built:
`make -C oracle`).  The oracle is the UNMODIFIED reference compiled against the mini-GSL
shim; see oracle/Makefile.  Usage: python tests/golden/make_golden.py"""
import gzip
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import conftest  # noqa: E402

if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as tmp:
        d1 = conftest.make_example1_dir(os.path.join(tmp, "a"))
        conftest.run_oracle_stage(d1, os.path.join(HERE, "example1_stage_1loop.npz"))
        d2 = conftest.make_example1_dir(os.path.join(tmp, "b"), switches=[1, 0, 1, 1])
        conftest.run_oracle_stage(d2, os.path.join(HERE, "example1_stage_full.npz"))
        # end-to-end oracle outputs
        for tag, d in (("1loop", d1), ("full", d2)):
            txt = conftest.run_oracle_binary(d)
            with gzip.open(os.path.join(HERE, "example1_oracle_%s.dat.gz" % tag), "wt") as f:
                f.write(txt)
        # variants with the README-sanctioned constant edits (oracle/Makefile): nk=256, the
        # high-accuracy growth/beta settings of config 3, and all PRINT* column groups (config 4)
        conftest.run_oracle_stage(d1, os.path.join(HERE, "example1_stage_1loop_nk256.npz"),
                                  lib="libredtime_stage_nk256.so", light=True)
        for tag, binary, d in (("nk256_1loop", "redTime_nk256", d1), ("nk256_full", "redTime_nk256", d2),
                               ("hiacc_full", "redTime_hiacc", d2), ("printall_1loop", "redTime_printall", d1),
                               ("HIGH_ACCURACY_1loop", "redTime_HIGH_ACCURACY", d1)):
            txt = conftest.run_oracle_binary(d, binary=binary)
            with gzip.open(os.path.join(HERE, "example1_oracle_%s.dat.gz" % tag), "wt") as f:
                f.write(txt)
    print("goldens written to", HERE)
