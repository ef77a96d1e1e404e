#!/usr/bin/env python3
"""Regenerate the committed stage-level goldens from the oracle (oracle/_ref must be built:
`make -C oracle`).  The oracle is the UNMODIFIED reference compiled against the mini-GSL
shim; see oracle/Makefile.  Usage: python tests/golden/make_golden.py"""
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
            import gzip
            with gzip.open(os.path.join(HERE, "example1_oracle_%s.dat.gz" % tag), "wt") as f:
                f.write(txt)
    print("goldens written to", HERE)
