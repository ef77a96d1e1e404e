"""The in-process seam, bound as the reference would bind it:
oracle/_ref/redTime_seam is the reference's UNMODIFIED main() and GSL driver loop
(src/redTime.cc:1551-1745, mini-GSL shim) with the GSL callback `derivatives` (:1416, registered at
:1596) served by rtrg_derivatives of libredtime_b200.so (oracle/seam_driver.cc).  Its stdout must
reproduce the oracle's: the stepper, the initial conditions and the printing are the reference's
own code, the right-hand side -- in full Time-RG mode the mode-coupling integrals of every stage
-- comes from the GPU."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ORACLE_REF, assert_table_parity, load_floor, parse_tables

pytestmark = pytest.mark.gpu
SEAM = os.path.join(ORACLE_REF, "redTime_seam")


@pytest.mark.skipif(not os.path.exists(SEAM), reason="oracle/_ref/redTime_seam not built (make -C oracle)")
@pytest.mark.parametrize("tag,fixture", [("full", "example1_full_dir"), ("1loop", "example1_dir")])
def test_reference_main_with_derivatives_from_the_c_abi(tag, fixture, request):
    d = request.getfixturevalue(fixture)
    p = subprocess.run([SEAM], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600,
                       env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert p.returncode == 0, p.stderr
    assert "derivatives() calls served by rtrg_derivatives" in p.stderr
    n_calls = int(p.stderr.split("seam driver:")[-1].split()[0])
    assert n_calls >= 100      # GSL's RKF45: 6 per attempt + dydt_in
    hdr, tab = parse_tables(p.stdout)
    with gzip.open(os.path.join(GOLDEN, "example1_oracle_%s.dat.gz" % tag), "rt") as f:
        rhdr, ref = parse_tables(f.read())
    assert hdr == rhdr         # same accepted steps: eta, a, z, H, sigma_v^2 to the 12 printed digits
    tab, ref = tab.reshape(7, 128, 17), ref.reshape(7, 128, 17)
    ex = assert_table_parity(tab, ref, load_floor(tag), what="seam " + tag)
    # columns 1-10 are far inside the tolerance: the step sequence is identical
    assert np.all(ex[:10] < 1e-2), ex
