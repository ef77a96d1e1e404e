"""End-to-end parity: examples/1_redTime through the batched device stepper vs
  (a) the reference's own golden example_redTime_result.dat (genuine GSL), and
  (b) the ORACLE output (unmodified reference + mini-GSL shim), 1-loop and full Time-RG.
Tolerance (north star): max relative error <= 1e-6 in columns 1-7 and <= 1e-5 in columns
8-17; columns 11-17 of the lowest k rows carry the reference's own FFT round-off (SURVEY H2,
V13), so there the bound is 1e-5 |ref| + floor(k, col) with the floor measured between the
genuine-GSL golden and the oracle."""
import gzip
import os

import numpy as np
import pytest

import redtime_b200 as rt
from conftest import GOLDEN, parse_tables, load_floor, assert_table_parity, local_scale, smooth_floor, FLOOR_C

pytestmark = pytest.mark.gpu
NK = 128


def load_oracle(tag):
    with gzip.open(os.path.join(GOLDEN, "example1_oracle_%s.dat.gz" % tag), "rt") as f:
        return parse_tables(f.read())


def run(dirname, **cfg):
    h = rt.RedTimeB200(**cfg)
    h.add_cosmology(rt.read_run_dir(dirname))
    h.prepare()
    tables, hdr, hdr0, status = h.run()
    cnt = h.counters(0)
    h.close()
    return tables[0], hdr[0], hdr0[0], status[0], cnt


def col_err(got, ref):
    return np.max(np.abs(got - ref) / (np.abs(ref) + 1e-300), axis=(0, 1))


def test_example1_1loop(example1_dir, golden_example1):
    tab, hdr, hdr0, status, cnt = run(example1_dir)
    assert status == 0
    _, gold = golden_example1
    gold = gold.reshape(7, NK, 17)
    _, orc = load_oracle("1loop")
    orc = orc.reshape(7, NK, 17)
    assert tab.shape == (7, NK, 17)
    # the step sequence genuine GSL took
    assert (cnt["attempts"], cnt["rejected"]) == (23, 4)
    # North-star tolerances at EVERY k: 1e-6 (columns 1-7), 1e-5 (8-10), and for the mode-coupling
    # columns 11-17 1e-5 of the local scale + 5 x the reference's own round-off floor, measured by
    # moving sigma_8 / n_s by one ulp (tests/golden/make_floor.py; SURVEY H2: columns 16-17 are
    # 1e3..1e9-fold cancellations of the J integrals below k = 3e-3 h/Mpc, rt:1182-1184).  The
    # genuine-GSL golden carries the noise of a different FFT as well: 8 x the floor there.
    floor = load_floor("1loop")
    assert_table_parity(tab, orc, floor, what="vs oracle")
    assert_table_parity(tab, gold, floor, c=8.0, what="vs genuine-GSL golden")


def test_example1_header_lines(example1_dir, golden_example1, tmp_path):
    tab, hdr, hdr0, status, cnt = run(example1_dir)
    p = str(tmp_path / "out.dat")
    rt.print_result(p, NK, tab, hdr, hdr0)
    mine = [l for l in open(p).read().split("\n") if l.startswith("#")]
    ref, _ = golden_example1
    assert mine == ref  # banner, eta_fin/sigmaV2 line and the 7 per-output lines, byte-identical


def test_example1_full_trg(example1_full_dir):
    tab, hdr, hdr0, status, cnt = run(example1_full_dir)
    assert status == 0
    _, orc = load_oracle("full")
    orc = orc.reshape(7, NK, 17)
    assert_table_parity(tab, orc, load_floor("full"), what="full Time-RG vs oracle")
    assert not tab[:, :, 13:].any()  # columns 14-17 are zeros in full-TRG mode


def test_batch_of_identical_and_mixed_cosmologies(example1_dir, example1_full_dir):
    """Batching must not change any result: [1-loop, full, 1-loop] in one handle."""
    single, *_ = run(example1_dir)
    h = rt.RedTimeB200()
    h.add_cosmology(rt.read_run_dir(example1_dir))
    h.add_cosmology(rt.read_run_dir(example1_full_dir))
    h.add_cosmology(rt.read_run_dir(example1_dir))
    h.prepare()
    tables, hdr, hdr0, status = h.run()
    assert not status.any()
    assert np.array_equal(tables[0], single) and np.array_equal(tables[2], single)
    _, orc = load_oracle("full")
    assert np.all(col_err(tables[1][:, :, :10], orc.reshape(7, NK, 17)[:, :, :10]) < 1e-5)
    h.close()


def test_batch_front_end_writes_reference_format(tmp_path, golden_example1):
    """redtime_b200.batch: run directories in, redTime_<MODEL>.dat out (the file a
    `redTime > redTime_<MODEL>.dat` run of the reference leaves, scripts/runRedTime:196-229)."""
    from redtime_b200 import batch
    from conftest import make_example1_dir
    d1 = make_example1_dir(str(tmp_path / "M001"))
    d2 = make_example1_dir(str(tmp_path / "M002"), switches=[1, 0, 1, 1])
    res = batch.run_batch([d1, d2])
    assert res == {d1: 0, d2: 0}
    hdr, arr = parse_tables(open(os.path.join(d1, "redTime_M001.dat")).read())
    ghdr, gold = golden_example1
    assert hdr == ghdr
    assert np.max(np.abs(arr[:, :10] / gold[:, :10] - 1)) < 1e-9
    _, full = parse_tables(open(os.path.join(d2, "redTime_M002.dat")).read())
    _, orc = load_oracle("full")
    assert np.max(np.abs(full[:, :10] / orc[:, :10] - 1)) < 1e-9


def test_pinned_read_back_equals_copying_run(example1_dir):
    h = rt.RedTimeB200()
    h.add_cosmologies([rt.read_run_dir(example1_dir)] * 3)
    h.prepare()
    t1, hdr1, hdr01, st1 = h.run()
    t2, hdr2, hdr02, st2 = h.run_pinned()
    for a, b in zip(t1, t2):
        assert np.array_equal(a, b)
    assert np.array_equal(hdr1, hdr2) and np.array_equal(hdr01, hdr02)
    # a second prepare() re-sends the staged tables (the log transform consumed them)
    h.prepare()
    t3, *_ = h.run()
    assert np.array_equal(t3[1], t1[1])
    h.close()


def test_packed_batch_kernel_is_bit_identical(example1_dir, example1_full_dir):
    """Batches of >= 6 cosmologies go through k_bilinear_packed (three (cosmology, spectrum)
    slots per CTA); the per-slot arithmetic is the same, so nothing may change."""
    single, *_ = run(example1_dir)
    single_full, *_ = run(example1_full_dir)
    h = rt.RedTimeB200()
    dirs = [example1_dir, example1_full_dir, example1_dir, example1_dir, example1_full_dir, example1_dir,
            example1_dir]
    h.add_cosmologies([rt.read_run_dir(d) for d in dirs])
    h.prepare()
    tables, hdr, hdr0, status = h.run()
    assert not status.any()
    for d, t in zip(dirs, tables):
        assert np.array_equal(t, single if d == example1_dir else single_full)
    h.close()


def test_page_locked_inputs_take_the_direct_path(example1_dir):
    """Tables in page-locked memory are DMA'd as they are and beta is formed on the device;
    pageable ones are staged with beta formed on the host.  IEEE division on both sides: the
    results must be bit-identical, also in a batch that mixes the two and after a re-prepare."""
    import torch
    base = rt.read_run_dir(example1_dir)
    pinned = dict(base)
    keep = []
    for key in ("k_T", "Tc_T", "Tb_T", "k_b", "Tc_b", "Tnu_b"):
        t = torch.empty(base[key].shape, dtype=torch.float64, pin_memory=True)
        t.numpy()[...] = base[key]
        keep.append(t)
        pinned[key] = t.numpy()
    h = rt.RedTimeB200()
    h.add_cosmologies([base, pinned, base, pinned])
    h.prepare()
    tables, hdr, hdr0, status = h.run()
    assert not status.any()
    for t in tables[1:]:
        assert np.array_equal(t, tables[0])
    h.prepare()
    again, *_ = h.run()
    assert np.array_equal(again[1], tables[0]) and np.array_equal(again[2], tables[0])
    h.close()


def test_out_of_range_cosmology_is_flagged_not_fatal(example1_dir):
    """Where the reference abort()s (D_dD for a < 1e-3: z_in = 1500; Beta_P for a > 1.001:
    z_out = -0.01) the cosmology gets a status word and the rest of the batch is untouched."""
    good = rt.read_run_dir(example1_dir)
    single, *_ = run(example1_dir)
    bad1 = dict(good, z_in=1500.0)
    bad2 = dict(good, z_out=np.array([1.0, -0.01]))
    h = rt.RedTimeB200()
    h.add_cosmologies([good, bad1, good, bad2])
    h.prepare()
    with pytest.raises(rt.RtrgError) as e:
        h.run()
    assert e.value.code == -5
    tables, hdr, hdr0, status = h.run(raise_on_ode_failure=False)
    assert list(status) == [0, 103, 0, 103]
    assert np.array_equal(tables[0], single) and np.array_equal(tables[2], single)
    assert not tables[1].any() and not tables[3].any()
    h.close()


def test_reduced_beta_upload(example1_dir, example1_full_dir):
    """reduce_beta = 1: the host sends beta(a=1,k) and the table pre-reduced in k instead of the
    full Beta_P(a,k) table.  Same interpolation rules in the other order: results
    agree to round-off (host code has no FMA contraction, the device code has)."""
    for d, tag in ((example1_dir, "1loop"), (example1_full_dir, "full")):
        full, hdr_f, hdr0_f, _, cnt_f = run(d)
        red, hdr_r, hdr0_r, _, cnt_r = run(d, reduce_beta=1)
        assert cnt_f == cnt_r
        # 1e-16 differences of beta are amplified by the cancelling mode-coupling sums inside the
        # full Time-RG right-hand side: measured 1.2e-10 on columns 8-10 (1-loop mode: < 1e-13)
        assert np.max(np.abs(red[:, :, :7] - full[:, :, :7]) / np.abs(full[:, :, :7]).clip(1e-300)) < 1e-12
        assert np.max(np.abs(red[:, :, 7:10] - full[:, :, 7:10]) / np.abs(full[:, :, 7:10])) < 2e-9
        # columns 11-17 amplify a 1-ulp change of the inputs by up to 1e8 at the lowest k: bounded at
        # every k by 1e-6 of the local scale + the reference's measured response to a 1-ulp change
        allowed = 1e-6 * local_scale(full) + FLOOR_C * smooth_floor(load_floor(tag))
        assert np.all(np.abs(red - full)[:, :, 10:] <= allowed[:, :, 10:])
        assert np.allclose(hdr_r[:7], hdr_f[:7], rtol=1e-13, atol=0) and np.allclose(hdr0_r, hdr0_f, rtol=1e-13, atol=0)
    h = rt.RedTimeB200(reduce_beta=1)
    h.add_cosmology(rt.read_run_dir(example1_dir))
    h.prepare()
    k = np.array([1e-3, 0.05, 1.0, 3.0])
    ref = rt.RedTimeB200()
    ref.add_cosmology(rt.read_run_dir(example1_dir))
    ref.prepare()
    assert np.max(np.abs(h.Beta_P(1.0, k) / ref.Beta_P(1.0, k) - 1)) < 1e-14
    with pytest.raises(rt.RtrgError):
        h.Beta_P(0.5, k)   # only a = 1 was uploaded
    h.close()
    ref.close()
