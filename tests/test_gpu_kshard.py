"""k-sharded single cosmology: G handles, each owning nk/G rows, exchanging the
ln P rows before every integral evaluation and max-reducing the error norm.  Tested on ONE GPU
with the in-process loopback transport (one host thread per rank; no kernel waits on another
kernel) -- the NCCL transport differs only in how the blocks move.  Sharding must not change a
single bit: every row is computed by the same CTA arithmetic and max() is order independent."""
import threading

import numpy as np
import pytest

import redtime_b200 as rt

pytestmark = pytest.mark.gpu


def run_single(d, **cfg):
    h = rt.RedTimeB200(**cfg)
    h.add_cosmology(rt.read_run_dir(d))
    h.prepare()
    tables, hdr, hdr0, status = h.run()
    cnt = h.counters(0)
    h.close()
    return tables[0], hdr[0], cnt


def run_sharded(d, G, **cfg):
    group = rt.LoopbackGroup(G)
    inp = rt.read_run_dir(d)
    hs = [rt.RedTimeB200(k_shards=G, k_rank=r, **cfg) for r in range(G)]
    res, errs = [None] * G, []

    def work(r):
        try:
            hs[r].add_cosmology(inp)
            hs[r].kshard_init_loopback(group)
            hs[r].prepare()
            res[r] = hs[r].run()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(r,)) for r in range(G)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    assert not errs, errs
    assert all(r is not None for r in res), "a rank did not finish"
    cnt = hs[0].counters(0)
    [h.close() for h in hs]
    group.close()
    return res, cnt


@pytest.mark.parametrize("G", [2, 4])
def test_kshard_full_trg_is_bit_identical(G, example1_full_dir):
    ref, hdr, cnt0 = run_single(example1_full_dir)
    res, cnt = run_sharded(example1_full_dir, G)
    assert (cnt["attempts"], cnt["rejected"]) == (cnt0["attempts"], cnt0["rejected"])
    for r in range(G):
        tables, hdr_r, hdr0, status = res[r]
        assert not status.any()
        assert np.array_equal(tables[0], ref), "rank %d" % r   # every rank holds the full table
        assert np.array_equal(hdr_r[0], hdr)


def test_kshard_1loop_is_bit_identical(example1_dir):
    ref, hdr, cnt0 = run_single(example1_dir)
    res, cnt = run_sharded(example1_dir, 2)
    assert (cnt["attempts"], cnt["rejected"]) == (23, 4)
    for r in range(2):
        assert np.array_equal(res[r][0][0], ref)


def test_kshard_needs_a_transport(example1_dir):
    h = rt.RedTimeB200(k_shards=2, k_rank=0)
    h.add_cosmology(rt.read_run_dir(example1_dir))
    h.prepare()
    with pytest.raises(rt.RtrgError):
        h.run()
    h.close()
    with pytest.raises(rt.RtrgError):
        rt.RedTimeB200(k_shards=3, k_rank=0)  # nk/8 = 16 rows blocks are not divisible by 3


def test_v_split_changes_only_round_off(example1_full_dir):
    """v_split > 1 splits the beta-side lags of a row block over several CTAs (shorter serial
    chains for tiny grids); partial sums are added in another order."""
    ref, hdr, cnt0 = run_single(example1_full_dir)
    for vs in (2, 4, 16):
        tab, _, cnt = run_single(example1_full_dir, v_split=vs)
        assert (cnt["attempts"], cnt["rejected"]) == (cnt0["attempts"], cnt0["rejected"])
        assert np.max(np.abs(tab[:, :, :10] / ref[:, :, :10] - 1)) < 2e-9
    # and sharding with a split stays bit-identical to the unsharded run with the same split
    ref4, _, _ = run_single(example1_full_dir, v_split=4)
    res, _ = run_sharded(example1_full_dir, 2, v_split=4)
    assert np.array_equal(res[0][0][0], ref4) and np.array_equal(res[1][0][0], ref4)


def test_fused_stage_kernel_is_bit_identical(example1_full_dir, monkeypatch):
    """Small launches run assembly + right-hand side + next-stage combination (or the error estimate)
    as ONE kernel (k_stage_post) built from the device functions of k_assemble / k_rhs / k_combine /
    k_final.  Switching it off (RTRG_NO_STAGE_FUSION, read by every rtrg_run) must not change a bit,
    sharded or not."""
    fused, _, cnt_f = run_single(example1_full_dir)
    res_f, _ = run_sharded(example1_full_dir, 2)
    monkeypatch.setenv("RTRG_NO_STAGE_FUSION", "1")
    plain, _, cnt_p = run_single(example1_full_dir)
    res_p, _ = run_sharded(example1_full_dir, 2)
    assert cnt_f == cnt_p
    assert np.array_equal(fused, plain)
    for r in range(2):
        assert np.array_equal(res_f[r][0][0], plain) and np.array_equal(res_p[r][0][0], plain)
