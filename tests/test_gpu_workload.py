"""GPU parity at workload scale: device-materialised synthetic cohorts of the BASELINE.json
shapes through the C-ABI (K1 pack -> K2 scan -> post-scan pipeline) against the C
restatement of the reference scan (oracle/scan_oracle.c, itself pinned to the reference's
golden vectors by tests/test_oracle_c.py). Bit-exact: every row, every column, in order."""

import numpy as np
import pytest

from crispr_hawk_b200 import _cabi, marshal, synth
from crispr_hawk_b200.workload import Workload
from oracle import c_oracle

pytestmark = pytest.mark.gpu

COLS = ("hap", "strand", "pos", "start", "stop")


def final_order(table):
    order = np.argsort(table["bucket"], kind="stable")
    return {k: table[k][order] for k in COLS + ("text",)}


def assert_tables_equal(got, want, what=""):
    assert len(got["hap"]) == len(want["hap"]), f"{what}: row count {len(got['hap'])} != {len(want['hap'])}"
    for k in COLS:
        bad = np.flatnonzero(got[k] != want[k])
        assert len(bad) == 0, f"{what}: column {k} differs first at row {bad[0]}"
    assert np.array_equal(got["text"], want["text"]), f"{what}: window text differs"


def oracle_table(wl, hap_indices, threads=4):
    buf, off, lens, a, b, is_ref, seg = wl.host_arrays_for_oracle(hap_indices)
    return c_oracle.search(buf, off, lens, a, b, is_ref, seg, wl.fwd, wl.rc, wl.guidelen, wl.right, threads=threads)


def check_bucket_ids(table):
    """bucket[i] == smallest emission index sharing (start, strand) (group_guides_position)."""
    key = table["start"].astype(np.int64) * 2 + table["strand"]
    _, first, inv = np.unique(key, return_index=True, return_inverse=True)
    assert np.array_equal(table["bucket"], first[inv])


def test_device_materialisation_matches_host():
    c = synth.make_cohort(bed_len=40_000, n_alt_hap=9, n_sites=900, mean_alts_per_hap=120, seed=5,
                          snv_frac=0.6, ins_frac=0.2, max_indel=7)  # fmt: skip
    d = synth.derive(c)
    dev = synth.materialize_device(c).cpu().numpy()
    texts = synth.materialize_host(c)
    buf, off, lens = marshal.stage_ascii(texts)
    assert off.tolist() == d.slot_off.tolist() and lens.tolist() == d.lens.tolist()
    assert np.array_equal(dev, buf)


@pytest.mark.parametrize("name,scale,n_alt", [("c1", 1.0, 20), ("c2", 0.2, 23), ("c3", 0.2, 23), ("c5shard", 0.004, 7)])
def test_workload_table_matches_c_oracle(name, scale, n_alt):
    k = synth.CONFIGS[name]
    c = synth.config_cohort(name, scale, n_alt_hap=n_alt)
    wl = Workload(c, k["pam"], k["guidelen"], k["right"])
    res = wl.step_resident()
    table = res.table()
    hits = [res.hits(0), res.hits(1)]
    assert res.scanned_bp == wl.scanned_bp
    res.close()
    check_bucket_ids(table)
    want = oracle_table(wl, np.arange(c.n_hap))
    assert want["scanned_bp"] == wl.scanned_bp
    assert_tables_equal(final_order(table), want, name)
    assert len(table["hap"]) > 100
    # host-buffer path gives the same table
    t2, h2d, d2h = wl.step_host_twocall()
    for kcol in COLS + ("bucket",):
        assert np.array_equal(t2[kcol], table[kcol])
    assert np.array_equal(t2["text"], table["text"])
    assert h2d > wl.d.total_slots and d2h >= len(table["hap"]) * (21 + (wl.guidelen + len(wl.fwd) + 20 + 15) // 16 * 16)
    # raw pam_search semantics on the same batch (re-encoded in full: a fused search leaves it sparse)
    wl.batch.repack(wl.ascii_dev.data_ptr())
    raw = _cabi.pam_search(wl.ctx, wl.batch, wl.params, wl.a, wl.b)
    for s in (0, 1):
        assert np.array_equal(raw.hits(s), want["hits"][s])
        # filtered hit list is a subset of the raw one, ascending
        assert np.all(np.diff(hits[s].astype(np.int64)) > 0)
        assert np.isin(hits[s], want["hits"][s]).all()
    raw.close()


def test_repeated_steps_are_idempotent():
    k = synth.CONFIGS["c2"]
    c = synth.config_cohort("c2", 0.05, n_alt_hap=40)
    wl = Workload(c, k["pam"], k["guidelen"], k["right"])
    first = wl.step_resident()
    t1 = first.table()
    first.close()
    for _ in range(3):
        r = wl.step_resident()
        t = r.table()
        r.close()
        for kcol in COLS + ("bucket",):
            assert np.array_equal(t[kcol], t1[kcol])
        assert np.array_equal(t["text"], t1["text"])


@pytest.mark.parametrize("name,min_rows", [("c2", 1e7), ("c3", 2e6)])
def test_full_size_config_properties(name, min_rows):
    """BASELINE.json configs 2 and 3 at full size (1 Mb x 5,009 haplotypes = 5.0 G hap-bp; NGG /
    20 nt / left and TTTV / 23 nt / right): the oracle checks a random subset of haplotypes row
    by row; the whole table is checked through size-independent properties (emission order,
    bucket ids)."""
    k = synth.CONFIGS[name]
    c = synth.config_cohort(name)
    wl = Workload(c, k["pam"], k["guidelen"], k["right"])
    res = wl.step_resident()
    table = res.table()
    res.close()
    assert wl.scanned_bp > 5.0e9
    n = len(table["hap"])
    assert n > min_rows
    # emission order: (hap, strand, pos) strictly ascending
    key = (table["hap"].astype(np.int64) << 33) | (table["strand"].astype(np.int64) << 32) | table["pos"].astype(np.int64)
    assert np.all(np.diff(key) > 0)
    check_bucket_ids(table)
    rng = np.random.default_rng(0)
    subset = np.sort(np.concatenate(([0], rng.choice(np.arange(1, c.n_hap), 12, replace=False))))
    want = oracle_table(wl, subset, threads=8)
    # oracle rows (final order of the subset) -> emission order
    okey = (want["hap"].astype(np.int64) << 33) | (want["strand"].astype(np.int64) << 32) | want["pos"].astype(np.int64)
    oo = np.argsort(okey, kind="stable")
    sel = np.flatnonzero(np.isin(table["hap"], subset))
    got = {kcol: table[kcol][sel] for kcol in COLS + ("text",)}
    got["hap"] = np.searchsorted(subset, got["hap"]).astype(np.int32)
    assert_tables_equal(got, {kcol: want[kcol][oo] for kcol in COLS + ("text",)}, f"{name} subset")


def _run_against_oracle(c, pam, G, right):
    wl = Workload(c, pam, G, right)
    res = wl.step_resident()
    table = res.table()
    res.close()
    check_bucket_ids(table)
    want = oracle_table(wl, np.arange(c.n_hap))
    assert_tables_equal(final_order(table), want, f"{pam}/{G}")
    return wl, table


def test_dense_hits_force_the_exact_retry():
    """PAM `N` hits every position on both strands: far more records than the staging
    estimate, so K2 is re-run with the exact per-warp segment sizes (and a larger staging
    area). The table must still be bit-exact."""
    c = synth.make_cohort(bed_len=150_000, n_alt_hap=5, n_sites=1500, mean_alts_per_hap=300, seed=8,
                          snv_frac=0.7, ins_frac=0.15, max_indel=6)  # fmt: skip
    wl, table = _run_against_oracle(c, "N", 20, False)
    assert len(table["hap"]) > 2 * 150_000 - 200  # every REF position, both strands
    wl.batch.repack(wl.ascii_dev.data_ptr())  # the raw scan reads every chunk: dense planes
    raw = _cabi.pam_search(wl.ctx, wl.batch, wl.params, wl.a, wl.b)
    assert raw.n_hits[0] == raw.n_hits[1] == wl.scanned_bp
    raw.close()


@pytest.mark.parametrize("pam,G,right", [("NGG", 40, False), ("TTTV", 70, True), ("NNGRRT", 33, False)])
def test_long_guides_take_the_generic_path(pam, G, right):
    """G > 32 or G + P > 33 leaves the 96-bit case-window fast path of the scan kernel."""
    c = synth.make_cohort(bed_len=60_000, n_alt_hap=9, n_sites=900, mean_alts_per_hap=150, seed=12,
                          snv_frac=0.6, ins_frac=0.2, max_indel=8)  # fmt: skip
    _, table = _run_against_oracle(c, pam, G, right)
    assert len(table["hap"]) > 500


def test_sacas9_and_short_guides():
    c = synth.make_cohort(bed_len=80_000, n_alt_hap=12, n_sites=1200, mean_alts_per_hap=200, seed=13,
                          snv_frac=0.7, ins_frac=0.15, max_indel=10)  # fmt: skip
    for pam, G, right in (("NNGRRT", 21, False), ("TTN", 23, True), ("NGG", 1, False), ("G", 5, True)):
        _run_against_oracle(c, pam, G, right)


def test_full_size_edit_lists_give_the_table_of_the_texts():
    """BASELINE config 2 at full size (5,009 x 1 Mb): the search from edit lists -- planes built
    only around the 4.2 M edits, texts never written (hawk_search_stream_edits) -- returns the
    table of the search over the resident texts, every column of all 14.66 M rows."""
    k = synth.CONFIGS["c2"]
    wl = Workload(synth.config_cohort("c2"), k["pam"], k["guidelen"], k["right"])
    res = wl.step_resident()
    want = res.table()
    res.close()
    got, h2d, d2h = wl.step_edits()
    assert len(got["hap"]) == len(want["hap"]) > 1e7
    for kcol in COLS + ("bucket",):
        assert np.array_equal(got[kcol], want[kcol]), kcol
    assert np.array_equal(got["text"], want["text"])
    assert h2d < 0.05 * wl.d.total_slots  # edit lists, not texts, crossed PCIe (the first call runs twice: it sizes its buffers)
    slim, _, d2h_slim = wl.step_edits(want_text=False)
    for kcol in COLS + ("bucket",):
        assert np.array_equal(slim[kcol], want[kcol]), kcol
    assert d2h_slim < 0.35 * d2h


@pytest.mark.parametrize("k", range(8))
def test_c5_slices_match_c_oracle(k):
    """BASELINE.json config 5 at the parity size SURVEY.md 8(d) names: 8 sub-regions of 20 kb x 32
    haplotypes with config 5's variant model (one alternate allele per kb and haplotype, SNV /
    insertion / deletion 90 / 5 / 5, indels up to 10 bases), every row against the C oracle."""
    cfg = synth.CONFIGS["c5shard"]
    c = synth.config_cohort("c5shard", 20_000 / cfg["bed_len"], seed_offset=100 + k, n_alt_hap=32)
    assert c.n_hap == 33
    _run_against_oracle(c, cfg["pam"], cfg["guidelen"], cfg["right"])


def test_full_size_config5_block():
    """BASELINE.json config 5, one rank's share at full size: a 50 Mb region x 626 haplotypes
    (31.3 G haplotype-bp, ~97 M guide rows). Whole table: emission order, bucket ids = first row
    of the (start, strand) key; REF rows per strand against a numpy restatement of the PAM test on
    the reference text; REF + two random haplotypes row by row against the C oracle."""
    cfg = synth.CONFIGS["c5shard"]
    c = synth.config_cohort("c5shard")
    wl = Workload(c, cfg["pam"], cfg["guidelen"], cfg["right"])
    res = wl.step_resident()
    table = res.table()
    res.close()
    assert wl.scanned_bp > 31.0e9
    n = len(table["hap"])
    assert n > 5e7
    key = (table["hap"].astype(np.int64) << 33) | (table["strand"].astype(np.int64) << 32) | table["pos"].astype(np.int64)
    assert np.all(np.diff(key) > 0)
    del key
    # bucket ids through a direct-address table (np.unique over 1e8 keys would take minutes)
    k2 = table["start"].astype(np.int64) * 2 + table["strand"]
    k2 -= k2.min()
    first = np.full(int(k2.max()) + 1, n, np.int64)
    order = np.arange(n - 1, -1, -1)
    first[k2[order]] = order  # later writes win: the smallest row index of every key stays
    assert np.array_equal(table["bucket"], first[k2])
    del k2, first, order
    # REF rows per strand == NGG / CCN occurrences whose PAM lies inside the BED interval
    ref = c.ref
    B = cfg["bed_len"]
    lo, hi = 100, 100 + B - 3  # compute_scan_start_stop (search_guides.py:49-84) on REF
    g = ref == ord("G")
    cc = ref == ord("C")
    fwd = int(np.count_nonzero(g[lo + 1 : hi + 1] & g[lo + 2 : hi + 2]))
    rev = int(np.count_nonzero(cc[lo:hi] & cc[lo + 1 : hi + 1]))
    ref_rows = table["hap"] == 0
    assert int(np.count_nonzero(ref_rows & (table["strand"] == 0))) == fwd
    assert int(np.count_nonzero(ref_rows & (table["strand"] == 1))) == rev
    rng = np.random.default_rng(5)
    subset = np.sort(np.concatenate(([0], rng.choice(np.arange(1, c.n_hap), 2, replace=False))))
    want = oracle_table(wl, subset, threads=8)
    okey = (want["hap"].astype(np.int64) << 33) | (want["strand"].astype(np.int64) << 32) | want["pos"].astype(np.int64)
    oo = np.argsort(okey, kind="stable")
    sel = np.flatnonzero(np.isin(table["hap"], subset))
    got = {kcol: table[kcol][sel] for kcol in COLS + ("text",)}
    got["hap"] = np.searchsorted(subset, got["hap"]).astype(np.int32)
    assert_tables_equal(got, {kcol: want[kcol][oo] for kcol in COLS + ("text",)}, "c5 block subset")
