"""hawk_encode_search_dev (K1 + K2 fused, flat over the slot space) must give the table of
hawk_batch_repack_dev + hawk_search (staged K2) bit for bit, and both equal the oracle in
tests/test_gpu_workload.py / test_gpu_unphased.py (which run the fused call by default)."""

import numpy as np
import pytest

from crispr_hawk_b200 import _cabi, synth
from crispr_hawk_b200.workload import UnphasedWorkload, Workload
from tests.test_gpu_workload import COLS

pytestmark = pytest.mark.gpu


def tables_equal(t1, t2):
    for k in COLS + ("bucket",):
        assert np.array_equal(t1[k], t2[k]), k
    assert np.array_equal(t1["text"], t2["text"])


def both_ways(wl):
    r = wl.step_resident(fused=False)
    staged, hs = r.table(), [r.hits(0), r.hits(1)]
    r.close()
    r = wl.step_resident(fused=True)
    fused, hf = r.table(), [r.hits(0), r.hits(1)]
    r.close()
    tables_equal(staged, fused)
    for s in (0, 1):
        assert np.array_equal(hs[s], hf[s])
    # the library's own choice by haplotype shape
    r = wl.step_resident()
    auto, ha = r.table(), [r.hits(0), r.hits(1)]
    r.close()
    tables_equal(staged, auto)
    for s in (0, 1):
        assert np.array_equal(hs[s], ha[s])
    # and again: a sparse batch re-encodes cleanly, the staged path works after a fused one
    r = wl.step_resident(fused=True)
    tables_equal(staged, r.table())
    r.close()
    r = wl.step_resident(fused=False)
    tables_equal(staged, r.table())
    r.close()
    return staged


@pytest.mark.parametrize("name,scale,n_alt", [("c1", 1.0, 20), ("c2", 0.2, 23), ("c3", 0.2, 23), ("c5shard", 0.004, 7),
                                              ("c2", 0.013, 300)])  # fmt: skip
def test_fused_equals_staged_phased(name, scale, n_alt):
    k = synth.CONFIGS[name]
    wl = Workload(synth.config_cohort(name, scale, n_alt_hap=n_alt), k["pam"], k["guidelen"], k["right"])
    assert len(both_ways(wl)["hap"]) > 100


@pytest.mark.parametrize("scale", [0.002, 0.02])
def test_fused_equals_staged_unphased(scale):
    k = synth.CONFIGS["c4"]
    wl = UnphasedWorkload(synth.config_cohort("c4", scale), k["pam"], k["guidelen"], k["right"])
    assert len(both_ways(wl)["hap"]) > 1000


@pytest.mark.parametrize("pam,G,right", [("N", 20, False), ("NNGRRT", 21, False), ("TTN", 23, True), ("NGG", 1, False),
                                         ("G", 5, True), ("NGG", 32, False), ("NGG", 30, True), ("NGG", 40, False)])  # fmt: skip
def test_fused_equals_staged_geometries(pam, G, right):
    """Every guide / PAM geometry of the fast scan form (G <= 32, G + P <= 33), the widest reach
    included, and one beyond it (the fused call then runs K1 and the staged K2)."""
    c = synth.make_cohort(bed_len=70_000, n_alt_hap=9, n_sites=1100, mean_alts_per_hap=160, seed=31,
                          snv_frac=0.6, ins_frac=0.2, max_indel=8)  # fmt: skip
    both_ways(Workload(c, pam, G, right))


def test_dense_variants_overflow_falls_back():
    """Variants every ~16 bases: more hit chunks than a non-REF segment holds, so the fused
    kernel reports the overflow and the staged K2 finishes on the planes it kept."""
    c = synth.make_cohort(bed_len=600_000, n_alt_hap=3, n_sites=40_000, mean_alts_per_hap=36_000, seed=33,
                          snv_frac=1.0, ins_frac=0.0, max_indel=1)  # fmt: skip
    t = both_ways(Workload(c, "N", 20, False))
    assert len(t["hap"]) > 1_000_000


def test_sparse_batch_refuses_what_it_cannot_serve():
    k = synth.CONFIGS["c2"]
    wl = Workload(synth.config_cohort("c2", 0.02, n_alt_hap=5), k["pam"], k["guidelen"], k["right"])
    wl.step_resident(fused=True).close()
    with pytest.raises(_cabi.HawkLibraryError):
        _cabi.pam_search(wl.ctx, wl.batch, wl.params, wl.a, wl.b)
    with pytest.raises(_cabi.HawkLibraryError):
        wl.batch.export_nibbles(1)
    long_params = _cabi.make_params(wl.fwd, wl.rc, 32, False, False)
    with pytest.raises(_cabi.HawkLibraryError):
        _cabi.search(wl.ctx, wl.batch, long_params, wl.a, wl.b, wl.d.is_ref)
    # a shorter guide is served from the kept planes
    short = _cabi.make_params(wl.fwd, wl.rc, 12, False, False)
    r1 = _cabi.search(wl.ctx, wl.batch, short, wl.a, wl.b, wl.d.is_ref)
    t1 = r1.table()
    r1.close()
    wl.batch.repack(wl.ascii_dev.data_ptr())
    r2 = _cabi.search(wl.ctx, wl.batch, short, wl.a, wl.b, wl.d.is_ref)
    tables_equal(t1, r2.table())
    r2.close()


def test_fused_reports_a_bad_character():
    k = synth.CONFIGS["c2"]
    wl = Workload(synth.config_cohort("c2", 0.02, n_alt_hap=5), k["pam"], k["guidelen"], k["right"])
    wl.prepare_resident()
    slot = int(wl.d.slot_off[3]) + 777
    keep = int(wl.ascii_dev[slot])
    wl.ascii_dev[slot] = ord("!")
    with pytest.raises(_cabi.HawkLibraryError) as e:
        wl.step_resident(fused=True)
    assert e.value.code == _cabi.HAWK_EIUPAC and e.value.bad_slot == slot
    wl.ascii_dev[slot] = keep
    wl.step_resident(fused=True).close()

