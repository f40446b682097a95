"""GPU parity for BASELINE config 4 (unphased, gnomAD density, SaCas9 NNGRRT / 21 nt) through
the C-ABI: the 8 x 5 kb slices built and searched by the reference itself
(tests/golden/config4_slices.json.gz), a 100 kb cohort row for row against the C oracle, and
the full 10 Mb configuration through size-independent properties plus an oracle subset."""

import numpy as np
import pytest

import crispr_hawk_b200 as hawk
from crispr_hawk_b200 import synth
from crispr_hawk_b200 import synth_unphased as SU
from crispr_hawk_b200.workload import UnphasedWorkload
from oracle import c_oracle
from tests.helpers import FxRegion
from tests.test_gpu_workload import COLS, assert_tables_equal, check_bucket_ids, final_order
from tests.test_synth_unphased import SLICES, slice_inputs, table_digest

pytestmark = pytest.mark.gpu


def oracle_for(wl, u, a, b, threads=4):
    return c_oracle.search(u.ascii, u.slot_off, u.lens, a, b, u.is_ref, u.seg, wl.fwd, wl.rc, wl.guidelen, wl.right,
                           threads=threads, unphased=True, alleles=u.alleles)  # fmt: skip


@pytest.mark.parametrize("spec", SLICES, ids=[f"c4_slice_{s['seed']}" for s in SLICES])
def test_c4_slices_match_the_reference(spec):
    c, u = slice_inputs(spec)
    wl = UnphasedWorkload(c, spec["pam"], spec["guidelen"], spec["right"], uset=u)
    res = wl.step_resident()
    table = res.table()
    res.close()
    check_bucket_ids(table)
    got = final_order(table)
    assert len(got["hap"]) == spec["n_guides"]
    digest, sample = table_digest(spec, got, spec["hap_order"])
    assert sample == [list(x) for x in spec["sample"]]
    assert digest == spec["sha256"]
    assert_tables_equal(got, oracle_for(wl, u, wl.a, wl.b), f"slice {spec['seed']}")
    # host-buffer path
    t2, h2d, d2h = wl.step_host()
    for k in COLS + ("bucket",):
        assert np.array_equal(t2[k], table[k])
    assert np.array_equal(t2["text"], table["text"])
    assert h2d > u.total_slots and d2h > 0


def test_c4_slice_through_the_python_api():
    """`crispr_hawk_b200.search` on haplotype OBJECTS (the drop-in seam) for one slice."""
    spec = SLICES[0]
    c, u = slice_inputs(spec)
    haps = SU.unphased_haplotypes(c, u)
    for k, h in enumerate(haps):
        h.id = f"h{spec['hap_order'][k]}"
    region = FxRegion("chr1", c.region_start, c.region_stop)
    pam = hawk.PAM(spec["pam"], spec["right"], True)
    pam.encode(0)
    bits = hawk.encode_region(haps, 0, True)
    guides = hawk.search(pam, region, haps, bits, spec["guidelen"], spec["right"], True, False, 0, True)
    assert len(guides) == spec["n_guides"]
    for row in spec["sample"]:
        g = guides[row[0]]
        assert [g.start, g.stop, g.strand, g.sequence, bool(g.right), int(g.hapid[1:])] == row[1:]


@pytest.mark.parametrize("scale", [0.002, 0.01])
def test_c4_scaled_matches_c_oracle(scale):
    k = synth.CONFIGS["c4"]
    c = synth.config_cohort("c4", scale)
    wl = UnphasedWorkload(c, k["pam"], k["guidelen"], k["right"])
    res = wl.step_resident()
    table = res.table()
    assert res.scanned_bp == wl.scanned_bp
    res.close()
    check_bucket_ids(table)
    want = oracle_for(wl, wl.d, wl.a, wl.b)
    assert want["scanned_bp"] == wl.scanned_bp
    assert_tables_equal(final_order(table), want, f"c4 x {scale}")
    assert len(table["hap"]) > 1000
    for _ in range(2):  # idempotent
        r = wl.step_resident()
        t = r.table()
        r.close()
        for kcol in COLS + ("bucket",):
            assert np.array_equal(t[kcol], table[kcol])
        assert np.array_equal(t["text"], table["text"])


def test_c4_full_size_properties():
    """BASELINE config 4 at full size (10 Mb, ~1.25 M sites, 11 full-length IUPAC haplotypes +
    ~4.3e5 indel-window haplotypes): emission order, bucket ids, and REF + one SNV haplotype +
    3,000 random window haplotypes row for row against the C oracle."""
    k = synth.CONFIGS["c4"]
    c = synth.config_cohort("c4")
    wl = UnphasedWorkload(c, k["pam"], k["guidelen"], k["right"])
    res = wl.step_resident()
    table = res.table()
    res.close()
    n = len(table["hap"])
    assert wl.scanned_bp > 1.9e8 and n > 1e7
    key = (table["hap"].astype(np.int64) << 33) | (table["strand"].astype(np.int64) << 32) | table["pos"].astype(np.int64)
    assert np.all(np.diff(key) >= 0)  # resolved strings of one hit share (hap, strand, pos)
    check_bucket_ids(table)
    rng = np.random.default_rng(0)
    n_full = int((wl.d.hap_kind < 2).sum())
    subset = np.sort(np.concatenate(([0, int(rng.integers(1, n_full))],
                                    rng.choice(np.arange(n_full, wl.d.n_hap), 3000, replace=False))))  # fmt: skip
    u, a, b = wl.oracle_subset(subset)
    want = oracle_for(wl, u, a, b, threads=8)
    okey = (want["hap"].astype(np.int64) << 33) | (want["strand"].astype(np.int64) << 32) | want["pos"].astype(np.int64)
    oo = np.argsort(okey, kind="stable")
    sel = np.flatnonzero(np.isin(table["hap"], subset))
    got = {kcol: table[kcol][sel] for kcol in COLS + ("text",)}
    got["hap"] = np.searchsorted(subset, got["hap"]).astype(np.int32)
    assert_tables_equal(got, {kcol: want[kcol][oo] for kcol in COLS + ("text",)}, "c4 subset")
