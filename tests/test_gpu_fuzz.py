"""Randomised GPU parity: seeded cohorts of random shape (region length, haplotype count,
variant density, indel mix, PAM / guide length / side) through every route of the host layer
-- resident search, streamed search from texts and from edit lists with a random number of
groups, N2 annotation -- against the C oracle of the scan and the N2 oracle. Bit-exact."""

import os

import numpy as np
import pytest

from crispr_hawk_b200 import synth
from crispr_hawk_b200.workload import Workload
from oracle import annot_oracle as A
from oracle import c_oracle

pytestmark = pytest.mark.gpu

COLS = ("hap", "strand", "pos", "start", "stop")
PAMS = [("NGG", 20, False), ("TTTV", 23, True), ("NNGRRT", 21, False), ("NG", 18, False), ("YTN", 12, True),
        ("NNNNGATT", 22, False), ("NGG", 34, False), ("TTTV", 45, True)]  # fmt: skip


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    bed_len = int(rng.integers(300, 30_000))
    n_alt = int(rng.integers(0, 14))
    n_sites = int(rng.integers(0, max(1, bed_len // 12)))
    mean = float(rng.uniform(0, max(1.0, n_sites * 0.6)))
    snv = float(rng.uniform(0.2, 1.0))
    ins = float(rng.uniform(0, 1.0 - snv))
    max_indel = int(rng.integers(1, 12))
    for bump in range(50):
        c = synth.make_cohort(bed_len, n_alt, n_sites, mean, seed=2000 + seed + 1000 * bump, snv_frac=snv, ins_frac=ins,
                              max_indel=max_indel)  # fmt: skip
        # a non-reference haplotype without a single variant would be a second REF (the library
        # refuses those like the reference does): draw again
        if np.all(np.diff(c.hap_off)[1:] > 0):
            break
        n_alt = max(0, n_alt - 1)
    pam, G, right = PAMS[int(rng.integers(0, len(PAMS)))]
    return c, pam, G, right, rng


@pytest.mark.parametrize("seed", range(int(os.environ.get("HAWK_FUZZ_SEEDS", "16"))))
def test_random_cohort_all_routes(seed):
    c, pam, G, right, rng = _case(seed)
    wl = Workload(c, pam, G, right)
    res = wl.step_resident()
    table = res.table()
    res.close()
    buf, off, lens, a, b, is_ref, seg = wl.host_arrays_for_oracle(np.arange(c.n_hap))
    want = c_oracle.search(buf, off, lens, a, b, is_ref, seg, wl.fwd, wl.rc, G, right, threads=2)
    order = np.argsort(table["bucket"], kind="stable")
    assert len(order) == len(want["hap"]), (pam, G, right)
    for col in COLS:
        assert np.array_equal(table[col][order], want[col]), col
    assert np.array_equal(table["text"][order], want["text"])
    # streamed routes give the same table, bucket ids included
    for step in (wl.step_host, wl.step_edits):
        got, _, _ = step(n_groups=int(rng.integers(0, 6)))
        for col in COLS + ("bucket",):
            assert np.array_equal(got[col], table[col]), (step.__name__, col)
        assert np.array_equal(got["text"], table["text"]), step.__name__
    # N2 on a sample of rows against the oracle (inside annotate_measure)
    if len(table["hap"]):
        m = wl.annotate_measure(reps=1, n_sample=64, oracle=A)
        assert m["rows"] == len(table["hap"])
