"""CPU-side randomised parity: the kernels' host/device core compiled for the CPU
(tests/hostcheck.py) against the C oracle of the reference scan on seeded random cohorts --
the same generator and shapes the GPU fuzz test uses, small enough for the CPU suite."""

import numpy as np
import pytest

from crispr_hawk_b200 import marshal, synth
from crispr_hawk_b200.pam import pam_patterns
from oracle import c_oracle
from tests import hostcheck

PAMS = [("NGG", 20, False), ("TTTV", 23, True), ("NNGRRT", 21, False), ("NG", 18, False), ("YTN", 12, True),
        ("NNNNGATT", 22, False), ("NGG", 34, False), ("TTTV", 45, True)]  # fmt: skip


@pytest.mark.parametrize("seed", range(24))
def test_kernel_core_on_cpu_matches_c_oracle(seed):
    rng = np.random.default_rng(3000 + seed)
    bed_len = int(rng.integers(300, 4000))
    n_alt = int(rng.integers(1, 8))
    n_sites = int(rng.integers(1, max(2, bed_len // 12)))
    snv = float(rng.uniform(0.2, 1.0))
    ins = float(rng.uniform(0, 1.0 - snv))
    for bump in range(50):
        c = synth.make_cohort(bed_len, n_alt, n_sites, float(rng.uniform(1, max(2.0, n_sites * 0.6))),
                              seed=4000 + seed + 1000 * bump, snv_frac=snv, ins_frac=ins, max_indel=int(rng.integers(1, 12)))  # fmt: skip
        if np.all(np.diff(c.hap_off)[1:] > 0):
            break
    pam, G, right = PAMS[int(rng.integers(0, len(PAMS)))]
    haps = synth.synth_haplotypes(c)
    region = synth.SynthRegion(c)
    tab = hostcheck.search(pam, region, haps, G, right, True, True)
    order = np.argsort(tab["bucket"], kind="stable")
    d = synth.derive(c)
    buf, off, lens = marshal.stage_ascii([h.sequence.sequence for h in haps])
    fwd, rc = pam_patterns(pam)
    a, b = synth.scan_bounds(c, len(fwd))
    want = c_oracle.search(buf, off, lens, a, b, d.is_ref, d.seg, fwd, rc, G, right, threads=1)
    assert len(order) == len(want["hap"])
    for col in ("hap", "strand", "pos", "start", "stop"):
        assert np.array_equal(tab[col][order], want[col]), col
    assert np.array_equal(tab["text"][order], want["text"])


@pytest.mark.parametrize("seed,pam,G,right", [(1, "NGG", 20, False), (2, "TTTV", 23, True)])
def test_indexed_posmap_search_over_many_buckets(seed, pam, G, right):
    """A region of 60 kb with indel-rich haplotypes: ~15 buckets of the coarse segment index
    (4 kb each) and dozens of posmap segments per haplotype -- row_coords' indexed search
    (hawk_core.h, the table the device builds in seg_index_kernel) against the C oracle's
    coordinates for every row."""
    c = synth.make_cohort(60_000, 6, 1500, 300.0, seed=7000 + seed, snv_frac=0.3, ins_frac=0.35, max_indel=9)
    haps = synth.synth_haplotypes(c)
    region = synth.SynthRegion(c)
    d = synth.derive(c)
    n_seg = np.diff(d.seg.seg_off)
    assert n_seg[1:].min() > 50 and d.lens.max() >> 12 >= 14
    tab = hostcheck.search(pam, region, haps, G, right, True, True)
    order = np.argsort(tab["bucket"], kind="stable")
    buf, off, lens = marshal.stage_ascii([h.sequence.sequence for h in haps])
    fwd, rc = pam_patterns(pam)
    a, b = synth.scan_bounds(c, len(fwd))
    want = c_oracle.search(buf, off, lens, a, b, d.is_ref, d.seg, fwd, rc, G, right, threads=2)
    assert len(order) == len(want["hap"]) > 1000
    for col in ("hap", "strand", "pos", "start", "stop"):
        assert np.array_equal(tab[col][order], want[col]), col
    assert np.array_equal(tab["text"][order], want["text"])
