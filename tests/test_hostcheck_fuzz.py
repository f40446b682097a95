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
