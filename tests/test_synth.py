"""The synthetic workload generator must produce haplotypes the reference's own builder
would produce from the equivalent phased VCF (live check), and the flat arrays derived
from it must agree with the oracle's view (CPU checks)."""

import numpy as np
import pytest

from crispr_hawk_b200 import marshal, synth
from crispr_hawk_b200.pam import pam_patterns
from oracle import c_oracle
from oracle import hawk_oracle as O
from oracle import refshim


def small_cohort(seed=3, n_alt=6):
    return synth.make_cohort(bed_len=1500, n_alt_hap=n_alt, n_sites=60, mean_alts_per_hap=12, seed=seed,
                             snv_frac=0.6, ins_frac=0.2, max_indel=6)  # fmt: skip


def test_derived_arrays_are_consistent():
    c = small_cohort()
    d = synth.derive(c)
    texts = synth.materialize_host(c)
    assert [len(t) for t in texts] == d.lens.tolist()
    assert texts[0].isupper() and d.is_ref.tolist() == [1] + [0] * (c.n_hap - 1)
    assert sum(sum(ch.islower() for ch in t) for t in texts) == d.variant_bases
    haps = synth.synth_haplotypes(c, texts)
    a, b = synth.scan_bounds(c, 3)
    for h, hap in enumerate(haps):
        # same bounds as the reference's dict-based rule
        assert marshal.scan_bounds(hap, c.region_start, c.region_stop, 3) == (a[h], b[h])
        # RLE of the dict posmap evaluates to the same map as the generator's segments
        r, g, s = marshal.posmap_segments(marshal.posmap_values(hap))
        vals = marshal.eval_segments(r, g, s, np.arange(len(hap)))
        assert vals.tolist() == [hap.posmap[i] for i in range(len(hap))]


@pytest.mark.ref
def test_generator_matches_reference_builder():
    c = small_cohort(seed=9, n_alt=8)
    lines, samples = synth.to_vcf_lines(c)
    region, ref_haps = refshim.build_case(c.ref.tobytes().decode(), c.bed_start, c.bed_stop, lines, samples, True)
    assert (region.start, region.stop) == (c.region_start, c.region_stop)
    want = {(h.sequence.sequence, tuple(h.posmap[i] for i in range(len(h)))) for h in ref_haps}
    mine = synth.synth_haplotypes(c)
    got = {(h.sequence.sequence, tuple(h.posmap[i] for i in range(len(h)))) for h in mine}
    assert got == want


@pytest.mark.parametrize("pam,G,right", [("NGG", 20, False), ("TTTV", 23, True), ("NNGRRT", 21, False)])
def test_c_oracle_equals_python_oracle_on_synth(pam, G, right):
    c = small_cohort(seed=21, n_alt=5)
    d = synth.derive(c)
    texts = synth.materialize_host(c)
    haps = synth.synth_haplotypes(c, texts)
    fwd, rc = pam_patterns(pam)
    a, b = synth.scan_bounds(c, len(fwd))
    buf, off, lens = marshal.stage_ascii(texts)
    assert off.tolist() == d.slot_off.tolist()
    out = c_oracle.search(buf, off, lens, a, b, d.is_ref, d.seg, fwd, rc, G, right, threads=2)
    want = O.search(pam, c.region_start, c.region_stop, [O.OracleHap.from_object(h) for h in haps], G, right, True, True)
    got = [(int(out["start"][i]), int(out["stop"][i]), int(out["strand"][i]), out["text"][i].tobytes().decode(),
            haps[int(out["hap"][i])].id) for i in range(len(out["hap"]))]  # fmt: skip
    assert got == [(g.start, g.stop, g.strand, g.sequence, g.hapid) for g in want]
    assert len(got) > 40
